#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_factorization_gpu.py tests/test_solve_golden.py tests/test_seeding.py tests/test_csr_ops_gpu.py tests/test_triangulated_precision.py tests/test_gradients_gpu.py -m gpu -q 2>&1 | tail -15 > gpurun_out/c_pytest.txt
tail -8 gpurun_out/c_pytest.txt
timeout 200 python scripts/solve_config.py 3 300 2 2>&1 | tail -2 | cut -c1-600
timeout 100 python scripts/callback_config.py 3 20
bash scripts/profile_r02.sh a 2>&1 | tail -45
