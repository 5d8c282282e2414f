#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_factorization_gpu.py tests/test_solve_golden.py tests/test_optimize_gpu.py -m gpu -q 2>&1 | cut -c1-300 > gpurun_out/g_pytest.txt
tail -6 gpurun_out/g_pytest.txt
echo "== spine"; timeout 200 python scripts/solve_config.py 3 300 2 2>&1 | tail -1 | cut -c1-700
echo "== no spine"; MRCAL_B200_CHOL_NO_SPINE=1 timeout 200 python scripts/solve_config.py 3 300 2 2>&1 | tail -1 | cut -c1-700
echo "== cfg5 spine"; timeout 300 python scripts/solve_config.py 5 40 2 2>&1 | tail -1 | cut -c1-700
echo "== cfg5 no spine"; MRCAL_B200_CHOL_NO_SPINE=1 timeout 300 python scripts/solve_config.py 5 40 2 2>&1 | tail -1 | cut -c1-700
