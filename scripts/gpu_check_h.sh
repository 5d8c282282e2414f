#!/bin/bash
mkdir -p gpurun_out
timeout 120 python scripts/spine_stamps.py 1270 | tail -5
timeout 600 python -m pytest tests/test_factorization_gpu.py tests/test_normal_gpu.py tests/test_optimize_gpu.py tests/test_solve_golden.py -m gpu -q 2>&1 | cut -c1-300 > gpurun_out/h_pytest.txt
tail -4 gpurun_out/h_pytest.txt
echo "== overlap"; timeout 200 python scripts/solve_config.py 3 300 3 2>&1 | tail -1 | cut -c1-700
echo "== no overlap"; MRCAL_B200_NO_OVERLAP=1 timeout 200 python scripts/solve_config.py 3 300 3 2>&1 | tail -1 | cut -c1-700
