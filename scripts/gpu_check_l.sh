#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_callback_gpu.py tests/test_gradients_gpu.py tests/test_optimize_gpu.py tests/test_cross_reprojection_gpu.py tests/test_normal_gpu.py -m gpu -q 2>&1 | cut -c1-300 > gpurun_out/l_pytest.txt
tail -5 gpurun_out/l_pytest.txt
timeout 100 python scripts/callback_config.py 3 20
timeout 200 python scripts/solve_config.py 3 300 3 2>&1 | tail -1 | cut -c1-700
