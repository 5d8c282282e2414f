#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_normal_gpu.py tests/test_optimize_gpu.py tests/test_solve_golden.py tests/test_factorization_gpu.py -m gpu -q 2>&1 | cut -c1-300 > gpurun_out/i_pytest.txt
tail -4 gpurun_out/i_pytest.txt
timeout 200 python scripts/solve_config.py 3 300 2 2>&1 | tail -1 | cut -c1-700
timeout 300 python scripts/solve_config.py 5 60 2 2>&1 | tail -1 | cut -c1-700
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_cfg5.csv python scripts/solve_config.py 5 6 1 > /dev/null 2>&1
python scripts/summarize_launches.py gpurun_out/launches_cfg5.csv 2>/dev/null | head -12
