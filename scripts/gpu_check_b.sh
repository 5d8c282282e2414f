#!/bin/bash
# quick GPU check: the new code paths first (assembly, device loop, outliers, TMA stores), then config-3 timing
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > gpurun_out/b_gpu.txt 2>&1
timeout 900 python -m pytest tests/test_normal_gpu.py tests/test_optimize_gpu.py tests/test_solve_golden.py tests/test_csr_ops_gpu.py tests/test_triangulated_precision.py tests/test_gradients_gpu.py tests/test_project_gpu.py -m gpu -q -x 2>&1 | tail -25 > gpurun_out/b_pytest1.txt
tail -6 gpurun_out/b_pytest1.txt
timeout 600 python -m pytest tests/test_callback_gpu.py tests/test_factorization_gpu.py tests/test_seeding.py -m gpu -q 2>&1 | tail -15 > gpurun_out/b_pytest2.txt
tail -4 gpurun_out/b_pytest2.txt
for i in 1 2; do timeout 200 python scripts/solve_config.py 3 300 2 2>&1 | tail -2; done > gpurun_out/b_solve_det.txt 2>&1
MRCAL_B200_ATOMIC_ASSEMBLY=1 timeout 200 python scripts/solve_config.py 3 300 2 2>&1 | tail -2 > gpurun_out/b_solve_atomic.txt
cat gpurun_out/b_solve_det.txt gpurun_out/b_solve_atomic.txt | cut -c1-900
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_r02a.csv python bench.py --profile --steps 1 --warmup 1 --max-iterations 12 > gpurun_out/b_ncu.log 2>&1
tail -2 gpurun_out/b_ncu.log | cut -c1-300
python scripts/summarize_launches.py gpurun_out/launches_r02a.csv 2>&1 | head -40
