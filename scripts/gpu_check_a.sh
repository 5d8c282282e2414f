#!/bin/bash
# first GPU check of round 2: parity tests, then config 3 timing with both assembly paths
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/a_gpu.txt 2>&1
timeout 1700 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/a_pytest.txt
tail -5 gpurun_out/a_pytest.txt
for i in 1 2; do timeout 300 python scripts/solve_config.py 3 300 2 2>&1 | tail -3; done > gpurun_out/a_solve_det.txt 2>&1
MRCAL_B200_ATOMIC_ASSEMBLY=1 timeout 300 python scripts/solve_config.py 3 300 2 > gpurun_out/a_solve_atomic.txt 2>&1
cat gpurun_out/a_solve_det.txt gpurun_out/a_solve_atomic.txt | cut -c1-700
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_r02a.csv python bench.py --profile --steps 1 --warmup 1 --max-iterations 12 > gpurun_out/a_ncu.log 2>&1
tail -2 gpurun_out/a_ncu.log | cut -c1-300
