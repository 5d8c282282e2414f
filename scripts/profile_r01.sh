#!/bin/bash
# ncu evidence for round 1 (run under gpurun; outputs land in gpurun_out/, summaries are copied to profiles/)
set -x
CMD="python bench.py --steps 1 --warmup 1 --max-iterations 5 --no-cpu-baseline"
# 1. every launch with its device time (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r01_bench.csv $CMD > gpurun_out/ncu_bench.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches_r01_bench.csv > gpurun_out/launches_r01_bench_summary.txt
# 2. the top kernels, full sets (a few launches each)
for K in syrk_dmma_kernel eval_boards_kernel assemble_items_dmma_kernel potrf_diag_kernel schur_groups_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 4 -c 3 -f -o gpurun_out/prof_r01_$K $CMD > gpurun_out/ncu_$K.log 2>&1
done
ls -la gpurun_out/
