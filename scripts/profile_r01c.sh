#!/bin/bash
# ncu evidence, round 1, final pass of the round (persistent Cholesky with merged diagonal steps, forked assembly streams). Run under gpurun; outputs land in
# gpurun_out/, the summaries are copied to profiles/ by hand.
set -x
CMD="python bench.py --steps 1 --warmup 1 --max-iterations 10 --profile"
# 1. every launch with its device time (cold-cache, serialised: compare SHARES)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r01c.csv $CMD > gpurun_out/ncu_r01c_list.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches_r01c.csv > gpurun_out/launches_r01c_summary.txt
# 2. the top kernels, full sets (two launches each, after the warm-up solve)
for K in chol_dataflow_kernel chol_backward_dataflow_kernel schur_groups_kernel assemble_items_dmma_kernel eval_boards_kernel; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$K -s 12 -c 2 -f -o gpurun_out/prof_r01c_$K $CMD > gpurun_out/ncu_r01c_$K.log 2>&1
done
python scripts/ncu_summary.py gpurun_out/prof_r01c_*.ncu-rep > gpurun_out/ncu_r01c_summary.md
ls -la gpurun_out/ | tail -20
