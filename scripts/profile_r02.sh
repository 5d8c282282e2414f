#!/bin/bash
# ncu evidence, round 2. Run under gpurun; outputs land in gpurun_out/, the summaries are copied to profiles/.
# usage: profile_r02.sh <tag>   (tag names this pass: a, b, ...)
TAG=${1:-a}
CMD="python bench.py --steps 1 --warmup 1 --max-iterations 10 --profile --no-config5"
# 1. every launch with its device time (cold-cache, serialised: compare SHARES)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02$TAG.csv $CMD > gpurun_out/ncu_r02${TAG}_list.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches_r02$TAG.csv > gpurun_out/launches_r02${TAG}_summary.txt
# 2. the top kernels, full sets (two launches each, after the warm-up solve)
for K in schur_tiles_kernel chol_spine_kernel fused_boards_kernel groups_panels_kernel quadform_items_kernel; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$K -s 12 -c 2 -f -o gpurun_out/prof_r02${TAG}_$K $CMD > gpurun_out/ncu_r02${TAG}_$K.log 2>&1
done
# the Jacobian fill of optimizer_callback() (eval_boards_kernel with J): its own small driver
timeout 300 ncu --set full --clock-control none --import-source on -k regex:eval_boards_kernel -s 2 -c 2 -f -o gpurun_out/prof_r02${TAG}_eval_boards_kernel python scripts/callback_config.py 3 6 > gpurun_out/ncu_r02${TAG}_eval.log 2>&1
python scripts/ncu_summary.py gpurun_out/prof_r02${TAG}_*.ncu-rep > gpurun_out/ncu_r02${TAG}_summary.md
cat gpurun_out/launches_r02${TAG}_summary.txt | head -32
cat gpurun_out/ncu_r02${TAG}_summary.md
