#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:schur_tiles_kernel -s 3 -c 1 -f -o gpurun_out/prof_cfg5_schur_tiles python scripts/solve_config.py 5 6 1 > gpurun_out/ncu_cfg5.log 2>&1
tail -3 gpurun_out/ncu_cfg5.log
