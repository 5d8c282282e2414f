#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/dbg_tri.py 2>&1 | tail -8
for cfg in "128 96" "96 48" "256 192" "160 400" "400 400" "64 64"; do
  set -- $cfg
  echo "== items/part $1 groups/part $2"
  MRCAL_B200_TILE_ITEMS_PER_PART=$1 MRCAL_B200_TILE_GROUPS_PER_PART=$2 timeout 200 python scripts/solve_config.py 3 300 2 2>&1 | tail -1 | cut -c1-700
done
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | cut -c1-300 > gpurun_out/e_pytest.txt
tail -12 gpurun_out/e_pytest.txt
