#!/bin/bash
N=${1:-2}
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 scripts/nccl_time.py 2>&1 | grep "all-reduce\|Error\|error" | head -8; }
run
NCCL_ALGO=Tree run
NCCL_ALGO=NVLS run
NCCL_PROTO=LL128 run
