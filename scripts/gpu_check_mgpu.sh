#!/bin/bash
# usage: gpu_check_mgpu.sh N   (run under gpurun --gpus N)
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 2>gpurun_out/mgpu${N}_bench.err | tee gpurun_out/mgpu${N}_bench.json | cut -c1-3000
tail -5 gpurun_out/mgpu${N}_bench.err | cut -c1-400
