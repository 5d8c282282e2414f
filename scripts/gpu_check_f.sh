#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/dbg_tri.py 2>&1 | grep -v "^\[\| \[" | tail -8
timeout 200 python scripts/solve_config.py 3 300 2 2>&1 | tail -1 | cut -c1-700
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | cut -c1-300 > gpurun_out/f_pytest.txt
tail -12 gpurun_out/f_pytest.txt
timeout 900 python bench.py 2>gpurun_out/f_bench.err | tee gpurun_out/f_bench.json | cut -c1-1500
tail -3 gpurun_out/f_bench.err
