#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_factorization_gpu.py tests/test_csr_ops_gpu.py -m gpu -q 2>&1 | cut -c1-400 > gpurun_out/d_pytest.txt
tail -8 gpurun_out/d_pytest.txt
timeout 300 python scripts/dbg_tri.py 2>&1 | tail -30
