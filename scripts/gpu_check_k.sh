#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_cross_reprojection_gpu.py -m gpu -q 2>&1 | cut -c1-400 > gpurun_out/k_pytest.txt
tail -40 gpurun_out/k_pytest.txt
