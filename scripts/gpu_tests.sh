#!/bin/bash
# usage: gpurun -- 'bash scripts/gpu_tests.sh [pytest args]'
python -m pytest tests -m gpu -q "$@" 2>&1 | tail -40
