#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | cut -c1-300 > gpurun_out/f_pytest.txt
tail -6 gpurun_out/f_pytest.txt
timeout 900 python bench.py 2>gpurun_out/f_bench.err > gpurun_out/f_bench.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/f_bench.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','iterations_per_step','phase_ms_per_iteration','host_syncs_per_iteration','e2e','roofline','roofline_assembly','roofline_jacobian_fill','config5'): print(k, d.get(k))
PY
tail -3 gpurun_out/f_bench.err
