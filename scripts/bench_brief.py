#!/usr/bin/env python3
"""Run bench.py with the given arguments and print the handful of numbers watched during kernel work."""
import json, subprocess, sys
r = subprocess.run([sys.executable, "bench.py"] + sys.argv[1:], capture_output=True, text=True)
line = [l for l in r.stdout.splitlines() if l.startswith("{")]
if not line:
    print(r.stdout[-2000:], r.stderr[-2000:]); sys.exit(1)
d = json.loads(line[-1])
print("value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "it/step", d.get("iterations_per_step"), "rms", d.get("rms_reproj_error__pixels"))
print("phases", {k: round(v, 4) for k, v in d.get("phase_ms_per_iteration", {}).items()}, "ms/iter", round(1e3 / d["value"], 4))
print("roofline", {k: d["roofline"][k] for k in ("achieved", "frac")}, "launches", d.get("gpu_launches"))
