#!/usr/bin/env python3
"""N cost-function evaluations WITH the Jacobian of one BASELINE config on the device (the optimizer_callback() path).
usage: callback_config.py [config=3] [N=10]"""
import sys

sys.path.insert(0, ".")
import mrcal_b200
from mrcal_b200 import synthetic

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 3
N = int(sys.argv[2]) if len(sys.argv) > 2 else 10
kw, _ = synthetic.baseline_config(cfg, pixel_noise=0.3)
P = mrcal_b200.Problem(**kw)
print("ms per evaluation with J: %.4f, without: %.4f" % (P.time_callback(N, True), P.time_callback(N, False)))
