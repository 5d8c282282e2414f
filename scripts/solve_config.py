#!/usr/bin/env python3
"""Run one BASELINE config through the device-resident solver and print the solve info.
usage: solve_config.py [config=3] [max_iterations] [repeats]"""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import mrcal_b200
from mrcal_b200 import synthetic

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 3
maxit = int(sys.argv[2]) if len(sys.argv) > 2 else 300
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
kw, truth = synthetic.baseline_config(cfg, pixel_noise=0.3)
t0 = time.time()
P = mrcal_b200.Problem(**kw)
print("create %.3f s; Nstate %d Nmeas %d nnz %d" % (time.time() - t0, P.Nstate, P.Nmeasurements, P.N_j_nonzero))
for r in range(reps):
    P.reset()
    t0 = time.time()
    s = P.optimize(max_iterations=maxit)
    wall = time.time() - t0
    print(json.dumps({k: (round(v, 4) if isinstance(v, float) else v) for k, v in s.items()}), "wall %.3f s" % wall)
