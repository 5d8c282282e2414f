#!/bin/bash
# usage: gpurun --gpus 2 -- 'bash scripts/dist_check.sh 2 small'
N=${1:-2}; WHICH=${2:-small}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/dist_check.py $WHICH 2>&1 | grep -v "^W0\|^\*\*\*\|Setting OMP" | tail -12
python - <<PY
import sys, numpy as np
sys.path.insert(0, ".")
import mrcal_b200
from mrcal_b200 import synthetic
which = "$WHICH"
if which == "small":
    kw, _ = synthetic.make_problem(lensmodel="LENSMODEL_SPLINED_STEREOGRAPHIC_order=3_Nx=8_Ny=6_fov_x_deg=100", Ncameras=2, Nframes=40, W=6, H=5, seed=2, pixel_noise=0.2)
elif which == "outliers":
    sys.path.insert(0, "tests")
    import problems
    kw, _ = synthetic.baseline_config(1, pixel_noise=0.3)
    problems.inject_gross_outliers(kw, 0.01, 101)
    kw["do_apply_outlier_rejection"] = True
elif which == "points":
    kw, _ = synthetic.make_problem(lensmodel="LENSMODEL_OPENCV4", Ncameras=3, Nframes=9, W=6, H=5, seed=4, pixel_noise=0.2, Npoints=12, Npoints_fixed=3, which="some")
else:
    kw, _ = synthetic.baseline_config(int(which), pixel_noise=0.3)
if which != "outliers":
    kw["do_apply_outlier_rejection"] = False
P = mrcal_b200.Problem(**kw)
s = P.optimize()
out = P.download(into_inputs=False)
g = np.load("/tmp/dist_check_solution.npz")
print("single GPU: outer passes", s["Nouter"], "outliers", s["Noutliers_board"], "iterations", s["Niterations"], "norm2", s["norm2_x_final"], "ms", round(s["ms_total"], 1), "| sharded: iterations", int(g["iterations"]), "norm2", float(g["norm2"]))
ok = abs(float(g["norm2"]) - s["norm2_x_final"]) <= 1e-8 * s["norm2_x_final"]
for k in ("intrinsics", "rt_cam_ref", "rt_ref_frame", "calobject_warp", "points"):
    if k in g.files and out.get(k) is not None and np.size(out[k]):
        d = np.abs(g[k] - out[k]).max()
        print(f"  {k}: max |sharded - single| = {d:.3g}")
        ok = ok and d < 1e-3 * max(1.0, np.abs(out[k]).max())
print("PARITY PASS" if ok else "PARITY FAIL")
PY
