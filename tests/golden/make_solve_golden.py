#!/usr/bin/env python3
"""Builds tests/golden/solve_cases.npz: what the reference's OWN mrcal_optimize() (mrcal.c:6179, compiled into
oracle/_ref with the restated libdogleg of oracle/port/dogleg_port.c underneath) returns for the problems of
tests/problems.py:solve_cases().

Run in the build container (needs /root/reference to build oracle/_ref):

    make -C oracle ref && python tests/golden/make_solve_golden.py [case ...]

Per case: b_packed (final packed state), rms_reproj_error__pixels, norm2_x, Noutliers_board,
Noutliers_triangulated_point, the indices of the board corners marked as outliers, the triangulated outlier
flags, the number of trust-region steps / evaluations / factorizations / outer passes, and a 64-entry sample
of x. The inputs are rebuilt from seeds by tests/problems.py wherever the repo is."""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import ref  # noqa: E402
import problems  # noqa: E402

OUT = os.path.join(HERE, "solve_cases.npz")


def main():
    want = set(sys.argv[1:])
    store = dict(np.load(OUT)) if os.path.exists(OUT) else {}
    for name, kw in problems.solve_cases():
        if want and name not in want:
            continue
        P = ref.Problem(kw)
        t0 = time.time()
        r = P.optimize()
        dt = time.time() - t0
        outl = np.flatnonzero(P.observations_board.reshape(-1, 3)[:, 2] < 0).astype(np.int32) if P.Nobs_board else np.zeros(0, np.int32)
        store[f"{name}/b_packed"] = r["b_packed"]
        store[f"{name}/scalars"] = np.array([r["rms_reproj_error__pixels"], r["norm2_x"]])
        store[f"{name}/counts"] = np.array([r["Noutliers_board"], r["Noutliers_triangulated_point"], r["iterations"],
                                            r["evaluations"], r["factorizations"], r["passes"], r["iterations_last_pass"]], np.int64)
        store[f"{name}/outliers_board"] = outl
        store[f"{name}/outliers_tri"] = getattr(P, "tri_outlier", np.zeros(0, np.int32))
        xs = np.linspace(0, len(r["x"]) - 1, 64).astype(np.int64)
        store[f"{name}/x_sample"] = r["x"][xs]
        print(f"{name}: {dt:.1f} s  rms {r['rms_reproj_error__pixels']:.9f}  iterations {r['iterations']} passes {r['passes']} "
              f"outliers {r['Noutliers_board']}/{r['Noutliers_triangulated_point']}  lambda {r['lambda_']:g}  "
              f"split callback {r['t_callback']:.2f} factor {r['t_factor']:.2f} products {r['t_products']:.2f} of {r['t_total']:.2f} s",
              flush=True)
        np.savez_compressed(OUT, **store)


if __name__ == "__main__":
    main()
