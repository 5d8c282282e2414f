#!/usr/bin/env python3
"""Builds tests/golden/*.npz from the reference's OWN test fixtures.

Run in the build container (needs /root/reference and oracle/_ref):

    python tests/golden/make_golden.py

What it stores:

optimizer_callback.npz
  The inputs of /root/reference/test/test-optimizer-callback.py (rebuilt here
  without numpysane/mrcal-python, following that script line by line, :45-88)
  and the six stored golden x / J arrays it compares against
  (test/data/test-optimizer-callback-ref-{x,J}-{0..5}.npy). J there is the
  Jacobian with respect to the UNPACKED state ("pack_state(J)", :177).

projections.npz
  The in-source known answers of /root/reference/test/test-projections.py
  (:337-514): (lensmodel, intrinsics, p, q_ref) for PINHOLE, STEREOGRAPHIC,
  LONLAT, LATLON, OPENCV4/5/8, SPLINED order 3 and 2. The numbers are parsed out
  of the reference test file, not retyped.

callback_cases.npz
  x, CSR J and layout numbers produced by the compiled reference
  (oracle/_ref/libmrcal_ref.so) on small seeded problems built by
  tests/problems.py: the same inputs can be regenerated anywhere, so the GPU
  box can check the CUDA path against the reference's outputs even if
  oracle/_ref were missing.
"""
import ast
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
REF = os.environ.get("MRCAL_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)

from oracle import ref  # noqa: E402


def read_cameramodel(path):
    txt = open(path).read()
    txt = re.sub(r"#.*", "", txt)
    return ast.literal_eval(txt.strip())


def linspace_shaped(*shape):
    return np.linspace(0, 1, int(np.prod(shape))).reshape(*shape)


def make_optimizer_callback():
    d = f"{REF}/test/data"
    rows = [l.split() for l in open(f"{d}/synthetic-board-observations.vnl") if not l.startswith("#")]
    names = sorted(set(r[0] for r in rows))
    obs = np.zeros((len(names), 10, 10, 3))
    idx = np.zeros((len(names), 3), np.int32)
    for i, n in enumerate(names):
        m = re.match(r"frame(\d+)-cam(\d+)", n)
        iframe, icam = int(m.group(1)), int(m.group(2))
        pts = np.array([[float(r[1]), float(r[2]), 0.5 ** float(r[3])] for r in rows if r[0] == n])
        obs[i] = pts.reshape(10, 10, 3)
        idx[i] = (iframe, icam, icam - 1)
    keep = (1, 2, 4, 5)  # test-optimizer-callback.py:51-54
    obs, idx = obs[keep, ...], idx[keep, ...]

    m0 = read_cameramodel(f"{d}/cam0.opencv8.cameramodel")
    m1 = read_cameramodel(f"{d}/cam1.opencv8.cameramodel")
    intrinsics = np.array([m0["intrinsics"], m1["intrinsics"]])
    rt_cam_ref = ref.compose_rt(np.array(m1["extrinsics"], float),
                                ref.invert_rt(np.array(m0["extrinsics"], float)))[None, :]
    imagersizes = np.array([m0["imagersize"], m1["imagersize"]], np.int32)
    rt_ref_frame = linspace_shaped(3, 6)
    rt_ref_frame[:, 5] += 5
    idx_pt = np.array(((0, 1, -1), (1, 0, -1), (1, 1, 0), (2, 0, -1), (2, 1, 0)), np.int32)
    points = 10. + 2. * linspace_shaped(3, 3)
    obs_pt = np.concatenate((1000. + 500. * linspace_shaped(5, 2),
                             np.array((0.9, 0.8, 0.9, 1.3, 1.8))[:, None]), axis=-1)
    out = dict(lensmodel=np.array(m0["lensmodel"]),
               intrinsics=intrinsics, rt_cam_ref=rt_cam_ref, imagersizes=imagersizes,
               rt_ref_frame=rt_ref_frame, points=points,
               observations_board=obs, indices_frame_camintrinsics_camextrinsics=idx,
               observations_point=obs_pt, indices_point_camintrinsics_camextrinsics=idx_pt,
               calobject_warp=np.array((1e-3, 2e-3)), calibration_object_spacing=np.array(0.1))
    for i in range(6):
        out[f"x_ref_{i}"] = np.load(f"{d}/test-optimizer-callback-ref-x-{i}.npy")
        out[f"J_ref_{i}"] = np.load(f"{d}/test-optimizer-callback-ref-J-{i}.npy")
    np.savez_compressed(f"{HERE}/optimizer_callback.npz", **out)
    print("optimizer_callback.npz:", {k: v.shape for k, v in out.items() if k.startswith(("x_", "J_"))})


def make_projections():
    """Parse the check(...) calls of test/test-projections.py:337-514."""
    src = open(f"{REF}/test/test-projections.py").read()
    tree = ast.parse(src)
    # names the reference test uses for its shared point array etc.
    env = {"np": np}
    cases = []
    for node in ast.walk(tree):
        if isinstance(node, ast.Assign) and len(node.targets) == 1 and isinstance(node.targets[0], ast.Name) \
           and node.targets[0].id == "p":
            try:
                env["p"] = eval(compile(ast.Expression(node.value), "<p>", "eval"), env)
            except Exception:
                pass
    for node in ast.walk(tree):
        if isinstance(node, ast.Call) and getattr(node.func, "id", None) == "check":
            try:
                args = [eval(compile(ast.Expression(a), "<a>", "eval"), env) for a in node.args]
            except Exception as e:
                print("skipping a check() call:", e)
                continue
            # check( (lensmodel, intrinsics), p, q_ref ): test-projections.py:44
            if len(args) >= 3 and isinstance(args[0], tuple) and isinstance(args[0][0], str):
                cases.append((args[0][0], args[0][1], args[1], args[2]))
    out = {}
    for i, (lensmodel, intr, p, q) in enumerate(cases):
        out[f"lensmodel_{i}"] = np.array(lensmodel)
        out[f"intrinsics_{i}"] = np.asarray(intr, float)
        out[f"p_{i}"] = np.asarray(p, float)
        out[f"q_{i}"] = np.asarray(q, float)
    out["N"] = np.array(len(cases))
    np.savez_compressed(f"{HERE}/projections.npz", **out)
    print("projections.npz:", [str(c[0])[:60] for c in cases])


def make_callback_cases():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import problems
    out = {}
    names = []
    for name, kw in problems.golden_cases():
        P = ref.Problem(kw)
        b, x, J = P.callback()
        names.append(name)
        out[f"{name}__b"] = b
        out[f"{name}__x"] = x
        out[f"{name}__Jp"] = J.indptr.astype(np.int32)
        out[f"{name}__Ji"] = J.indices.astype(np.int32)
        out[f"{name}__Jx"] = J.data
        out[f"{name}__layout"] = np.array(problems.layout_numbers(P), np.int64)
    out["names"] = np.array(names)
    np.savez_compressed(f"{HERE}/callback_cases.npz", **out)
    print("callback_cases.npz:", names)


if __name__ == "__main__":
    what = sys.argv[1:] or ["optimizer_callback", "projections", "callback_cases"]
    if "optimizer_callback" in what:
        make_optimizer_callback()
    if "projections" in what:
        make_projections()
    if "callback_cases" in what:
        make_callback_cases()
