"""Pins the oracle. The compiled reference (oracle/_ref) must reproduce the
reference's own golden vectors (test/test-optimizer-callback.py) and the
in-source projection known answers (test/test-projections.py:337-514); the
committed callback_cases fixtures must be what it produces today."""
import os

import numpy as np
import pytest

import problems

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")

# test/test-optimizer-callback.py:90-131
SELECTIONS = [
    dict(do_optimize_intrinsics_core=False, do_optimize_intrinsics_distortions=True, do_optimize_extrinsics=False,
         do_optimize_frames=False, do_optimize_calobject_warp=False, do_apply_regularization=True),
    dict(do_optimize_intrinsics_core=True, do_optimize_intrinsics_distortions=False, do_optimize_extrinsics=False,
         do_optimize_frames=False, do_optimize_calobject_warp=False, do_apply_regularization=True),
    dict(do_optimize_intrinsics_core=False, do_optimize_intrinsics_distortions=False, do_optimize_extrinsics=False,
         do_optimize_frames=True, do_optimize_calobject_warp=False, do_apply_regularization=True),
    dict(do_optimize_intrinsics_core=True, do_optimize_intrinsics_distortions=True, do_optimize_extrinsics=False,
         do_optimize_frames=True, do_optimize_calobject_warp=False, do_apply_regularization=True),
    dict(do_optimize_intrinsics_core=True, do_optimize_intrinsics_distortions=True, do_optimize_extrinsics=True,
         do_optimize_frames=True, do_optimize_calobject_warp=True, do_apply_regularization=False),
    dict(do_optimize_intrinsics_core=True, do_optimize_intrinsics_distortions=True, do_optimize_extrinsics=True,
         do_optimize_frames=True, do_optimize_calobject_warp=True, do_apply_regularization=False),
]


def optimizer_callback_golden_case(i):
    """(kwargs, x_ref, J_ref_unpacked) of case i of the reference's test-optimizer-callback.py."""
    g = np.load(os.path.join(GOLDEN, "optimizer_callback.npz"))
    kw = {k: g[k] for k in g.files if not k.startswith(("x_ref", "J_ref"))}
    kw["lensmodel"] = str(kw["lensmodel"])
    kw["calibration_object_spacing"] = float(kw["calibration_object_spacing"])
    kw.update(SELECTIONS[i])
    if i == 5:   # outlier_indices = (1,2): test-optimizer-callback.py:131,139-143
        o = kw["observations_board"].copy()
        o.reshape(-1, 3)[[1, 2], 2] = -1.
        kw["observations_board"] = o
    kw = {k: (np.ascontiguousarray(v) if isinstance(v, np.ndarray) else v) for k, v in kw.items()}
    return kw, g[f"x_ref_{i}"], g[f"J_ref_{i}"]


@pytest.mark.parametrize("i", range(6))
def test_compiled_reference_reproduces_reference_goldens(ref, i):
    kw, x_ref, J_ref = optimizer_callback_golden_case(i)
    P = ref.Problem(kw)
    b, x, J = P.callback()
    Jd = J.toarray()
    P.pack_vector(Jd)     # the goldens hold dx/d(unpacked state): test-optimizer-callback.py:177
    ireg = P.measurement_index("regularization")
    n = ireg if ireg >= 0 else len(x)
    assert x.shape == x_ref.shape and Jd.shape == J_ref.shape
    # the reference's own bar is 1e-6 RMS (test/testutils.py:113); the compiled code is exact to roundoff.
    # Rows >= ireg: the stored regularization rows predate the current scale constants
    assert np.abs(x[:n] - x_ref[:n]).max() < 1e-9
    assert np.abs(Jd[:n] - J_ref[:n]).max() < 1e-9 * (1 + np.abs(J_ref[:n]).max())


def test_compiled_reference_reproduces_projection_known_answers(ref):
    g = np.load(os.path.join(GOLDEN, "projections.npz"))
    n = 0
    for i in range(int(g["N"])):
        lm = str(g[f"lensmodel_{i}"])
        intr, p, q = g[f"intrinsics_{i}"], g[f"p_{i}"], g[f"q_{i}"]
        for k in range(p.shape[0]):
            ii = intr[k] if intr.ndim == 2 else intr
            qk = ref.project(p[k:k + 1], lm, ii)[0]
            assert np.abs(qk - q[k]).max() < 2e-6 * max(1., np.abs(q[k]).max()), (lm, k, qk, q[k])   # stored to ~10 digits
            n += 1
    assert n >= 30


def test_committed_callback_fixtures_are_current(ref):
    g = np.load(os.path.join(GOLDEN, "callback_cases.npz"))
    cases = problems.golden_cases()
    assert [c[0] for c in cases] == [str(s) for s in g["names"]]
    for name, kw in cases:
        P = ref.Problem(kw)
        b, x, J = P.callback()
        assert np.array_equal(g[f"{name}__Jp"], J.indptr) and np.array_equal(g[f"{name}__Ji"], J.indices), name
        assert np.allclose(g[f"{name}__x"], x, rtol=1e-12, atol=1e-12), name
        assert np.allclose(g[f"{name}__Jx"], J.data, rtol=1e-12, atol=1e-9), name
        assert np.allclose(g[f"{name}__b"], b, rtol=1e-15, atol=0), name
