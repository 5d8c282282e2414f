"""Solve-level parity against the reference's OWN mrcal_optimize() (mrcal.c:6179).

tests/golden/solve_cases.npz holds what the compiled reference (oracle/_ref: mrcal.c unmodified, with the
restated libdogleg of oracle/port/dogleg_port.c underneath) returned for tests/problems.py:solve_cases():
BASELINE configs 1-3 at full size exactly as bench.py solves them, configs 1-2 with gross outliers and outlier
rejection on, and small problems with discrete and triangulated points. Made by
tests/golden/make_solve_golden.py in the build container; the inputs are rebuilt from seeds here.

Gates: |b_packed - b_ref|_inf <= 1e-5, |rms - rms_ref| <= 1e-7 px, norm2_x relative 1e-9,
the same outlier set, the same number of outer passes. Where a gate is relaxed the reason is written next to it.
"""
import os

import numpy as np
import pytest

import problems

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "solve_cases.npz")


@pytest.fixture(scope="module")
def gold():
    if not os.path.exists(GOLD):
        pytest.skip("tests/golden/solve_cases.npz missing: run tests/golden/make_solve_golden.py")
    return np.load(GOLD)


@pytest.fixture(scope="module")
def cases():
    return dict(problems.solve_cases())


SMALL = ["baseline1", "baseline2", "baseline1_outliers", "baseline2_outliers", "splined3_outliers",
         "opencv4_points_outliers", "tri_pinhole_unity_only_rejection", "tri_opencv4_boards_points_rejection",
         "tri_stereographic_unity_rejection", "tri_divergent_rejection"]
ALL = SMALL + ["baseline3"]

# State gate per case. The splined solves end while still creeping along the flattest knot directions (the
# stopping rule is "squared step < 1e-7", mrcal.c:6297): the state there is defined by the iterate sequence,
# not by the cost, so roundoff-level differences in the factorization show up at ~1e-4 in those knots.
TOL_B = {"baseline3": 2e-3, "splined3_outliers": 2e-3,
         # triangulated points only, the scale held by the unity regularization alone: costs agree to 1e-6 relative when
         # the (absolute) step threshold stops both solves, the poorly constrained translations to ~3e-4
         "tri_divergent_rejection": 1e-3}
# Cost gate per case (relative). The triangulated-only problems end with costs of ~1e-6 rad^2; the stopping rule is an
# ABSOLUTE step length, so in relative terms they are less converged when the loop stops
TOL_COST = {"tri_pinhole_unity_only_rejection": 1e-6, "tri_stereographic_unity_rejection": 1e-6, "tri_divergent_rejection": 1e-6,
            "tri_opencv4_boards_points_rejection": 1e-8}


def _clone(kw):
    return {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in kw.items()}


@pytest.mark.parametrize("name", SMALL)
def test_oracle_reproduces_its_golden(ref, gold, cases, name):
    """CPU: the oracle's solve is deterministic and the committed fixture is what it produces today."""
    if f"{name}/b_packed" not in gold:
        pytest.skip("not in the fixture file")
    P = ref.Problem(_clone(cases[name]))
    r = P.optimize()
    assert np.abs(r["b_packed"] - gold[f"{name}/b_packed"]).max() <= 1e-9
    assert abs(r["rms_reproj_error__pixels"] - gold[f"{name}/scalars"][0]) <= 1e-12
    c = gold[f"{name}/counts"]
    assert (r["Noutliers_board"], r["Noutliers_triangulated_point"], r["iterations"], r["passes"]) == (c[0], c[1], c[2], c[5])


@pytest.mark.gpu
@pytest.mark.parametrize("name", ALL)
def test_optimize_matches_reference_solution(gold, cases, name):
    import mrcal_b200
    if f"{name}/b_packed" not in gold:
        pytest.skip("not in the fixture file")
    kw = _clone(cases[name])
    r = mrcal_b200.optimize(**kw)
    b_ref = gold[f"{name}/b_packed"]
    rms_ref, norm2_ref = gold[f"{name}/scalars"]
    c = gold[f"{name}/counts"]
    norm2 = float(r["x"] @ r["x"])
    assert abs(norm2 - norm2_ref) <= TOL_COST.get(name, 1e-9) * norm2_ref, (norm2, norm2_ref)
    assert abs(r["rms_reproj_error__pixels"] - rms_ref) <= 1e-7
    assert r["Noutliers_board"] == c[0]
    assert r["Noutliers_triangulated_point"] == c[1]
    if "observations_board" in kw:
        outl = np.flatnonzero(kw["observations_board"].reshape(-1, 3)[:, 2] < 0)
        assert np.array_equal(outl, gold[f"{name}/outliers_board"])
    assert np.abs(r["b_packed"] - b_ref).max() <= TOL_B.get(name, 1e-5), np.abs(r["b_packed"] - b_ref).max()
    xs = np.linspace(0, len(r["x"]) - 1, 64).astype(np.int64)
    assert np.abs(r["x"][xs] - gold[f"{name}/x_sample"]).max() <= (1e-4 if name in TOL_B else 1e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["baseline1", "baseline2", "baseline3", "baseline2_outliers"])
def test_iteration_counts_match_reference(gold, cases, name):
    """The trust-region step count of the device solver against the CPU run of the same algorithm. Not pinned by
    the reference itself (libdogleg publishes no iterate sequences), but a different count means a different
    accept/reject history somewhere."""
    import mrcal_b200
    if f"{name}/b_packed" not in gold:
        pytest.skip("not in the fixture file")
    P = mrcal_b200.Problem(**_clone(cases[name]))
    s = P.optimize()
    c = gold[f"{name}/counts"]
    # splined config 3: ~300 steps creeping along a flat valley with rho hovering around the 0.25 threshold;
    # a last-bit difference in rho flips one trust-region update and shifts the count by a few steps
    slack = 12 if name == "baseline3" else 0
    assert abs(s["Niterations"] - int(c[2])) <= slack, (s["Niterations"], int(c[2]))
    assert s["Nouter"] == int(c[5])
