"""The reference's gradient self-consistency test (test/test-gradients.py + test/test-gradients.c): every reported
Jacobian entry against a finite difference of the residuals, per lens model, through
mrcal_optimize(check_gradient=true) of the C-ABI (the reference does this with libdogleg's dogleg_testGradient)."""
import numpy as np
import pytest

import problems

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["pinhole_2cam_all", "stereographic_2cam_all", "cahvor_2cam_all", "opencv4_2cam_all",
                                  "splined3_3cam_all", "splined2_2cam_coreonly", "opencv8_points_fixed",
                                  "tri_opencv4_boards_points"])
def test_reported_gradients_match_finite_differences(name):
    from mrcal_b200 import api
    kw = dict(problems.golden_cases())[name]
    g = api.check_gradient(**kw)
    rep, obs = g[:, 2], g[:, 3]
    # forward difference with step 1e-6 in the packed state: second-order terms are ~1e-6 |d2x/db2|; the
    # reference's script accepts relative errors of a few percent on entries that are not tiny
    scale = np.maximum(np.abs(rep), np.abs(obs))
    big = scale > 1e-3 * scale.max()
    assert big.sum() > 100
    rel = np.abs(rep - obs)[big] / scale[big]
    assert np.percentile(rel, 99) < 2e-3, np.percentile(rel, 99)
    assert rel.max() < 5e-2, rel.max()
    # nothing reported as structurally zero has a real slope
    assert np.abs(obs[np.abs(rep) == 0]).max(initial=0.) < 1e-3 * scale.max()
