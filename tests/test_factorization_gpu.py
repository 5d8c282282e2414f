"""The factorization object (stand-in for mrcal.CHOLMOD_factorization). The
known-answer case is the reference's test/test-CHOLMOD-factorization.py; the
larger cases exercise the blocked DMMA Cholesky (several 64-blocks and
256-panels) against numpy."""
import numpy as np
import pytest
import scipy.sparse

import mrcal_b200

pytestmark = pytest.mark.gpu


def test_reference_known_answer():
    # test/test-CHOLMOD-factorization.py:20-50: a 4x3 J; solve_xt_JtJ_bt(bt) == solve(JtJ, bt')'
    indptr = np.array([0, 2, 3, 6, 8])
    indices = np.array([0, 2, 2, 0, 1, 2, 1, 2])
    data = np.array([1, 2, 3, 4, 5, 6, 7, 8], dtype=float)
    Jsparse = scipy.sparse.csr_matrix((data, indices, indptr))
    Jdense = Jsparse.toarray()
    Jdense_ref = np.array(((1, 0, 2), (0, 0, 3), (4, 5, 6), (0, 7, 8)), dtype=float)
    assert np.array_equal(Jdense, Jdense_ref)
    bt = np.array(((1., 5., 3.), (2., -2., -8)))
    F = mrcal_b200.CHOLMOD_factorization(Jsparse)
    xt = F.solve_xt_JtJ_bt(bt)
    JtJ = Jdense.T @ Jdense
    xt_ref = np.linalg.solve(JtJ, bt.T).T
    assert np.allclose(xt, xt_ref, rtol=1e-6, atol=0)       # the reference's bar
    assert np.allclose(xt, xt_ref, rtol=1e-12, atol=1e-14)  # ours
    # broadcasting over leading dims, and a single vector
    assert np.allclose(F.solve_xt_JtJ_bt(bt[0]), xt_ref[0], rtol=1e-12)
    assert np.allclose(F.solve_xt_JtJ_bt(bt[None, ...]), xt_ref[None, ...], rtol=1e-12)
    L = np.linalg.cholesky(JtJ)
    d = np.diag(L)
    assert np.isclose(F.rcond(), (d.min() / d.max()) ** 2, rtol=1e-10)
    assert np.allclose(F.solve_xt_JtJ_bt(bt, sys="L"), np.linalg.solve(L, bt.T).T, rtol=1e-12)


@pytest.mark.parametrize("n,m,density", [(50, 200, 0.2), (64, 300, 0.1), (200, 1500, 0.05), (700, 6000, 0.02),
                                         (1100, 9000, 0.01)])
def test_random_sparse(n, m, density):
    rng = np.random.default_rng(n)
    J = scipy.sparse.random(m, n, density=density, random_state=np.random.RandomState(n), format="csr")
    J = (J + scipy.sparse.vstack([scipy.sparse.eye(n) * 0.5, scipy.sparse.csr_matrix((m - n, n))])).tocsr()
    Jd = J.toarray()
    H = Jd.T @ Jd
    bt = rng.normal(size=(3, n))
    F = mrcal_b200.CHOLMOD_factorization(J)
    xt = F.solve_xt_JtJ_bt(bt)
    r = (H @ xt.T).T - bt
    assert np.abs(r).max() / np.abs(bt).max() < 1e-10   # ||JtJ d - Jtx|| / ||Jtx|| <= 1e-10
    assert np.allclose(xt, np.linalg.solve(H, bt.T).T, rtol=1e-7, atol=1e-9)


def test_not_positive_definite_gives_no_object():
    # a zero column makes JtJ singular; the reference returns None for the
    # factorization in that case (mrcal-pywrap.c:1981-1988)
    J = scipy.sparse.csr_matrix(np.array(((1., 0., 2.), (0., 0., 3.), (4., 0., 6.))))
    with pytest.raises(RuntimeError, match="positive definite"):
        mrcal_b200.CHOLMOD_factorization(J)


def test_optimizer_callback_returns_factorization():
    from mrcal_b200 import synthetic
    kw, _ = synthetic.make_problem(lensmodel="LENSMODEL_OPENCV4", Ncameras=2, Nframes=6, W=6, H=5)
    b, x, J, F = mrcal_b200.optimizer_callback(**kw)
    assert F is not None
    Jd = J.toarray()
    g = Jd.T @ x
    d = F.solve_xt_JtJ_bt(g)
    assert np.abs(Jd.T @ (Jd @ d) - g).max() / np.abs(g).max() < 1e-9


def test_solve_systems():
    """solve_xt_JtJ_bt(sys=...) (mrcal-pywrap.c:467-493). With P = I and D = I the factor L is THE Cholesky factor
    of JtJ (unique), so 'L' and 'Lt' are pinned by numpy; the identities the reference's callers rely on
    (mrcal/model_analysis.py:837-841: P, then L, then D, and A2 A3' = b' inv(JtJ) b) are checked as such."""
    rng = np.random.default_rng(5)
    J = scipy.sparse.random(400, 150, density=0.08, random_state=7, format="csr") + \
        scipy.sparse.vstack((scipy.sparse.eye(150), scipy.sparse.csr_matrix((250, 150))))
    F = mrcal_b200.CHOLMOD_factorization(J)
    JtJ = (J.T @ J).toarray()
    L = np.linalg.cholesky(JtJ)
    bt = rng.normal(size=(7, 150))
    tol = dict(rtol=0, atol=1e-9 * np.abs(bt).max())
    xA = F.solve_xt_JtJ_bt(bt)
    assert np.allclose(xA, np.linalg.solve(JtJ, bt.T).T, **tol)
    assert np.allclose(F.solve_xt_JtJ_bt(bt, sys="LDLt"), xA, **tol)
    for name in ("L", "LD", "CHOLMOD_L"):
        assert np.allclose(F.solve_xt_JtJ_bt(bt, sys=name), np.linalg.solve(L, bt.T).T, **tol)
    for name in ("Lt", "DLt"):
        assert np.allclose(F.solve_xt_JtJ_bt(bt, sys=name), np.linalg.solve(L.T, bt.T).T, **tol)
    for name in ("D", "P", "Pt"):
        assert np.array_equal(F.solve_xt_JtJ_bt(bt, sys=name), bt)
    # the chain the reference's uncertainty code runs
    A1 = F.solve_xt_JtJ_bt(bt, sys="P")
    A2 = F.solve_xt_JtJ_bt(A1, sys="L")
    A3 = F.solve_xt_JtJ_bt(A2, sys="D")
    assert np.allclose(A2 @ A3.T, bt @ np.linalg.solve(JtJ, bt.T), rtol=1e-9, atol=1e-12)
    # and Lt after L is A (after undoing P)
    assert np.allclose(F.solve_xt_JtJ_bt(F.solve_xt_JtJ_bt(F.solve_xt_JtJ_bt(A2, sys="D"), sys="Lt"), sys="Pt"), xA, **tol)
    with pytest.raises(RuntimeError, match="Unknown sys"):
        F.solve_xt_JtJ_bt(bt, sys="LL")
    # a single right-hand side takes the graph path
    assert np.allclose(F.solve_xt_JtJ_bt(bt[0], sys="L"), np.linalg.solve(L, bt[0]), **tol)
