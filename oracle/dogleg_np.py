"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the solver half of the hot path.

The reference's trust-region loop is not in the reference tree: mrcal_optimize()
calls dogleg_optimize2() (mrcal.c:6435) from libdogleg (github.com/dkogan/libdogleg,
"at least version 0.15.3", doc/install.org:63), which factors JtJ with CHOLMOD.
Neither library is available here, so this module restates libdogleg's published
algorithm (Powell's dogleg, doc/formulation.org:293-294) with the parameters
mrcal sets (mrcal.c:6289-6299) and libdogleg's defaults for the rest:

    trustregion0 = 1e3, decrease 0.1 below rho < 0.25, increase 2 above rho > 0.75
    (only if the step was limited by the trust region), Jt_x_threshold = 0,
    update_threshold = 1e-7 (compared with the SQUARED step length),
    trustregion_threshold = 0, max_iterations = 300; on a non-positive-definite
    JtJ add lambda*I, lambda = 1e-10 then x10, kept for the rest of the solve.

PARITY UNPINNED at the iterate level: the reference publishes no iterate
sequence, iteration count or final state for any problem, and libdogleg's source
is not here to confirm the defaults above. What IS pinned: the cost function this
loop drives (oracle/_ref, checked against the reference's goldens), and the
linear solve (against numpy). The GPU solver is compared with this restatement at
the optimum.

Also restates markOutliers() (mrcal.c:3978-4402, board part) and the outer
re-solve loop (mrcal.c:6430-6481).
"""
import numpy as np
import scipy.sparse
import scipy.sparse.linalg

DEFAULTS = dict(max_iterations=300, trustregion0=1e3,
                trustregion_decrease_factor=0.1, trustregion_decrease_threshold=0.25,
                trustregion_increase_factor=2.0, trustregion_increase_threshold=0.75,
                Jt_x_threshold=0.0, update_threshold=1e-7, trustregion_threshold=0.0)


class NotPositiveDefinite(Exception):
    pass


def factor_solve(JtJ, rhs, dense_limit=3500):
    """Solve JtJ d = rhs by Cholesky (raises NotPositiveDefinite). The CPU
    stand-in for cholmod_factorize + cholmod_solve."""
    n = JtJ.shape[0]
    if n <= dense_limit:
        A = JtJ.toarray() if scipy.sparse.issparse(JtJ) else JtJ
        try:
            L = np.linalg.cholesky(A)
        except np.linalg.LinAlgError:
            raise NotPositiveDefinite()
        y = scipy.linalg.solve_triangular(L, rhs, lower=True)
        return scipy.linalg.solve_triangular(L.T, y, lower=False)
    lu = scipy.sparse.linalg.splu(scipy.sparse.csc_matrix(JtJ), permc_spec="MMD_AT_PLUS_A", diag_pivot_thresh=0.0,
                                  options=dict(SymmetricMode=True))
    if not (lu.U.diagonal() > 0).all():
        raise NotPositiveDefinite()
    return lu.solve(rhs)


import scipy.linalg  # noqa: E402


def dogleg_optimize(callback, p0, verbose=False, factor=factor_solve, **params):
    """callback(p) -> (x, J) with J scipy CSR. Returns dict(p, x, norm2_x, iterations,
    evaluations, factorizations, lambda_)."""
    par = dict(DEFAULTS)
    par.update(params)
    stats = dict(iterations=0, evaluations=0, factorizations=0)
    lam = 0.0

    class Point:
        pass

    def operating_point(p):
        pt = Point()
        pt.p = p
        pt.x, pt.J = callback(p)
        stats["evaluations"] += 1
        pt.norm2_x = float(pt.x @ pt.x)
        pt.Jt_x = pt.J.T @ pt.x
        pt.cauchy = None
        pt.gn = None
        return pt

    def cauchy(pt):
        if pt.cauchy is None:
            g2 = float(pt.Jt_x @ pt.Jt_x)
            Jg = pt.J @ pt.Jt_x
            k = -g2 / float(Jg @ Jg)
            pt.cauchy = k * pt.Jt_x
            pt.cauchy_lensq = k * k * g2
        return pt.cauchy

    def gauss_newton(pt):
        nonlocal lam
        if pt.gn is None:
            JtJ = (pt.J.T @ pt.J).tocsc()
            while True:
                try:
                    A = JtJ if lam == 0.0 else JtJ + lam * scipy.sparse.identity(JtJ.shape[0], format="csc")
                    stats["factorizations"] += 1
                    pt.gn = -factor(A, pt.Jt_x)
                    break
                except NotPositiveDefinite:
                    lam = 1e-10 if lam == 0.0 else lam * 10.0
                    if verbose:
                        print(f"singular JtJ. Adding {lam} I from now on")
            pt.gn_lensq = float(pt.gn @ pt.gn)
        return pt.gn

    def take_step(pt, trustregion):
        a = cauchy(pt)
        if pt.cauchy_lensq >= trustregion * trustregion:
            step = a * (trustregion / np.sqrt(pt.cauchy_lensq))
            lensq = trustregion * trustregion
            edge = True
        else:
            b = gauss_newton(pt)
            if pt.gn_lensq <= trustregion * trustregion:
                step, lensq, edge = b, pt.gn_lensq, False
            else:
                d = b - a
                l2 = float(d @ d)
                c = float(a @ d)
                disc = c * c - l2 * (pt.cauchy_lensq - trustregion * trustregion)
                k = (-c + np.sqrt(max(disc, 0.0))) / l2
                step = a + k * d
                lensq = trustregion * trustregion
                edge = True
        Js = pt.J @ step
        expected = -float(Js @ Js) - 2.0 * float(pt.x @ Js)
        return step, lensq, expected, edge

    trustregion = par["trustregion0"]
    before = operating_point(np.array(p0, dtype=float))
    done = not (np.abs(before.Jt_x) > par["Jt_x_threshold"]).any()
    while not done and stats["iterations"] < par["max_iterations"]:
        while True:
            step, lensq, expected, edge = take_step(before, trustregion)
            if lensq < par["update_threshold"]:
                done = True
                break
            after = operating_point(before.p + step)
            rho = (before.norm2_x - after.norm2_x) / expected
            if verbose:
                print(f"iter {stats['iterations']}: norm2_x {before.norm2_x:.9g} -> {after.norm2_x:.9g} "
                      f"|step| {np.sqrt(lensq):.3g} rho {rho:.3g} trustregion {trustregion:.3g}")
            if rho < par["trustregion_decrease_threshold"]:
                trustregion *= par["trustregion_decrease_factor"]
            elif rho > par["trustregion_increase_threshold"] and edge:
                trustregion *= par["trustregion_increase_factor"]
            if rho > 0.0:
                before = after
                break
            if trustregion < par["trustregion_threshold"]:
                done = True
                break
        if done:
            break
        stats["iterations"] += 1
        if not (np.abs(before.Jt_x) > par["Jt_x_threshold"]).any():
            break
    return dict(p=before.p, x=before.x, norm2_x=before.norm2_x, lambda_=lam, **stats)


def mark_outliers(observations_board, x_boards):
    """markOutliers(), board part (mrcal.c:4105-4357). observations_board (...,3) is
    modified in place (weights negated). Returns (found_new, Noutliers)."""
    k0, k1 = 4.0, 5.0
    pool = observations_board.reshape(-1, 3)
    w = pool[:, 2]
    inl = w > 0.0
    xx = x_boards.reshape(-1, 2)
    Nout = int((~inl).sum())
    var = float((xx[inl] ** 2).sum()) / (2.0 * int(inl.sum()))
    if not ((xx[inl] ** 2) > k1 * k1 * var).any():
        return False, Nout
    bad = inl & (((xx ** 2) > k0 * k0 * var).any(axis=1))
    w[bad] *= -1.0
    return True, Nout + int(bad.sum())


def optimize(kw, verbose=False, factor=factor_solve, **params):
    """CPU restatement of mrcal_optimize() (mrcal.c:6179-6624) on an optimization_inputs dict,
    with the compiled reference (oracle/_ref) as the cost function. Does not modify kw.
    Returns dict(b_packed, x, rms_reproj_error__pixels, Noutliers_board, + solver stats, problem)."""
    from . import ref
    P = ref.Problem(kw)
    Nstate = P.num_states()
    scale = np.ones(Nstate)
    P.unpack_vector(scale)          # scale[i] = unpacked value of a unit packed step

    def set_state(b):
        v = b * scale
        i = 0
        L = ref.lib()
        ni = P.num_states_of("intrinsics")
        if ni:
            per = ni // P.Ncam_i
            Nintr = P.intrinsics.shape[1]
            core = bool(P.selection_bits & 1) and True
            dist = bool(P.selection_bits & 2)
            for c in range(P.Ncam_i):
                blk = v[i:i + per]
                k = 0
                if core:
                    P.intrinsics[c, :4] = blk[:4]
                    k = 4
                if dist:
                    P.intrinsics[c, 4:] = blk[k:k + Nintr - 4]
                i += per
        ne = P.num_states_of("extrinsics")
        if ne:
            P.rt_cam_ref[...] = v[i:i + ne].reshape(-1, 6)
            i += ne
        nf = P.num_states_of("frames")
        if nf:
            P.rt_ref_frame[...] = v[i:i + nf].reshape(-1, 6)
            i += nf
        npt = P.num_states_of("points")
        if npt:
            P.points[:npt // 3] = v[i:i + npt].reshape(-1, 3)
            i += npt
        nw = P.num_states_of("calobject_warp")
        if nw:
            P.calobject_warp[...] = v[i:i + 2]
            i += 2
        assert i == Nstate

    def callback(b):
        set_state(b)
        _, x, J = P.callback()
        return x, J

    b0, _, _ = P.callback(no_jacobian=True)
    Noutliers = int((P.observations_board.reshape(-1, 3)[:, 2] < 0).sum()) if P.Nobs_board else 0
    outlier_rejection = bool(P.selection_bits & (1 << 6))
    passes = 0
    total = dict(iterations=0, evaluations=0, factorizations=0)
    b = b0
    while True:
        passes += 1
        r = dogleg_optimize(callback, b, verbose=verbose, factor=factor, **params)
        for k in total:
            total[k] += r[k]
        b = r["p"]
        if not (outlier_rejection and P.Nobs_board):
            break
        nb = P.num_measurements_of("boards")
        found, Noutliers = mark_outliers(P.observations_board, r["x"][:nb])
        if not found:
            break
    set_state(b)
    return dict(b_packed=b, x=r["x"], norm2_x=r["norm2_x"],
                rms_reproj_error__pixels=float(np.sqrt(r["norm2_x"] / len(r["x"]))),
                Noutliers_board=Noutliers, passes=passes, lambda_=r["lambda_"], problem=P, **total)
