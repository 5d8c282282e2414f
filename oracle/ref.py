"""TEST INFRASTRUCTURE ONLY -- ctypes driver for oracle/_ref/libmrcal_ref.so.

libmrcal_ref.so is the reference's OWN residual/Jacobian code (mrcal.c, opencv.c,
poseutils*.c/.cc, triangulation.cc, cahvore.cc) compiled by oracle/Makefile from
the sources where they lie under /root/reference. It is the ground truth for the
callback half of the hot path: x, the CSR Jacobian, the state/measurement
layout and pack/unpack. The solver half (libdogleg+CHOLMOD) is NOT in the
reference tree: oracle/port/dogleg_port.c restates libdogleg behind its own
interface and is linked into libmrcal_ref.so, so the reference's own
mrcal_optimize() (markOutliers, outer loop, unpack, statistics) runs:
Problem.optimize(). (oracle/dogleg_np.py is the same restatement in numpy.)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference legs
may import this module. The product (mrcal_b200/) never does.

ABI facts used below (x86-64 gcc):
  mrcal_problem_selections_t  1 byte passed by value; bit0 core, 1 distortions,
                              2 extrinsics, 3 frames, 4 calobject_warp,
                              5 regularization, 6 outlier rejection, 7 unity_cam01
                              (types.h:283-307)
  mrcal_lensmodel_t           16 bytes: int type @0, config union @8 (types.h:122-136)
  mrcal_observation_board_t   {int icam_intrinsics, icam_extrinsics, iframe} (types.h:206-214)
  mrcal_observation_point_t   {int icam_intrinsics, icam_extrinsics, i_point} (types.h:218-228)
  buffer sizes are in BYTES (mrcal.c:6083-6121)
"""
import ctypes as C
import os
import resource

import numpy as np
import scipy.sparse

_HERE = os.path.dirname(os.path.abspath(__file__))
LIBPATH = os.path.join(_HERE, "_ref", "libmrcal_ref.so")


def available():
    return os.path.exists(LIBPATH)


class Lensmodel(C.Structure):
    _fields_ = [("type", C.c_int), ("_pad", C.c_int), ("config", C.c_uint16 * 4)]


class Selections(C.Structure):
    _fields_ = [("bits", C.c_uint8)]


class Stats(C.Structure):
    """mrcal_stats_t (types.h:319-344)"""
    _fields_ = [("rms_reproj_error__pixels", C.c_double), ("Noutliers_board", C.c_int),
                ("Noutliers_triangulated_point", C.c_int)]


class CholmodSparse(C.Structure):
    _fields_ = [("nrow", C.c_size_t), ("ncol", C.c_size_t), ("nzmax", C.c_size_t),
                ("p", C.c_void_p), ("i", C.c_void_p), ("nz", C.c_void_p),
                ("x", C.c_void_p), ("z", C.c_void_p),
                ("stype", C.c_int), ("itype", C.c_int), ("xtype", C.c_int),
                ("dtype", C.c_int), ("sorted", C.c_int), ("packed", C.c_int)]


SELECTION_BITS = ("do_optimize_intrinsics_core",
                  "do_optimize_intrinsics_distortions",
                  "do_optimize_extrinsics",
                  "do_optimize_frames",
                  "do_optimize_calobject_warp",
                  "do_apply_regularization",
                  "do_apply_outlier_rejection",
                  "do_apply_regularization_unity_cam01")

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError(f"{LIBPATH} is missing: run `make -C oracle ref` where /root/reference exists")
        # The reference keeps big VLAs on the stack (mrcal.c:4534,4645-4663)
        try:
            soft, hard = resource.getrlimit(resource.RLIMIT_STACK)
            resource.setrlimit(resource.RLIMIT_STACK, (hard, hard))
        except Exception:
            pass
        _lib = C.CDLL(LIBPATH)
        _lib.mrcal_lensmodel_from_name.restype = C.c_bool
        _lib.mrcal_optimizer_callback.restype = C.c_bool
        _lib.mrcal_project.restype = C.c_bool
        for n in ("mrcal_lensmodel_num_params", "mrcal_num_states", "mrcal_num_measurements",
                  "_mrcal_num_j_nonzero", "mrcal_num_intrinsics_optimization_params",
                  "mrcal_state_index_intrinsics", "mrcal_state_index_extrinsics",
                  "mrcal_state_index_frames", "mrcal_state_index_points",
                  "mrcal_state_index_calobject_warp",
                  "mrcal_num_states_intrinsics", "mrcal_num_states_extrinsics",
                  "mrcal_num_states_frames", "mrcal_num_states_points",
                  "mrcal_num_states_calobject_warp",
                  "mrcal_measurement_index_boards", "mrcal_measurement_index_points",
                  "mrcal_measurement_index_regularization",
                  "mrcal_num_measurements_boards", "mrcal_num_measurements_points",
                  "mrcal_num_measurements_regularization"):
            getattr(_lib, n).restype = C.c_int
    return _lib


def lensmodel_from_name(name):
    lm = Lensmodel()
    if not lib().mrcal_lensmodel_from_name(C.byref(lm), name.encode()):
        raise ValueError(f"reference could not parse lensmodel '{name}'")
    return lm


def lensmodel_num_params(name):
    return lib().mrcal_lensmodel_num_params(C.byref(lensmodel_from_name(name)))


def _dp(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None and a.size else None


class ObservationPointTriangulated(C.Structure):
    """mrcal_observation_point_triangulated_t (types.h:243-263); bits: 1 = last_in_set, 2 = outlier."""
    _fields_ = [("icam_intrinsics", C.c_int), ("icam_extrinsics", C.c_int),
                ("bits", C.c_uint8), ("px", C.c_double * 3)]


def unproject(q, lensmodel_name, intrinsics):
    """Observation rays of pixels q (N,2). Closed-form models: the reference's mrcal_unproject (mrcal.c:3082).
    The others: the reference inverts them with libdogleg, which this build stubs out, so the same fixed point
    is found here by Newton's method on the reference's OWN mrcal_project() and its gradients, in the same
    stereographic parametrisation (mrcal.c:3106-3270)."""
    q = np.ascontiguousarray(q, dtype=np.float64).reshape(-1, 2)
    intr = np.ascontiguousarray(intrinsics, dtype=np.float64)
    if lensmodel_name in ("LENSMODEL_PINHOLE", "LENSMODEL_STEREOGRAPHIC", "LENSMODEL_LONLAT", "LENSMODEL_LATLON"):
        lm = lensmodel_from_name(lensmodel_name)
        v = np.zeros((q.shape[0], 3))
        lib().mrcal_unproject.restype = C.c_bool
        if not lib().mrcal_unproject(_dp(v), _dp(q), q.shape[0], C.byref(lm), _dp(intr)):
            raise RuntimeError("reference mrcal_unproject() failed")
        return v
    xy = (q - intr[2:4]) / intr[0:2]
    u = 2. * xy / (np.sqrt((xy * xy).sum(-1, keepdims=True) + 1.) + 1.)
    for _ in range(50):
        v = np.concatenate((u, 1. - 0.25 * (u * u).sum(-1, keepdims=True)), -1)
        qq, g = project(v, lensmodel_name, intr, gradients=True)
        r = qq - q
        J = g[..., :2] - 0.5 * g[..., 2:3] * u[:, None, :]
        du = -np.einsum("nij,nj->ni", np.linalg.pinv(J), r)
        u = u + du
        if np.abs(du).max() < 1e-15:
            break
    v = np.concatenate((u, 1. - 0.25 * (u * u).sum(-1, keepdims=True)), -1)
    return v


class Problem:
    """Normalised view of a mrcal `optimization_inputs` dict (mrcal-pywrap.c:890-937).

    Triangulated points: `observations_point_triangulated` (N,3: pixel x, y, weight) with
    `indices_point_triangulated_camintrinsics_camextrinsics` (N,3) as in the reference's Python API; the rays
    come from unproject() above, as mrcal-pywrap.c:1383-1401 does it."""

    def __init__(self, kw):
        kw = dict(kw)
        for old, new in (("extrinsics_rt_fromref", "rt_cam_ref"), ("frames_rt_toref", "rt_ref_frame")):
            if old in kw and new not in kw:
                kw[new] = kw[old]
        f8 = lambda k, shape: np.ascontiguousarray(
            kw[k] if kw.get(k) is not None else np.zeros(shape), dtype=np.float64)
        i4 = lambda k, shape: np.ascontiguousarray(
            kw[k] if kw.get(k) is not None else np.zeros(shape), dtype=np.int32)
        self.lensmodel_name = kw["lensmodel"]
        self.lensmodel = lensmodel_from_name(self.lensmodel_name)
        self.intrinsics = f8("intrinsics", (0, 0)).copy()
        self.imagersizes = i4("imagersizes", (0, 2))
        self.rt_cam_ref = f8("rt_cam_ref", (0, 6)).copy()
        self.rt_ref_frame = f8("rt_ref_frame", (0, 6)).copy()
        self.points = f8("points", (0, 3)).copy()
        self.observations_board = f8("observations_board", (0, 0, 0, 3)).copy()
        self.indices_board = i4("indices_frame_camintrinsics_camextrinsics", (0, 3))
        self.observations_point = f8("observations_point", (0, 3)).copy()
        self.indices_point = i4("indices_point_camintrinsics_camextrinsics", (0, 3))
        self.observations_tri = f8("observations_point_triangulated", (0, 3)).copy()
        self.indices_tri = i4("indices_point_triangulated_camintrinsics_camextrinsics", (0, 3))
        self.Nobs_tri = self.indices_tri.shape[0]
        self.c_tri = None
        if self.Nobs_tri:
            rays = np.zeros((self.Nobs_tri, 3))
            for icam in np.unique(self.indices_tri[:, 1]):
                sel = np.flatnonzero(self.indices_tri[:, 1] == icam)
                rays[sel] = unproject(self.observations_tri[sel, :2], self.lensmodel_name, self.intrinsics[icam])
            self.triangulated_rays = rays
            last = np.concatenate((np.diff(self.indices_tri[:, 0]) != 0, [True]))
            self.c_tri = (ObservationPointTriangulated * self.Nobs_tri)()
            for i in range(self.Nobs_tri):
                o = self.c_tri[i]
                o.icam_intrinsics, o.icam_extrinsics = int(self.indices_tri[i, 1]), int(self.indices_tri[i, 2])
                o.bits = (1 if last[i] else 0) | (2 if self.observations_tri[i, 2] <= 0.0 else 0)
                o.px[0], o.px[1], o.px[2] = rays[i]
        cw = kw.get("calobject_warp")
        self.calobject_warp = None if cw is None else np.ascontiguousarray(cw, dtype=np.float64).copy()
        self.Npoints_fixed = int(kw.get("Npoints_fixed", 0) or 0)
        self.spacing = float(kw.get("calibration_object_spacing", 0.0) or 0.0)
        self.Ncam_i = self.intrinsics.shape[0]
        self.Ncam_e = self.rt_cam_ref.shape[0]
        self.Nframes = self.rt_ref_frame.shape[0]
        self.Npoints = self.points.shape[0]
        self.Nobs_board = self.indices_board.shape[0]
        self.Nobs_point = self.indices_point.shape[0]
        if self.Nobs_board:
            self.H, self.W = self.observations_board.shape[1:3]
        else:
            self.H = self.W = 0

        # do_optimize_* defaults: "optimize if that thing exists" (mrcal-pywrap.c:1447-1457)
        d = dict(do_optimize_intrinsics_core=self.Ncam_i > 0,
                 do_optimize_intrinsics_distortions=self.Ncam_i > 0,
                 do_optimize_extrinsics=self.Ncam_e > 0,
                 do_optimize_frames=self.Nframes > 0,
                 do_optimize_calobject_warp=self.Nobs_board > 0,
                 do_apply_regularization=True,
                 do_apply_outlier_rejection=True,
                 do_apply_regularization_unity_cam01=False)
        bits = 0
        for ib, name in enumerate(SELECTION_BITS):
            v = kw.get(name)
            v = d[name] if v is None else bool(v)
            bits |= (1 << ib) if v else 0
        self.selection_bits = bits
        self.selections = Selections(bits)

        # struct order differs from the Python index arrays (mrcal-pywrap.c:1255-1261)
        self.c_obs_board = np.ascontiguousarray(self.indices_board[:, (1, 2, 0)]) \
            if self.Nobs_board else np.zeros((0, 3), np.int32)
        self.c_obs_point = np.ascontiguousarray(self.indices_point[:, (1, 2, 0)]) \
            if self.Nobs_point else np.zeros((0, 3), np.int32)

    # The layout functions take the raw bits: has_calobject_warp() (mrcal.c:737)
    # and modelHasCore_fxfycxcy() (mrcal.c:291; true for every model) apply the
    # same normalisation the callback does at mrcal.c:6055-6060
    def effective_selections(self):
        return Selections(self.selection_bits)

    def counts(self):
        return (self.Ncam_i, self.Ncam_e, self.Nframes, self.Npoints, self.Npoints_fixed, self.Nobs_board)

    def num_states(self):
        return lib().mrcal_num_states(*self.counts(), self.effective_selections(), C.byref(self.lensmodel))

    def num_measurements(self):
        return lib().mrcal_num_measurements(
            self.Nobs_board, self.Nobs_point, self.c_tri, self.Nobs_tri, self.W, self.H,
            self.Ncam_i, self.Ncam_e, self.Nframes, self.Npoints, self.Npoints_fixed,
            self.effective_selections(), C.byref(self.lensmodel))

    def num_j_nonzero(self):
        return lib()._mrcal_num_j_nonzero(
            self.Nobs_board, self.Nobs_point, self.c_tri, self.Nobs_tri, self.W, self.H,
            self.Ncam_i, self.Ncam_e, self.Nframes, self.Npoints, self.Npoints_fixed,
            _dp(self.c_obs_board), _dp(self.c_obs_point),
            self.effective_selections(), C.byref(self.lensmodel))

    def state_index(self, what, i=0):
        f = getattr(lib(), f"mrcal_state_index_{what}")
        args = self.counts() + (self.effective_selections(), C.byref(self.lensmodel))
        if what == "calobject_warp":
            return f(*args)
        return f(i, *args)

    def num_states_of(self, what):
        l, s = lib(), self.effective_selections()
        if what == "intrinsics":
            return l.mrcal_num_states_intrinsics(self.Ncam_i, s, C.byref(self.lensmodel))
        if what == "extrinsics":
            return l.mrcal_num_states_extrinsics(self.Ncam_e, s)
        if what == "frames":
            return l.mrcal_num_states_frames(self.Nframes, s)
        if what == "points":
            return l.mrcal_num_states_points(self.Npoints, self.Npoints_fixed, s)
        if what == "calobject_warp":
            return l.mrcal_num_states_calobject_warp(s, self.Nobs_board)
        raise KeyError(what)

    def measurement_index(self, what, i=0):
        l, s = lib(), self.effective_selections()
        if what == "boards":
            return l.mrcal_measurement_index_boards(i, self.Nobs_board, self.Nobs_point, self.W, self.H)
        if what == "points":
            return l.mrcal_measurement_index_points(i, self.Nobs_board, self.Nobs_point, self.W, self.H)
        if what == "regularization":
            return l.mrcal_measurement_index_regularization(
                self.c_tri, self.Nobs_tri, self.W, self.H, self.Ncam_i, self.Ncam_e, self.Nframes,
                self.Npoints, self.Npoints_fixed, self.Nobs_board, self.Nobs_point,
                s, C.byref(self.lensmodel))
        raise KeyError(what)

    def num_measurements_of(self, what):
        l, s = lib(), self.effective_selections()
        if what == "boards":
            return l.mrcal_num_measurements_boards(self.Nobs_board, self.W, self.H)
        if what == "points":
            return l.mrcal_num_measurements_points(self.Nobs_point)
        if what == "regularization":
            return l.mrcal_num_measurements_regularization(
                self.Ncam_i, self.Ncam_e, self.Nframes, self.Npoints, self.Npoints_fixed,
                self.Nobs_board, s, C.byref(self.lensmodel))
        raise KeyError(what)

    def callback(self, no_jacobian=False):
        """One evaluation of the reference's mrcal_optimizer_callback (mrcal.c:5972).

        Returns (b_packed, x, J) with J a scipy CSR matrix (or None)."""
        Nstate, Nmeas = self.num_states(), self.num_measurements()
        b = np.zeros(Nstate)
        x = np.zeros(Nmeas)
        Jt, keep = None, None
        if not no_jacobian:
            nnz = self.num_j_nonzero()
            P = np.zeros(Nmeas + 1, np.int32)
            I = np.zeros(nnz, np.int32)
            X = np.zeros(nnz, np.float64)
            Jt = CholmodSparse(nrow=Nstate, ncol=Nmeas, nzmax=nnz,
                               p=P.ctypes.data, i=I.ctypes.data, x=X.ctypes.data,
                               sorted=1, packed=1)
            keep = (P, I, X)
        ok = lib().mrcal_optimizer_callback(
            _dp(b), C.c_int(b.nbytes), _dp(x), C.c_int(x.nbytes),
            C.byref(Jt) if Jt is not None else None,
            _dp(self.intrinsics), _dp(self.rt_cam_ref), _dp(self.rt_ref_frame), _dp(self.points),
            _dp(self.calobject_warp) if self.calobject_warp is not None else None,
            self.Ncam_i, self.Ncam_e, self.Nframes, self.Npoints, self.Npoints_fixed,
            _dp(self.c_obs_board), _dp(self.c_obs_point), self.Nobs_board, self.Nobs_point,
            self.c_tri, self.Nobs_tri,
            _dp(self.observations_board), _dp(self.observations_point),
            C.byref(self.lensmodel), _dp(self.imagersizes),
            self.selections, None,
            C.c_double(self.spacing), self.W, self.H, C.c_bool(False))
        if not ok:
            raise RuntimeError("reference mrcal_optimizer_callback() failed")
        J = None
        if keep is not None:
            P, I, X = keep
            J = scipy.sparse.csr_matrix((X, I, P), shape=(Nmeas, Nstate))
        return b, x, J

    def drt_cross_reprojection__dbpacked(self, icam_intrinsics=-1):
        """K = drt_ref_refperturbed/db_packed (icam_intrinsics < 0) or drt_cam_camperturbed/db_packed, shape (6, Nstate),
        as mrcal.drt_cross_reprojection__dbpacked() returns it (mrcal-pywrap.c:2016-2110): the callback at the current
        state, then _mrcal_drt_cross_reprojection__dbpacked() (uncertainty.c:798) into the extrinsics / frames / points /
        calobject_warp columns of a zero matrix."""
        b, x, J = self.callback()
        Nstate, Nmeas = J.shape[1], J.shape[0]
        P = np.ascontiguousarray(J.indptr, np.int32)
        I = np.ascontiguousarray(J.indices, np.int32)
        X = np.ascontiguousarray(J.data, np.float64)
        Jt = CholmodSparse(nrow=Nstate, ncol=Nmeas, nzmax=len(X), p=P.ctypes.data, i=I.ctypes.data, x=X.ctypes.data,
                           sorted=1, packed=1)
        K = np.zeros((6, Nstate))
        s0, s1 = K.strides

        def sub(what):
            i0 = self.state_index(what) if what == "calobject_warp" else self.state_index(what, 0)
            if i0 < 0 or self.num_states_of(what) == 0:
                return None
            return C.c_void_p(K.ctypes.data + 8 * i0)
        f = lib()._mrcal_drt_cross_reprojection__dbpacked
        f.restype = C.c_bool
        ok = f(sub("extrinsics"), C.c_int(s0), C.c_int(s1), sub("frames"), C.c_int(s0), C.c_int(s1),
               sub("points"), C.c_int(s0), C.c_int(s1), sub("calobject_warp"), C.c_int(s0), C.c_int(s1),
               C.c_int(icam_intrinsics), _dp(b), C.c_int(b.nbytes), C.byref(Jt),
               self.Ncam_i, self.Ncam_e, self.Nframes, self.Npoints, self.Npoints_fixed, self.Nobs_board, self.Nobs_point,
               C.byref(self.lensmodel), self.effective_selections(), self.W, self.H)
        if not ok:
            raise RuntimeError("reference _mrcal_drt_cross_reprojection__dbpacked() failed")
        return K, b, J

    def optimize(self, verbose=False, iteration_cap=0):
        """The reference's own mrcal_optimize() (mrcal.c:6179) on top of the restated libdogleg
        (oracle/port/dogleg_port.c). Modifies this Problem's state arrays and observation weights in
        place, as the reference does. iteration_cap > 0 bounds the steps of each dogleg_optimize2() call
        (benchmark samples). Returns dict(b_packed, x, rms_reproj_error__pixels, Noutliers_board,
        Noutliers_triangulated_point, + the solver's counters and timers)."""
        L = lib()
        L.mrcal_optimize.restype = Stats
        L.dogleg_port_set_iteration_cap(C.c_int(int(iteration_cap)))
        L.dogleg_port_reset_stats()
        Nstate, Nmeas = self.num_states(), self.num_measurements()
        b = np.zeros(Nstate)
        x = np.zeros(Nmeas)
        st = L.mrcal_optimize(
            _dp(b), C.c_int(b.nbytes), _dp(x), C.c_int(x.nbytes),
            _dp(self.intrinsics), _dp(self.rt_cam_ref), _dp(self.rt_ref_frame), _dp(self.points),
            _dp(self.calobject_warp) if self.calobject_warp is not None else None,
            self.Ncam_i, self.Ncam_e, self.Nframes, self.Npoints, self.Npoints_fixed,
            _dp(self.c_obs_board), _dp(self.c_obs_point), self.Nobs_board, self.Nobs_point,
            self.c_tri, self.Nobs_tri,
            _dp(self.observations_board), _dp(self.observations_point),
            C.byref(self.lensmodel), _dp(self.imagersizes),
            self.selections, None,
            C.c_double(self.spacing), self.W, self.H, C.c_bool(bool(verbose)), C.c_bool(False))
        L.dogleg_port_set_iteration_cap(C.c_int(0))
        if st.rms_reproj_error__pixels < 0:
            raise RuntimeError("reference mrcal_optimize() failed")
        o = (C.c_double * 16)()
        L.dogleg_port_get_stats(o)
        if self.c_tri is not None:
            self.tri_outlier = np.array([(self.c_tri[i].bits >> 1) & 1 for i in range(self.Nobs_tri)], np.int32)
        return dict(b_packed=b, x=x, rms_reproj_error__pixels=float(st.rms_reproj_error__pixels),
                    Noutliers_board=int(st.Noutliers_board),
                    Noutliers_triangulated_point=int(st.Noutliers_triangulated_point),
                    norm2_x=float(x @ x),
                    iterations=int(o[0]), evaluations=int(o[1]), factorizations=int(o[2]),
                    symbolic=int(o[3]), Lnnz=int(o[4]), passes=int(o[5]),
                    t_callback=o[6], t_factor=o[7], t_products=o[8], t_total=o[9], lambda_=o[10],
                    iterations_last_pass=int(o[12]))

    def pack_vector(self, b):
        b = np.ascontiguousarray(b, dtype=np.float64)
        flat = b.reshape(-1, b.shape[-1])
        for row in flat:
            lib().mrcal_pack_solver_state_vector(_dp(row), *self.counts(),
                                                 self.effective_selections(), C.byref(self.lensmodel))
        return b

    def unpack_vector(self, b):
        b = np.ascontiguousarray(b, dtype=np.float64)
        flat = b.reshape(-1, b.shape[-1])
        for row in flat:
            lib().mrcal_unpack_solver_state_vector(_dp(row), *self.counts(),
                                                   self.effective_selections(), C.byref(self.lensmodel))
        return b


def project_with_intrinsics_gradient(p, lensmodel_name, intrinsics):
    """Reference mrcal_project with both gradients: q (N,2), dq_dp (N,2,3), dq_dintrinsics (N,2,Nintrinsics)."""
    lm = lensmodel_from_name(lensmodel_name)
    p = np.ascontiguousarray(p, dtype=np.float64).reshape(-1, 3)
    intrinsics = np.ascontiguousarray(intrinsics, dtype=np.float64)
    N = p.shape[0]
    q = np.zeros((N, 2))
    dq_dp = np.zeros((N, 2, 3))
    dq_di = np.zeros((N, 2, intrinsics.shape[0]))
    if not lib().mrcal_project(_dp(q), _dp(dq_dp), _dp(dq_di), _dp(p), N, C.byref(lm), _dp(intrinsics)):
        raise RuntimeError("reference mrcal_project() failed")
    return q, dq_dp, dq_di


def project(p, lensmodel_name, intrinsics, gradients=False):
    """Reference mrcal_project (mrcal.h:165). p: (N,3). Returns q (N,2) [, dq_dp (N,2,3)]."""
    lm = lensmodel_from_name(lensmodel_name)
    p = np.ascontiguousarray(p, dtype=np.float64).reshape(-1, 3)
    intrinsics = np.ascontiguousarray(intrinsics, dtype=np.float64)
    N = p.shape[0]
    q = np.zeros((N, 2))
    if not gradients:
        if not lib().mrcal_project(_dp(q), None, None, _dp(p), N, C.byref(lm), _dp(intrinsics)):
            raise RuntimeError("reference mrcal_project() failed")
        return q
    # one point per call: with N > 1 and no intrinsics gradients the reference writes every point's dq_dp
    # into the first slot (mrcal.c:2894-2907 passes dq_dp, not &dq_dp[2*i])
    dq_dp = np.zeros((N, 2, 3))
    for i in range(N):
        qi, gi, pi = np.zeros(2), np.zeros((2, 3)), np.ascontiguousarray(p[i])
        if not lib().mrcal_project(_dp(qi), _dp(gi), None, _dp(pi), 1, C.byref(lm), _dp(intrinsics)):
            raise RuntimeError("reference mrcal_project() failed")
        q[i], dq_dp[i] = qi, gi
    return q, dq_dp


def compose_rt(rt0, rt1):
    """Reference mrcal_compose_rt_full (poseutils.c:745); no gradients."""
    rt0 = np.ascontiguousarray(rt0, dtype=np.float64)
    rt1 = np.ascontiguousarray(rt1, dtype=np.float64)
    out = np.zeros(6)
    z = (None, 0, 0)
    lib().mrcal_compose_rt_full(_dp(out), 0, *z, *z, *z, *z, *z, *z,
                                _dp(rt0), 0, _dp(rt1), 0, C.c_bool(False), C.c_bool(False))
    return out


def invert_rt(rt):
    """Reference mrcal_invert_rt_full (poseutils.c)."""
    rt = np.ascontiguousarray(rt, dtype=np.float64)
    out = np.zeros(6)
    lib().mrcal_invert_rt_full(_dp(out), 0, None, 0, 0, None, 0, 0, _dp(rt), 0)
    return out
