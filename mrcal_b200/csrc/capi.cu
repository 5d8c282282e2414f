// The reference-named entry points of the hot path (Part 1d of
// include/mrcal_b200.h): thin drivers over the device-resident problem.
#include "problem_impl.h"

using namespace mb200;

#include <mutex>

namespace {
// The reference-named entry points are called over and over with problems of the same shape
// (mrcal-calibrate-cameras solves 5-6 times; the uncertainty tools hundreds of times). Creating a
// device problem means ~0.7 GB of cudaMalloc, pinned allocations and CUDA-graph captures: ~0.5 s.
// So the last problem is kept and re-used when the next call has the same shape; only the values
// (seed, observations) are re-uploaded. One entry, guarded by a mutex: the reference is single-threaded.
struct ProblemCache
{
    std::mutex mtx;
    mrcal_b200_problem_t* p = nullptr;
    std::vector<int> imagersizes;
    bool have_warp = false;
    double spacing = 0.;
    ~ProblemCache() { /* the CUDA context may already be gone at process exit: leak on purpose */ }
} g_cache;

mrcal_b200_problem_t* acquire_problem(const double* intrinsics, const mrcal_pose_t* rt_cam_ref, const mrcal_pose_t* rt_ref_frame,
                                      const mrcal_point3_t* points, const mrcal_calobject_warp_t* calobject_warp,
                                      int Ncam_i, int Ncam_e, int Nframes, int Npoints, int Npoints_fixed,
                                      const mrcal_observation_board_t* ob, const mrcal_observation_point_t* op,
                                      int Nob, int Nop, const mrcal_point3_t* pool_b, const mrcal_point3_t* pool_p,
                                      const mrcal_lensmodel_t* lensmodel, const int* imagersizes,
                                      mrcal_problem_selections_t sel, double spacing, int W, int H,
                                      const mrcal_observation_point_triangulated_t* otri = nullptr, int Notri = 0)
{
    if(otri == nullptr || Notri < 0) Notri = 0;
    if(Nob < 0) Nob = 0;
    if(Nop < 0) Nop = 0;
    mrcal_b200_problem_t* P = g_cache.p;
    // (problems with triangulated observations carry their rays on the device: not reused)
    bool same = P != nullptr && Notri == 0 && P->dp.Ntri == 0;
    if(same)
    {
        const Dims& d = P->L.d;
        mrcal_problem_selections_t s2 = sel;
        if(Nob <= 0) s2.do_optimize_calobject_warp = false;
        same = d.Ncam_i == Ncam_i && d.Ncam_e == Ncam_e && d.Nframes == Nframes && d.Npoints == Npoints &&
               d.Npoints_fixed == Npoints_fixed && d.Nobs_board == Nob && d.Nobs_point == Nop &&
               (Nob == 0 || (d.W == W && d.H == H)) &&
               memcmp(&P->L.lensmodel, lensmodel, sizeof(*lensmodel)) == 0 &&
               memcmp(&P->L.sel, &s2, sizeof(s2)) == 0 &&
               g_cache.have_warp == (calobject_warp != nullptr) && g_cache.spacing == spacing &&
               (int)g_cache.imagersizes.size() == 2 * Ncam_i &&
               (Ncam_i == 0 || memcmp(g_cache.imagersizes.data(), imagersizes, 2 * Ncam_i * sizeof(int)) == 0) &&
               !P->sharded;
        for(int i = 0; same && i < Nob; i++)
            same = P->h_obs_board[3 * i] == ob[i].icam.intrinsics &&
                   P->h_obs_board[3 * i + 1] == (ob[i].icam.extrinsics < 0 ? -1 : ob[i].icam.extrinsics) &&
                   P->h_obs_board[3 * i + 2] == ob[i].iframe;
        for(int i = 0; same && i < Nop; i++)
            same = P->h_obs_point[3 * i] == op[i].icam.intrinsics &&
                   P->h_obs_point[3 * i + 1] == (op[i].icam.extrinsics < 0 ? -1 : op[i].icam.extrinsics) &&
                   P->h_obs_point[3 * i + 2] == op[i].i_point;
    }
    if(same)
    {
        if(!mrcal_b200_problem_upload(P, intrinsics, rt_cam_ref, rt_ref_frame, points, calobject_warp, pool_b, pool_p))
            return nullptr;
        return P;
    }
    if(P) { mrcal_b200_problem_destroy(P); g_cache.p = nullptr; }
    P = mrcal_b200_problem_create_triangulated(intrinsics, rt_cam_ref, rt_ref_frame, points, calobject_warp,
                                               Ncam_i, Ncam_e, Nframes, Npoints, Npoints_fixed, ob, op, Nob, Nop, otri, Notri,
                                               pool_b, pool_p, lensmodel, imagersizes, sel, spacing, W, H);
    if(P == nullptr) return nullptr;
    g_cache.p = P;
    g_cache.imagersizes.assign(imagersizes, imagersizes + 2 * Ncam_i);
    g_cache.have_warp = calobject_warp != nullptr;
    g_cache.spacing = spacing;
    return P;
}

}  // namespace

namespace mb200 {
// factorization_schur.cu takes over the problem the last callback left in the cache
mrcal_b200_problem_t* capi_steal_cached_problem()
{
    std::lock_guard<std::mutex> lock(g_cache.mtx);
    mrcal_b200_problem_t* P = g_cache.p;
    g_cache.p = nullptr;
    return P;
}
}  // namespace mb200

extern "C" bool mrcal_optimizer_callback(double* b_packed, int buffer_size_b_packed,
                                         double* x, int buffer_size_x,
                                         mrcal_b200_sparse_t* Jt,
                                         const double* intrinsics, const mrcal_pose_t* rt_cam_ref,
                                         const mrcal_pose_t* rt_ref_frame, const mrcal_point3_t* points,
                                         const mrcal_calobject_warp_t* calobject_warp,
                                         int Ncameras_intrinsics, int Ncameras_extrinsics, int Nframes,
                                         int Npoints, int Npoints_fixed,
                                         const mrcal_observation_board_t* observations_board,
                                         const mrcal_observation_point_t* observations_point,
                                         int Nobservations_board, int Nobservations_point,
                                         const mrcal_observation_point_triangulated_t* observations_point_triangulated,
                                         int Nobservations_point_triangulated,
                                         const mrcal_point3_t* observations_board_pool,
                                         const mrcal_point3_t* observations_point_pool,
                                         const mrcal_lensmodel_t* lensmodel, const int* imagersizes,
                                         mrcal_problem_selections_t problem_selections,
                                         const mrcal_problem_constants_t* problem_constants,
                                         double calibration_object_spacing,
                                         int calibration_object_width_n, int calibration_object_height_n,
                                         bool verbose)
{
    (void)problem_constants; (void)verbose;
    if(b_packed == nullptr || x == nullptr)
    {
        set_error("mrcal_optimizer_callback(): b_packed and x may not be NULL");
        return false;
    }
    std::lock_guard<std::mutex> lock(g_cache.mtx);
    struct { mrcal_b200_problem_t* p; } g{acquire_problem(intrinsics, rt_cam_ref, rt_ref_frame, points, calobject_warp,
                                             Ncameras_intrinsics, Ncameras_extrinsics, Nframes, Npoints, Npoints_fixed,
                                             observations_board, observations_point, Nobservations_board, Nobservations_point,
                                             observations_board_pool, observations_point_pool, lensmodel, imagersizes,
                                             problem_selections, calibration_object_spacing,
                                             calibration_object_width_n, calibration_object_height_n,
                                             observations_point_triangulated, Nobservations_point_triangulated)};
    if(g.p == nullptr) return false;
    const int Nstate = g.p->L.Nstate, Nmeas = g.p->L.Nmeas;
    if(buffer_size_b_packed != Nstate * (int)sizeof(double))
    {
        set_error("The buffer passed to fill-in b_packed has the wrong size. Needed exactly %d bytes, but got %d bytes",
                  Nstate * (int)sizeof(double), buffer_size_b_packed);
        return false;
    }
    if(buffer_size_x != Nmeas * (int)sizeof(double))
    {
        set_error("The buffer passed to fill-in x has the wrong size. Needed exactly %d bytes, but got %d bytes",
                  Nmeas * (int)sizeof(double), buffer_size_x);
        return false;
    }
    if(Jt != nullptr && (Jt->p == nullptr || Jt->i == nullptr || Jt->x == nullptr))
    {
        set_error("mrcal_optimizer_callback(): Jt given, but its p/i/x arrays are not");
        return false;
    }
    return mrcal_b200_problem_callback(g.p, b_packed, x,
                                       Jt ? (int32_t*)Jt->p : nullptr, Jt ? (int32_t*)Jt->i : nullptr,
                                       Jt ? (double*)Jt->x : nullptr);
}

extern "C" mrcal_stats_t mrcal_optimize(double* b_packed_final, int buffer_size_b_packed_final,
                                        double* x_final, int buffer_size_x_final,
                                        double* intrinsics, mrcal_pose_t* rt_cam_ref, mrcal_pose_t* rt_ref_frame,
                                        mrcal_point3_t* points, mrcal_calobject_warp_t* calobject_warp,
                                        int Ncameras_intrinsics, int Ncameras_extrinsics, int Nframes,
                                        int Npoints, int Npoints_fixed,
                                        const mrcal_observation_board_t* observations_board,
                                        const mrcal_observation_point_t* observations_point,
                                        int Nobservations_board, int Nobservations_point,
                                        const mrcal_observation_point_triangulated_t* observations_point_triangulated,
                                        int Nobservations_point_triangulated,
                                        mrcal_point3_t* observations_board_pool,
                                        mrcal_point3_t* observations_point_pool,
                                        const mrcal_lensmodel_t* lensmodel, const int* imagersizes,
                                        mrcal_problem_selections_t problem_selections,
                                        const mrcal_problem_constants_t* problem_constants,
                                        double calibration_object_spacing,
                                        int calibration_object_width_n, int calibration_object_height_n,
                                        bool verbose, bool check_gradient)
{
    (void)problem_constants; (void)verbose;
    mrcal_stats_t bad = {};
    bad.rms_reproj_error__pixels = -1.0;
    std::lock_guard<std::mutex> lock(g_cache.mtx);
    struct { mrcal_b200_problem_t* p; } g{acquire_problem(intrinsics, rt_cam_ref, rt_ref_frame, points, calobject_warp,
                                             Ncameras_intrinsics, Ncameras_extrinsics, Nframes, Npoints, Npoints_fixed,
                                             observations_board, observations_point, Nobservations_board, Nobservations_point,
                                             observations_board_pool, observations_point_pool, lensmodel, imagersizes,
                                             problem_selections, calibration_object_spacing,
                                             calibration_object_width_n, calibration_object_height_n,
                                             observations_point_triangulated, Nobservations_point_triangulated)};
    if(g.p == nullptr) return bad;
    const int Nstate = g.p->L.Nstate, Nmeas = g.p->L.Nmeas;
    if(b_packed_final != nullptr && buffer_size_b_packed_final != Nstate * (int)sizeof(double))
    {
        set_error("The buffer passed to fill-in b_packed_final has the wrong size. Needed exactly %d bytes, but got %d bytes",
                  Nstate * (int)sizeof(double), buffer_size_b_packed_final);
        return bad;
    }
    if(x_final != nullptr && buffer_size_x_final != Nmeas * (int)sizeof(double))
    {
        set_error("The buffer passed to fill-in x_final has the wrong size. Needed exactly %d bytes, but got %d bytes",
                  Nmeas * (int)sizeof(double), buffer_size_x_final);
        return bad;
    }
    if(Nmeas <= Nstate)
        fprintf(stderr, "mrcal_b200: WARNING: problem isn't overdetermined: Nmeasurements=%d, Nstate=%d\n", Nmeas, Nstate);

    if(check_gradient)
    {
        // The reference hands every state variable to libdogleg's dogleg_testGradient() (mrcal.c:6601-6605), which
        // compares the reported Jacobian column with a forward difference and prints a vnlog
        // (test/test-gradients.py parses it). The same, with our own callback. Nothing is solved; like the
        // reference, rms comes out as sqrt(-1/Nmeasurements) = NaN
        const int nnz = mrcal_b200_problem_num_j_nonzero(g.p);
        std::vector<double> b0(Nstate), x0(Nmeas), x1(Nmeas), Jv(nnz), b1;
        std::vector<int32_t> Jp(Nmeas + 1), Ji(nnz);
        if(!mrcal_b200_problem_callback(g.p, b0.data(), x0.data(), Jp.data(), Ji.data(), Jv.data())) return bad;
        // J by columns
        std::vector<std::vector<std::pair<int, double>>> cols(Nstate);
        for(int m = 0; m < Nmeas; m++)
            for(int e = Jp[m]; e < Jp[m + 1]; e++) cols[Ji[e]].push_back({m, Jv[e]});
        const double delta = 1e-6;
        printf("# ivar imeasurement gradient_reported gradient_observed error error_relative\n");
        for(int ivar = 0; ivar < Nstate; ivar++)
        {
            b1 = b0;
            b1[ivar] += delta;
            if(!mrcal_b200_problem_reset(g.p, b1.data()) || !mrcal_b200_problem_callback(g.p, nullptr, x1.data(), nullptr, nullptr, nullptr)) return bad;
            size_t k = 0;
            for(int m = 0; m < Nmeas; m++)
            {
                double rep = 0.;
                bool have = false;
                while(k < cols[ivar].size() && cols[ivar][k].first == m) { rep += cols[ivar][k].second; have = true; k++; }
                const double obs = (x1[m] - x0[m]) / delta;
                if(!have && obs == 0.) continue;
                const double err = rep - obs, den = (fabs(rep) + fabs(obs)) / 2.;
                printf("%d %d %.6g %.6g %.6g %.6g\n", ivar, m, rep, obs, err, den > 0. ? fabs(err) / den : 0.);
            }
        }
        fflush(stdout);
        mrcal_b200_problem_reset(g.p, b0.data());
        mrcal_stats_t st = {};
        st.rms_reproj_error__pixels = sqrt(-1.0 / (double)Nmeas);
        return st;
    }
    mrcal_stats_t stats = bad;
    if(!mrcal_b200_problem_optimize(g.p, nullptr, &stats, nullptr)) return bad;
    if(!mrcal_b200_problem_download(g.p, b_packed_final, x_final, intrinsics, rt_cam_ref, rt_ref_frame, points,
                                    calobject_warp, observations_board_pool))
        return bad;
    if(observations_point_triangulated != nullptr && Nobservations_point_triangulated > 0 &&
       problem_selections.do_apply_outlier_rejection)
    {
        // the reference marks new triangulated outliers in its (nominally const) input array
        // (mrcal.c:4231-4232,4370-4371 through the cast at mrcal.c:6463): so does this
        std::vector<int> flags(Nobservations_point_triangulated);
        if(mrcal_b200_problem_triangulated_outliers(g.p, flags.data(), Nobservations_point_triangulated) < 0) return bad;
        mrcal_observation_point_triangulated_t* o = (mrcal_observation_point_triangulated_t*)observations_point_triangulated;
        for(int i = 0; i < Nobservations_point_triangulated; i++) if(flags[i]) o[i].outlier = 1;
    }
    return stats;
}
