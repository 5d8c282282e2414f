// The reference-named entry points of the hot path (Part 1d of
// include/mrcal_b200.h): thin drivers over the device-resident problem.
#include "problem_impl.h"

using namespace mb200;

namespace {
struct ProblemGuard
{
    mrcal_b200_problem_t* p;
    ~ProblemGuard() { mrcal_b200_problem_destroy(p); }
};

bool have_triangulated(const mrcal_observation_point_triangulated_t* o, int N) { return o != nullptr && N > 0; }
}  // namespace

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
    if(have_triangulated(observations_point_triangulated, Nobservations_point_triangulated))
    {
        set_error("ERROR: triangulated points are not implemented in the CUDA path yet");
        return false;
    }
    if(b_packed == nullptr || x == nullptr)
    {
        set_error("mrcal_optimizer_callback(): b_packed and x may not be NULL");
        return false;
    }
    ProblemGuard g{mrcal_b200_problem_create(intrinsics, rt_cam_ref, rt_ref_frame, points, calobject_warp,
                                             Ncameras_intrinsics, Ncameras_extrinsics, Nframes, Npoints, Npoints_fixed,
                                             observations_board, observations_point, Nobservations_board, Nobservations_point,
                                             observations_board_pool, observations_point_pool, lensmodel, imagersizes,
                                             problem_selections, calibration_object_spacing,
                                             calibration_object_width_n, calibration_object_height_n)};
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
    if(check_gradient)
    {
        set_error("mrcal_optimize(check_gradient=true) is a libdogleg debugging facility and is not provided by the CUDA path");
        return bad;
    }
    if(have_triangulated(observations_point_triangulated, Nobservations_point_triangulated))
    {
        set_error("ERROR: triangulated points are not implemented in the CUDA path yet");
        return bad;
    }
    ProblemGuard g{mrcal_b200_problem_create(intrinsics, rt_cam_ref, rt_ref_frame, points, calobject_warp,
                                             Ncameras_intrinsics, Ncameras_extrinsics, Nframes, Npoints, Npoints_fixed,
                                             observations_board, observations_point, Nobservations_board, Nobservations_point,
                                             observations_board_pool, observations_point_pool, lensmodel, imagersizes,
                                             problem_selections, calibration_object_spacing,
                                             calibration_object_width_n, calibration_object_height_n)};
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

    mrcal_stats_t stats = bad;
    if(!mrcal_b200_problem_optimize(g.p, nullptr, &stats, nullptr)) return bad;
    if(!mrcal_b200_problem_download(g.p, b_packed_final, x_final, intrinsics, rt_cam_ref, rt_ref_frame, points,
                                    calobject_warp, observations_board_pool))
        return bad;
    return stats;
}
