// Internal: the Problem object behind mrcal_b200_problem_t.
#pragma once
#include <memory>
#include <vector>

#include "problem.h"

namespace mb200 {

// Owns a set of device allocations; frees them on destruction
struct DeviceArena
{
    std::vector<void*> ptrs;
    size_t bytes = 0;
    ~DeviceArena() { release(); }
    void release()
    {
        for(void* p : ptrs) cudaFree(p);
        ptrs.clear();
        bytes = 0;
    }
    template <typename T> bool alloc(T** out, size_t n, bool zero = false)
    {
        *out = nullptr;
        const size_t nb = (n ? n : 1) * sizeof(T);
        void* p = nullptr;
        cudaError_t e = cudaMalloc(&p, nb);
        if(e != cudaSuccess) { set_error("cudaMalloc(%zu bytes) failed: %s", nb, cudaGetErrorString(e)); return false; }
        if(zero) cudaMemset(p, 0, nb);
        ptrs.push_back(p);
        bytes += nb;
        *out = (T*)p;
        return true;
    }
};

struct SolverWorkspace;    // solver.cu
struct OutlierWorkspace;   // outliers.cu

}  // namespace mb200

// The opaque handle of the C-ABI
struct mrcal_b200_problem
{
    mb200::Layout     L;
    mb200::DevProblem dp;
    mb200::DeviceArena arena;
    cudaStream_t stream = nullptr;
    int device = 0;
    int nnz = 0;

    // seed kept on the device so reset() needs no host traffic
    double* d_seed_intr = nullptr; double* d_seed_rtcam = nullptr; double* d_seed_rtframe = nullptr;
    double* d_seed_points = nullptr; double* d_seed_warp = nullptr;
    double* d_pool_board = nullptr;       // live (outlier-marked) observation pool
    double* d_pool_board_seed = nullptr;  // as given
    double* d_pool_point = nullptr;
    double* d_scale = nullptr;            // [Nstate] pack scales
    int*    d_rowptr = nullptr;           // [Nmeas+1], filled on demand

    mb200::EvalBuffers op[2];             // two operating points (before/after step)
    int cur = 0;                          // which of op[] is the accepted state

    std::vector<int> h_board_j0, h_point_j0;
    std::vector<int> h_obs_board;         // [Nobs][3] icam_i, icam_e, iframe
    std::vector<int> h_obs_point;
    int Nobs_tri = 0;                     // triangulated observations
    int* d_tri_outlier_seed = nullptr;    // their outlier flags as given (reset() restores them)

    std::unique_ptr<mb200::SolverWorkspace, void (*)(mb200::SolverWorkspace*)> ws{nullptr, nullptr};
    std::unique_ptr<mb200::OutlierWorkspace, void (*)(mb200::OutlierWorkspace*)> ows{nullptr, nullptr};
    int launches = 0;                     // kernel launches so far (this library's kernels only)

    // multi-GPU sharding (identity when single GPU)
    bool sharded = false;
    int frame_offset = 0, Nframes_global = 0, point_offset = 0, Npoints_global = 0;
};

namespace mb200 {
bool problem_pack_seed(mrcal_b200_problem* P);                 // seed -> op[cur].p
bool problem_evaluate(mrcal_b200_problem* P, int which, bool with_jacobian, bool with_rowptr);
bool problem_unpack_to_seed_layout(mrcal_b200_problem* P, int which, double* d_intr, double* d_rtcam,
                                   double* d_rtframe, double* d_points, double* d_warp);
bool outliers_mark(mrcal_b200_problem* P, bool* found, int* Noutliers_board, int* Noutliers_tri);   // outliers.cu
bool solver_run(mrcal_b200_problem* P, const mrcal_b200_solver_parameters_t* params,
                mrcal_stats_t* stats, mrcal_b200_solve_info_t* info);
}  // namespace mb200
