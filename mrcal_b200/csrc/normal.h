// Internal: normal-equation assembly + Schur elimination (normal.cu)
#pragma once
#include "problem.h"

namespace mb200 {

// All device pointers. "work item" = one board observation or one point
// observation; "group" = one eliminated block (a frame: 6 unknowns, or a
// non-fixed point: 3 unknowns) with the work items that see it.
struct NormalBuffers
{
    double* S;        // [ldS][ldS] row-major, lower: reduced normal matrix (then its Cholesky factor)
    int     ldS;      // n_r padded to a multiple of 64
    int     n_r;      // number of shared (non-eliminated) unknowns
    double* gs;       // [ldS] reduced gradient g' (reduced numbering)
    double* g_full;   // [Nstate] J'x, state numbering
    int*    info;     // 0, or a code for the first non-PD block found
    int     e0, e1;   // eliminated state range [e0,e1)

    int     cap;      // capacity (columns) of one work item's B rows
    int*    wi_nsh;   // [Nwi] shared columns the item touches
    int*    wi_cols;  // [Nwi][cap] their reduced indices, increasing
    double* wi_B;     // [Nwi][6][cap] (eliminated x shared) block of the item's Gram matrix
    double* wi_D;     // [Nwi][36]
    double* wi_gf;    // [Nwi][6]

    int     Ngroups, Nframe_groups;
    int*    grp_ptr;    // [Ngroups+1]
    int*    grp_items;  // work item ids
    double* grp_Dinv;   // [Ngroups][36]
    double* grp_gf;     // [Ngroups][6]

    __host__ __device__ int reduced_index(int c) const { return c < e0 ? c : c - (e1 - e0); }
    __host__ __device__ int state_index(int r) const { return r < e0 ? r : r + (e1 - e0); }
};

// S, g', g_full at the operating point `op` (needs its x and Jacobian). lambda: diagonal loading
bool normal_assemble(const DevProblem& dp, const NormalBuffers& N, const EvalBuffers& op, const int* d_rowptr,
                     double lambda, cudaStream_t s, int* nlaunch);
// eliminated part of the step from the reduced solution ds (reduced numbering)
bool normal_backsubstitute(const NormalBuffers& N, const double* ds, double* step_full, int e0, cudaStream_t s, int* nlaunch);

}  // namespace mb200
