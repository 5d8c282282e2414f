// Internal: normal-equation assembly + Schur elimination (normal.cu)
#pragma once
#include "problem.h"

namespace mb200 {

// All device pointers. "work item" = one board observation or one point
// observation; "group" = one eliminated block (a frame: 6 unknowns, or a
// non-fixed point: 3 unknowns) with the work items that see it.
constexpr int kYld = 68;              // row stride of the panels of Ypan (doubles)
constexpr int kYpanel = 6 * kYld;     // one (group, block) panel

struct NormalBuffers
{
    double* S;        // [ldS][ldS] row-major, lower: reduced normal matrix over the ACTIVE shared unknowns
                      // (compact numbering), then its Cholesky factor
    int     ldS;      // n_c padded to a multiple of 64 (changes per assembly)
    int     ldS_max;  // n_r padded: the allocation
    int     n_r;      // number of shared (non-eliminated) unknowns
    int     n_c;      // ... of which touched by some observation: the coupled ("active") ones
    int     max_ntot; // widest work item (local columns) at the operating point of the last normal_prepare()
    double* gs;       // [ldS_max] reduced gradient g' (compact numbering)
    double* gsh;      // [ldS_max] staging for the shared part of J'x (reduced numbering); follows gs in memory
    int*    active;   // [n_r] flag
    int*    cidx;     // [n_r] reduced -> compact index, -1 if inactive
    int*    cinv;     // [ldS_max] compact -> reduced index
    int*    stat;     // device: [0] n_c, [1] widest item (local columns)
    int*    h_stat;   // pinned host mirror of stat
    bool    splined;
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
    int     schur_split;   // CTAs per group in the Schur kernel (<= most items in any group)
    int*    grp_ptr;    // [Ngroups+1]
    int*    grp_items;  // work item ids
    double* grp_Dinv;   // [Ngroups][36]
    double* grp_gf;     // [Ngroups][6]

    // ---- the atomics-free assembly (normal_det.cu): every entry of S is summed by ONE thread in a fixed order
    bool    det_available;   // the buffers below exist (MRCAL_B200_ATOMIC_ASSEMBLY=1 turns the path off)
    bool    det;          // in use for this assembly
    bool    selfchecked;  // the one-time cross-check against the atomic path has run
    bool    fused;        // the board rows are evaluated by fused_eval.cu: blocks straight from the projection, no J
    double* norm_part;    // [Nobs_board] per-observation |x|^2
    double* qf_part;      // [Nobs_board] per-observation g'(J'J)g
    cudaStream_t s_side[2];               // forked streams of the Gram kernel's size classes (owned by the workspace)
    cudaEvent_t  ev_fork, ev_join[2];
    int     capA;         // most local columns (incl. the two gradient rows) an item may have in the A pool
    double* wi_A;         // pool of per-item Gram blocks over the shared columns, [lda][lda] row-major, lower triangle
    long long* wi_Aoff;   // [Nwi] offset of each item's block in the pool (doubles)
    long long  A_pool;    // pool size (doubles)
    int*    wi_lda;       // [Nwi] nsh + 2 rounded up to even; rows nsh, nsh+1: -J'x over the item's shared columns
    unsigned short* wi_ccol;    // [Nwi][capA] compact index of each local column (increasing); then n_c, n_c+1
    unsigned char*  wi_segoff;  // [Nwi][nblk_max+1] first local column of each 64-column block of S
    int     nblk_max;     // ldS_max / 64
    double* Ypan;         // [Ngroups][nblk_max][6][kYld] inv(L_D) B of each group, dense per 64-column block (present blocks
                          // only). Rows padded to kYld: a panel goes to shared memory as ONE bulk copy and lands with a
                          // row stride whose DMMA fragment loads are free of bank conflicts
    unsigned* grp_present;   // [nblk_max][gwords] bit g: group g has columns in this block
    unsigned* wi_present;    // [nblk_max][wwords]
    unsigned* grp_blkmask;   // [Ngroups][bwords] the blocks a group is present in
    int     gwords, wwords, bwords;
    double* grp_Linv;     // [Ngroups][36] inverse of the Cholesky factor of D (lower)
    double* grp_h;        // [Ngroups][6]  inv(L_D) gf
    double* part_scratch; // partial tiles of the tiles whose contributors are split over several CTAs
    int*    part_arrive;  // their arrival counters
    double* S_packed;     // sharded solves: the lower-triangle tiles of S, each 64x64 contiguous: what is all-reduced

    __host__ __device__ int reduced_index(int c) const { return c < e0 ? c : c - (e1 - e0); }
    __host__ __device__ int state_index(int r) const { return r < e0 ? r : r + (e1 - e0); }
};

// The same in pieces, for a caller that wants to pick its own moment to wait for the device:
// prepare (device work + an async copy of the sizes) -> [synchronise the stream] -> adopt_sizes -> finish
bool normal_clear_marks(NormalBuffers& N, cudaStream_t s);
bool normal_prepare(const DevProblem& dp, NormalBuffers& N, const EvalBuffers& op, cudaStream_t s, int* nlaunch, bool boards_done,
                    double* rider = nullptr, bool* rider_summed = nullptr);
bool normal_adopt_sizes(NormalBuffers& N);
bool normal_finish(const DevProblem& dp, NormalBuffers& N, const EvalBuffers& op, const int* d_rowptr,
                   double lambda, cudaStream_t s, int* nlaunch, bool boards_done);
// fused_eval.cu
bool launch_fused_boards(const DevProblem& dp, NormalBuffers& N, const EvalBuffers& out, double* norm_part, cudaStream_t s, int* nlaunch);
bool launch_quadform_boards(const DevProblem& dp, const NormalBuffers& N, const double* g_full, double* part, double* out, cudaStream_t s, int* nlaunch);
bool normal_selfcheck(const DevProblem& dp, NormalBuffers& N, const EvalBuffers& op, const int* d_rowptr,
                      double lambda, cudaStream_t s, int* nlaunch);
// S, g', g_full at the operating point `op` (needs its x and Jacobian). lambda: diagonal loading.
// Updates N.n_c / N.ldS (one small device->host read)
bool normal_assemble(const DevProblem& dp, NormalBuffers& N, const EvalBuffers& op, const int* d_rowptr,
                     double lambda, cudaStream_t s, int* nlaunch);
// rhs <- y, L y = -g': read from the augmented row of the factor (valid after chol_factor)
bool normal_extract_y(const NormalBuffers& N, double* rhs, cudaStream_t s, int* nlaunch);
// rhs <- -g' (compact numbering, ldS entries)
bool normal_rhs(const NormalBuffers& N, double* rhs, cudaStream_t s, int* nlaunch);
// full-length Gauss-Newton step from the compact solution: active shared unknowns, inactive shared
// unknowns (from their regularization blocks), eliminated unknowns (back-substitution).
// ds_r: scratch, n_r doubles
bool normal_expand_step(const DevProblem& dp, const NormalBuffers& N, const EvalBuffers& op, double lambda,
                        const double* sol, double* ds_r, double* step_full, cudaStream_t s, int* nlaunch);

// normal_det.cu: the stages after the per-item Gram matrices, without atomics
bool normal_det_finish(const DevProblem& dp, NormalBuffers& N, const EvalBuffers& op, const int* d_rowptr,
                       double lambda, cudaStream_t s, int* nlaunch);
bool normal_det_item_prepare(const DevProblem& dp, NormalBuffers& N, cudaStream_t s, int* nlaunch);   // after the compaction
bool normal_det_item_offsets(const DevProblem& dp, NormalBuffers& N, cudaStream_t s, int* nlaunch);   // before it
size_t normal_det_part_scratch_doubles(int nblk_max);
int normal_det_part_arrive_ints(int nblk_max);
size_t normal_det_packed_doubles(int nblk_max);
bool normal_det_rhs(const NormalBuffers& N, cudaStream_t s, int* nlaunch);   // after the cross-rank reduction of S
bool normal_det_backsub(const NormalBuffers& N, const double* sol_compact, double* step_full, cudaStream_t s, int* nlaunch);

}  // namespace mb200
