// Internal: the solver workspace (solver.cu) for the pieces that build on it (factorization_schur.cu)
#pragma once
#include <vector>

#include "chol.h"
#include "normal.h"
#include "problem_impl.h"

namespace mb200 {

struct SolverWorkspace
{
    DeviceArena arena;
    NormalBuffers N{};
    double* invL = nullptr;
    double* rhs = nullptr;        // [ldS] compact solution
    double* ds_r = nullptr;       // [n_r] the same in reduced numbering
    double* step_gn = nullptr;    // [Nstate]
    double* step = nullptr;       // [Nstate]
    double* scal = nullptr;       // [64] device scalars
    double* h_scal = nullptr;     // pinned mirror
    int*    h_info = nullptr;     // pinned
    int*    ictl = nullptr;       // [8] device control flags of the step logic
    int*    h_ictl = nullptr;     // pinned
    CholScratch chol;             // this workspace's own flags of the persistent factorization kernels
    std::vector<cudaEvent_t> ev;
    ~SolverWorkspace()
    {
        chol_forget_graphs(N.S);
        for(int k = 0; k < 2; k++)
        {
            if(N.s_side[k]) cudaStreamDestroy(N.s_side[k]);
            if(N.ev_join[k]) cudaEventDestroy(N.ev_join[k]);
        }
        if(N.ev_fork) cudaEventDestroy(N.ev_fork);
        if(h_scal) cudaFreeHost(h_scal);
        if(h_info) cudaFreeHost(h_info);
        chol_scratch_destroy(&chol);
        for(auto e : ev) cudaEventDestroy(e);
    }
};

bool solver_build_workspace(mrcal_b200_problem* P);   // solver.cu

}  // namespace mb200

// The factorization object of the C-ABI: either a dense Cholesky of JtJ for an arbitrary CSR J (factorization.cu),
// or the structured factor of a calibration problem (factorization_schur.cu), which owns that problem
struct mrcal_b200_factorization
{
    mb200::DeviceArena arena;
    cudaStream_t stream = nullptr;
    int n = 0, npad = 0;
    double* H = nullptr;      // npad x npad, lower: L after factorization
    double* invL = nullptr;
    int*    info = nullptr;
    double* minmax = nullptr;
    mb200::CholScratch chol;  // this object's own flags of the persistent kernels
    // structured variant
    mrcal_b200_problem* P = nullptr;   // owned
    double* ia_L = nullptr;            // [n_r][3] Cholesky factor (l11, l21, l22) of each inactive unknown's regularization block
    int*    ia_first = nullptr;        // [n_r] is this the first unknown of its block (1), the second (2), alone (3), or active (0)
    int     Nelim = 0;
};

namespace mb200 {
bool schur_factorization_solve(mrcal_b200_factorization* F, double* out, const double* bt, int Nrhs, int sys);
double schur_factorization_rcond(mrcal_b200_factorization* F);
void schur_factorization_release(mrcal_b200_factorization* F);
}  // namespace mb200
