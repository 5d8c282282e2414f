// TEMPORARY stubs, replaced as the solver lands
#include "problem_impl.h"
using namespace mb200;
extern "C" void mrcal_b200_default_solver_parameters(mrcal_b200_solver_parameters_t* p)
{
    p->max_iterations = 300; p->trustregion0 = 1e3;
    p->trustregion_decrease_factor = 0.1; p->trustregion_decrease_threshold = 0.25;
    p->trustregion_increase_factor = 2.0; p->trustregion_increase_threshold = 0.75;
    p->Jt_x_threshold = 0; p->update_threshold = 1e-7; p->trustregion_threshold = 0;
}
extern "C" bool mrcal_b200_problem_optimize(mrcal_b200_problem_t*, const mrcal_b200_solver_parameters_t*, mrcal_stats_t*, mrcal_b200_solve_info_t*)
{ set_error("solver not built yet"); return false; }
extern "C" bool mrcal_b200_nccl_get_unique_id(void*) { set_error("nccl not built yet"); return false; }
extern "C" bool mrcal_b200_nccl_comm_init(const void*, int, int, int) { set_error("nccl not built yet"); return false; }
extern "C" void mrcal_b200_nccl_comm_destroy(void) {}
extern "C" bool mrcal_b200_problem_set_sharding(mrcal_b200_problem_t*, int, int, int, int) { set_error("nccl not built yet"); return false; }
extern "C" mrcal_b200_factorization_t* mrcal_b200_factorization_create(const int32_t*, const int32_t*, const double*, int, int) { set_error("factorization not built yet"); return nullptr; }
extern "C" void mrcal_b200_factorization_destroy(mrcal_b200_factorization_t*) {}
extern "C" bool mrcal_b200_factorization_solve_xt_JtJ_bt(mrcal_b200_factorization_t*, double*, const double*, int) { return false; }
extern "C" double mrcal_b200_factorization_rcond(mrcal_b200_factorization_t*) { return -1; }
