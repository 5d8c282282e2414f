// TEST INFRASTRUCTURE ONLY. Minimal stand-in for libdogleg's <dogleg.h>, which
// is an external dependency of the reference (mrcal.c:16) that is not present
// in this image. It declares just enough for /root/reference/mrcal.c to
// compile, so that the reference's own residual/Jacobian code
// (mrcal_optimizer_callback and friends) can be built into oracle/_ref/ and
// used as the parity oracle. The solver entry points are stubs that fail: the
// reference's trust-region loop itself lives in libdogleg and is restated in
// oracle/dogleg_np.py.
#pragma once
#include <stddef.h>
#include <stdbool.h>

#ifdef __cplusplus
extern "C" {
#endif

// Only p,i,x are touched by mrcal.c (mrcal.c:4461-4463). Field order follows
// the public CHOLMOD cholmod_sparse layout so ctypes users can rely on it
typedef struct cholmod_sparse_struct
{
    size_t nrow, ncol, nzmax;
    void  *p, *i, *nz, *x, *z;
    int    stype, itype, xtype, dtype, sorted, packed;
} cholmod_sparse;

typedef struct
{
    double* p;
    double* x;
    double  norm2_x;
} dogleg_operatingPoint_t;

typedef struct
{
    dogleg_operatingPoint_t* beforeStep;
} dogleg_solverContext_t;

#define DOGLEG_DEBUG_VNLOG 1

typedef struct
{
    int    max_iterations;
    int    dogleg_debug;
    double trustregion0;
    double trustregion_decrease_factor;
    double trustregion_decrease_threshold;
    double trustregion_increase_factor;
    double trustregion_increase_threshold;
    double Jt_x_threshold;
    double update_threshold;
    double trustregion_threshold;
} dogleg_parameters2_t;

typedef void (dogleg_callback_t)(const double* p, double* x, cholmod_sparse* Jt, void* cookie);
typedef void (dogleg_callback_dense_t)(const double* p, double* x, double* J, void* cookie);

void   dogleg_getDefaultParameters(dogleg_parameters2_t* parameters);
double dogleg_optimize2(double* p, unsigned int Nstate, unsigned int Nmeas, unsigned int NJnnz,
                        dogleg_callback_t* f, void* cookie,
                        const dogleg_parameters2_t* parameters,
                        dogleg_solverContext_t** returnContext);
double dogleg_optimize_dense2(double* p, unsigned int Nstate, unsigned int Nmeas,
                              dogleg_callback_dense_t* f, void* cookie,
                              const dogleg_parameters2_t* parameters,
                              dogleg_solverContext_t** returnContext);
void   dogleg_freeContext(dogleg_solverContext_t** ctx);
void   dogleg_testGradient(unsigned int var, const double* p0,
                           unsigned int Nstate, unsigned int Nmeas, unsigned int NJnnz,
                           dogleg_callback_t* f, void* cookie);

#ifdef __cplusplus
}
#endif
