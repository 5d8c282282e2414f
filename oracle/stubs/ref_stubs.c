// TEST INFRASTRUCTURE ONLY. Link-time stand-ins for the external symbols the
// reference's sources pull in but that the oracle never exercises:
// libdogleg's solver (mrcal.c:3244,6290,6435,6603,6621) and LAPACK's dgesdd_
// (poseutils.c:1440, procrustes only). Every solver stub reports failure.
#include <stdio.h>
#include <string.h>
#include "dogleg.h"

void dogleg_getDefaultParameters(dogleg_parameters2_t* p)
{
    memset(p, 0, sizeof(*p));
}
double dogleg_optimize2(double* p, unsigned int Nstate, unsigned int Nmeas, unsigned int NJnnz,
                        dogleg_callback_t* f, void* cookie,
                        const dogleg_parameters2_t* parameters,
                        dogleg_solverContext_t** returnContext)
{
    fprintf(stderr, "oracle/_ref: dogleg_optimize2() is a stub: libdogleg is not available\n");
    if(returnContext) *returnContext = NULL;
    return -1.0;
}
double dogleg_optimize_dense2(double* p, unsigned int Nstate, unsigned int Nmeas,
                              dogleg_callback_dense_t* f, void* cookie,
                              const dogleg_parameters2_t* parameters,
                              dogleg_solverContext_t** returnContext)
{
    fprintf(stderr, "oracle/_ref: dogleg_optimize_dense2() is a stub: libdogleg is not available\n");
    if(returnContext) *returnContext = NULL;
    return -1.0;
}
void dogleg_freeContext(dogleg_solverContext_t** ctx) { if(ctx) *ctx = NULL; }
void dogleg_testGradient(unsigned int var, const double* p0,
                         unsigned int Nstate, unsigned int Nmeas, unsigned int NJnnz,
                         dogleg_callback_t* f, void* cookie)
{
    fprintf(stderr, "oracle/_ref: dogleg_testGradient() is a stub\n");
}
void dgesdd_(void) { fprintf(stderr, "oracle/_ref: dgesdd_() is a stub\n"); }
