// TEST INFRASTRUCTURE ONLY. Link-time stand-ins for the LAPACK symbols the reference's sources pull in (LAPACK is
// not in this image): dgesdd_ (poseutils.c:1440, procrustes only: never exercised, a stub) and the packed Cholesky
// pair dpptrf_/dpptrs_ that uncertainty.c:1528,1551 calls on a 6x6 system -- restated here from LAPACK's documented
// algorithm (unblocked Cholesky of a symmetric matrix in packed storage, then two triangular solves). libdogleg's entry points
// (mrcal.c:3244,6290,6435,6603,6621) are provided by the restatement in
// oracle/port/dogleg_port.c.
#include <stdio.h>

void dgesdd_(void) { fprintf(stderr, "oracle/_ref: dgesdd_() is a stub\n"); }

#include <math.h>

// A = L L' for UPLO='L', packed column-major lower: ap[i + j(2n-j-1)/2] = A(i,j), i >= j (LAPACK dpptrf). info: 0 or
// the order of the first leading minor that is not positive definite
void dpptrf_(const char* uplo, const int* n_, double* ap, int* info)
{
    const int n = *n_;
    *info = 0;
    if(*uplo != 'L' && *uplo != 'l') { *info = -1; return; }
#define AP(i, j) ap[(i) + (j) * (2 * n - (j) - 1) / 2]
    for(int j = 0; j < n; j++)
    {
        double d = AP(j, j);
        for(int k = 0; k < j; k++) d -= AP(j, k) * AP(j, k);
        if(!(d > 0.)) { *info = j + 1; return; }
        d = sqrt(d);
        AP(j, j) = d;
        for(int i = j + 1; i < n; i++)
        {
            double v = AP(i, j);
            for(int k = 0; k < j; k++) v -= AP(i, k) * AP(j, k);
            AP(i, j) = v / d;
        }
    }
}
// solves A X = B with the factor from dpptrf_ (LAPACK dpptrs); B column-major, leading dimension ldb
void dpptrs_(const char* uplo, const int* n_, const int* nrhs, const double* ap, double* b, const int* ldb, int* info)
{
    const int n = *n_;
    *info = 0;
    if(*uplo != 'L' && *uplo != 'l') { *info = -1; return; }
    for(int r = 0; r < *nrhs; r++)
    {
        double* x = b + (size_t)r * *ldb;
        for(int i = 0; i < n; i++) { double v = x[i]; for(int k = 0; k < i; k++) v -= AP(i, k) * x[k]; x[i] = v / AP(i, i); }
        for(int i = n - 1; i >= 0; i--) { double v = x[i]; for(int k = i + 1; k < n; k++) v -= AP(k, i) * x[k]; x[i] = v / AP(i, i); }
    }
#undef AP
}
