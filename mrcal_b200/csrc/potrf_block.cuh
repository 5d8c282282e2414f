// Cholesky factor and inverse factor of ONE 64x64 diagonal block, in shared memory,
// by one CTA of 256 threads. This is the serial spine of the dense factorization
// (chol.cu, chol_dataflow.cu): n/64 of these run strictly one after the other, so
// what matters here is the length of the dependent-instruction chain, not flops.
//
//   4 panels of 16 columns. WARP 0 is the chain: it factors the 16x16 diagonal
//   sub-block D_p with a row per lane in registers, in the square-root-free form
//         a_ik -= (a_ij a_kj) / d_j
//   Per pivot the chain is one shuffle, one reciprocal (hardware seed + one cubic
//   Newton step, no branches) and one FMA; the square roots are taken in the shadow of
//   that chain. Each finished column of the unit-triangular factor Lt is published to
//   shared memory at once, and every 4 columns warp 0 ARRIVES on a named barrier
//   (never waits); the threads that own the rows below (WARPS 1..7, a row per thread)
//   follow one group of 4 columns behind. The same warps then do
//       B2  the trailing update A -= X X' on the tensor pipe (DMMA 8x8x4 tiles), the three
//           tiles of the NEXT diagonal sub-block first -- warp 0 is released as soon as
//           those are done -- and row block p of inv(L)
//       B3  the products that row block p+1 of inv(L) will need
//   Named barriers (bar.sync / bar.arrive) tie the two groups together.
//
// Measured on B200 (scripts/potrf_stamps.py): fp64 ops have ~20 cycles of dependent latency,
// a pivot costs ~125 cycles on the chain, and straight-line code is fetched from a cold
// instruction cache on every launch -- hence loops over panels, not full unrolling.
//
// Stands in for the innermost part of CHOLMOD's numeric factorization as libdogleg
// drives it (call site mrcal.c:6435); not a translation of anything in the reference.
#pragma once
#include <cuda_runtime.h>

namespace mb200 {

constexpr int PB = 64;    // block
constexpr int PLD = 68;   // leading dimension in shared memory: DMMA fragment loads are conflict-free
constexpr int PSB = 16;   // panel

struct PotrfSmem
{
    double L[PB * PLD];    // in: the block (lower triangle).  out: L (lower triangle; strictly upper: scratch)
    double X[PB * PLD];    // out: inv(L), lower triangular, zeros above
    double T[PSB * PLD];   // scratch
    double R[PB];          // 1 / L_jj
};

__device__ __forceinline__ void pb_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void pb_bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }
// 1/d = x0 (1 + e2) up to e^3, e = 1 - d x0, e2 = e + e^2: x0 is the hardware seed (>= 20 bits)
__device__ __forceinline__ void pb_rcp_parts(double d, double& x0, double& e2)
{
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x0) : "d"(d));
    const double e = fma(-d, x0, 1.);
    e2 = fma(e, e, e);
}
__device__ __forceinline__ double pb_rsqrt(double d)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    const double t = d * y;
    const double e = fma(-t, y, 1.);
    const double q = e * fma(0.375, e, 0.5);
    return fma(y, q, y);                                     // y (1 + e/2 + 3e^2/8)
}
__device__ __forceinline__ void pb_dmma(double& c0, double& c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// All 256 threads of the CTA call this after storing the block to sm.L (no barrier needed in between). pivot_base: global index of the block's first
// pivot (for the not-positive-definite report: *info = 1 + index of the first bad pivot < nreal).
// Ends with a __syncthreads().
// STAMP: debugging aid, writes clock64() at the phase boundaries of warp 0 / warp 1 to stamps[]
template <bool STAMP = false>
__device__ __forceinline__ void potrf_block(PotrfSmem& sm, int* __restrict__ info, int pivot_base, int nreal, long long* stamps = nullptr)
{
#define PB_STAMP(i) do { if(STAMP && lane == 0) stamps[i] = clock64(); } while(0)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    __syncthreads();   // the caller's stores to sm.L
    PB_STAMP(warp == 0 ? 0 : warp == 1 ? 32 : 63);
    for(int e = tid; e < PB * PLD; e += 256) sm.X[e] = 0.;
    __syncthreads();

    if(warp == 0)
    {
        const int li = lane & 15;   // lanes 16..31 shadow lanes 0..15
        for(int p = 0; p < PB / PSB; p++)
        {
            const int c0 = p * PSB;
            if(p > 0) pb_bar_sync(2, 128);   // the next diagonal sub-block has all its updates
            PB_STAMP(1 + 4 * p);
            double a[PSB], dd[PSB], sq[PSB];
            int firstbad = PSB;
#pragma unroll
            for(int k = 0; k < PSB; k++) a[k] = k <= li ? sm.L[(c0 + li) * PLD + c0 + k] : 0.;
            __syncwarp();   // (the loads above may touch the strictly upper part, which the loop below overwrites)
#pragma unroll
            for(int j = 0; j < PSB; j++)
            {
                double d = __shfl_sync(0xffffffffu, a[j], j);
                // not positive definite (or NaN): remember the first failing pivot, carry on with a harmless value.
                // (integer test of the high word: positive, normal, finite -- keeps the fp64 pipe off the chain)
                const int hi = __double2hiint(d);
                const bool good = hi >= 0x03f00000 && hi < 0x7ff00000;
                firstbad = (!good && firstbad == PSB) ? j : firstbad;
                d = good ? d : 1.;
                dd[j] = d;
                const double aj = a[j];
                double x0, e2;
                pb_rcp_parts(d, x0, e2);
                if(j + 1 < PSB)
                {
                    // the entry the NEXT pivot comes from: a + p/d = (a + p x0) + (p x0) e2, one op after e2
                    const double akj = __shfl_sync(0xffffffffu, aj, j + 1);
                    const double q = -(aj * akj) * x0;
                    a[j + 1] = fma(q, e2, a[j + 1] + q);
                }
                const double r = fma(x0, e2, x0);
                // column j of Lt: published for the rows below, and read back (broadcast) by this warp too --
                // cheaper than 15 64-bit shuffles and a multiply per entry
                if(lane > j && lane < PSB) sm.L[(c0 + j) * PLD + c0 + lane] = aj * r;
                sq[j] = pb_rsqrt(d);
                if(lane == j) sm.R[c0 + j] = sq[j];
                __syncwarp();
#pragma unroll
                for(int k = j + 2; k < PSB; k++) a[k] = fma(-aj, sm.L[(c0 + j) * PLD + c0 + k], a[k]);
                if((j & 3) == 3) pb_bar_arrive(4 + (j >> 2), 256);   // columns j-3..j of Lt (and of R) are out
            }
            PB_STAMP(2 + 4 * p);
            // L_ij = a_ij / sqrt(d_j)
            if(lane < PSB)
            {
#pragma unroll
                for(int j = 0; j < PSB; j++)
                    if(j <= li) sm.L[(c0 + li) * PLD + c0 + j] = (j == li ? dd[j] : a[j]) * sq[j];
            }
            if(firstbad < PSB && lane == 0 && pivot_base + c0 + firstbad < nreal) atomicCAS(info, 0, pivot_base + c0 + firstbad + 1);
            PB_STAMP(3 + 4 * p);
        }
    }
    else
    {
        const int bt = tid - 32, bw = warp - 1;   // 224 threads, 7 warps
        for(int p = 0; p < PB / PSB; p++)
        {
            const int c0 = p * PSB;
            const int m = PB - PSB - c0;   // rows below this panel
            if(warp == 1) PB_STAMP(33 + 6 * p);
            // ---- B1: rows below D_p, X = A inv(D_p)', a row per thread, one group of 4 columns behind warp 0; warp 7
            // inverts D_p the same way (column `lane` of inv(D_p) = diag(1/L_jj) inv(Lt); lanes 16..31 shadow).
            // The roles are per WARP, and every warp of the group waits at the SAME four barrier instructions
            const bool consumer = bw * 32 < m;
            const bool inverter = !consumer && warp == 7;
            const bool mine = consumer && bt < m;
            const int row = mine ? c0 + PSB + bt : PB - 1;
            const int li = lane & 15;
            double v[PSB];   // consumer: the row being solved;  inverter: column li of inv(Lt)
#pragma unroll
            for(int k = 0; k < PSB; k++) v[k] = consumer ? sm.L[row * PLD + c0 + k] : (k == li ? 1. : 0.);
#pragma unroll
            for(int jg = 0; jg < PSB / 4; jg++)
            {
                pb_bar_sync(4 + jg, 256);   // columns 4 jg .. 4 jg + 3 of Lt and of R are in shared memory
                if(consumer)
                {
#pragma unroll
                    for(int j = 4 * jg; j < 4 * jg + 4; j++)
                    {
                        if(mine) sm.L[row * PLD + c0 + j] = v[j] * sm.R[c0 + j];   // x_j is final
#pragma unroll
                        for(int k = j + 1; k < PSB; k++) v[k] = fma(-v[j], sm.L[(c0 + j) * PLD + c0 + k], v[k]);
                    }
                }
                else if(inverter)
                {
#pragma unroll
                    for(int k = 4 * jg; k < 4 * jg + 4; k++)
#pragma unroll
                        for(int i = k + 1; i < PSB; i++) v[i] = fma(-v[k], sm.L[(c0 + k) * PLD + c0 + i], v[i]);
                }
            }
            if(inverter && lane < PSB)
            {
#pragma unroll
                for(int i = 0; i < PSB; i++) sm.X[(c0 + i) * PLD + c0 + li] = i >= li ? v[i] * sm.R[c0 + i] : 0.;
            }
            if(warp == 1) PB_STAMP(34 + 6 * p);
            pb_bar_sync(3, 224);
            if(warp == 1) PB_STAMP(35 + 6 * p);
            // ---- B2: trailing update, 8x8 tiles of the m x m lower triangle, next diagonal sub-block first
            {
                const int ns = m / 8, ntiles = ns * (ns + 1) / 2;
                int ti = 0, tj = 0;   // tile number tl <-> (ti, tj), row-major over the lower triangle
                for(int q = 0; q < bw; q++) { if(tj == ti) { ti++; tj = 0; } else tj++; }
                for(int tl = bw; tl < ntiles; tl += 7)
                {
                    const int r0 = c0 + PSB + 8 * ti, q0 = c0 + PSB + 8 * tj;
                    double acc0 = 0., acc1 = 0., acc2 = 0., acc3 = 0.;
#pragma unroll
                    for(int ks = 0; ks < PSB / 4; ks += 2)
                    {
                        pb_dmma(acc0, acc1, sm.L[(r0 + g) * PLD + c0 + ks * 4 + t], sm.L[(q0 + g) * PLD + c0 + ks * 4 + t]);
                        pb_dmma(acc2, acc3, sm.L[(r0 + g) * PLD + c0 + ks * 4 + 4 + t], sm.L[(q0 + g) * PLD + c0 + ks * 4 + 4 + t]);
                    }
                    double* c = &sm.L[(r0 + g) * PLD + q0 + 2 * t];
                    // (the strictly upper part of a diagonal tile is not ours: Lt gets published there)
                    if(ti != tj || 2 * t <= g) c[0] -= acc0 + acc2;
                    if(ti != tj || 2 * t + 1 <= g) c[1] -= acc1 + acc3;
                    if(tl < 3)
                    {
                        __threadfence_block();
                        pb_bar_arrive(2, 128);
                        if(warp == 1) PB_STAMP(36 + 6 * p);
                    }
                    for(int q = 0; q < 7; q++) { if(tj == ti) { ti++; tj = 0; } else tj++; }
                }
                // row block p of inv(L): X_pj = -inv(D_p) T_pj, j < p   (T from the previous panel's B3)
                for(int tl = (bw + 3) % 7; tl < 2 * (2 * p); tl += 7)
                {
                    const int mi = tl & 1, nj = tl >> 1;
                    double acc0 = 0., acc1 = 0.;
#pragma unroll
                    for(int ks = 0; ks < PSB / 4; ks++)
                        pb_dmma(acc0, acc1, sm.X[(c0 + 8 * mi + g) * PLD + c0 + ks * 4 + t], sm.T[(ks * 4 + t) * PLD + 8 * nj + g]);
                    double2* c = reinterpret_cast<double2*>(&sm.X[(c0 + 8 * mi + g) * PLD + 8 * nj + 2 * t]);
                    *c = make_double2(-acc0, -acc1);
                }
            }
            if(warp == 1) PB_STAMP(37 + 6 * p);
            if(p == PB / PSB - 1) break;
            pb_bar_sync(3, 224);
            // ---- B3: T_{p+1,j} = sum_{k=j..p} L_{p+1,k} X_kj, j <= p: 2 x 2(p+1) tiles
            for(int tl = bw; tl < 2 * (2 * (p + 1)); tl += 7)
            {
                const int mi = tl & 1, nj = tl >> 1;
                const int r0 = c0 + PSB + 8 * mi;
                double acc0 = 0., acc1 = 0.;
                for(int k = (nj >> 1) * PSB; k < c0 + PSB; k += 4)
                    pb_dmma(acc0, acc1, sm.L[(r0 + g) * PLD + k + t], sm.X[(k + t) * PLD + 8 * nj + g]);
                *reinterpret_cast<double2*>(&sm.T[(8 * mi + g) * PLD + 8 * nj + 2 * t]) = make_double2(acc0, acc1);
            }
        }
    }
    __syncthreads();
    PB_STAMP(warp == 0 ? 20 : warp == 1 ? 60 : 63);
#undef PB_STAMP
}

}  // namespace mb200
