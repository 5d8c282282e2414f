// TEST INFRASTRUCTURE ONLY -- CPU restatement of the solver half of the hot path.
//
// The reference's trust-region loop is not in the reference tree: mrcal_optimize()
// calls dogleg_optimize2() (mrcal.c:6435) from libdogleg (github.com/dkogan/libdogleg,
// "at least version 0.15.3", doc/install.org:63), which factors JtJ with CHOLMOD
// (simplicial, supernodal=0; mirrored at mrcal-pywrap.c:179-183). Neither library
// is in this image. This file restates libdogleg's published algorithm (Powell's
// dogleg, doc/formulation.org:293-294) behind libdogleg's own interface (the
// declarations mrcal.c needs: oracle/stubs/dogleg.h), so that the reference's OWN
// mrcal_optimize() -- its pack/unpack, its markOutliers(), its outer re-solve loop,
// its statistics (mrcal.c:6179-6624) -- runs unmodified on top of it.
//
// PARITY: the loop below is a restatement from the library's documented behaviour,
// not a copy of its source (which is not available here). What it pins: everything
// in mrcal.c around the solver. What stays unpinned: libdogleg's own constants
// (trustregion0 1e3, 0.1 @ rho<0.25, 2 @ rho>0.75 on edge steps, lambda 1e-10 x10,
// thresholds 1e-8, 100 iterations; mrcal overrides four of them, mrcal.c:6296-6299).
//
// The factorization is a plain CPU sparse Cholesky written for this file:
// exact minimum-degree ordering (computed once per solve, as libdogleg calls
// cholmod_analyze once), elimination tree + up-looking simplicial LL' redone
// whenever the sparsity pattern of J moves (CHOLMOD's simplicial path tolerates
// that too, which is why the reference turns supernodal off).
//
// Nothing under mrcal_b200/ links or loads this.
#define _GNU_SOURCE
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "dogleg.h"

#define SAY(fmt, ...) fprintf(stderr, "dogleg_port: " fmt "\n", ##__VA_ARGS__)

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

////////////////////////////////////////////////////////////////////////////////
// Sparse Cholesky of JtJ + lambda I
////////////////////////////////////////////////////////////////////////////////
typedef struct
{
    int     n;
    // pattern of J the symbolic data below belongs to
    int     Nmeas, nnzJ;
    int*    Jp_saved;   // [Nmeas+1]
    int*    Ji_saved;   // [nnzJ]
    // J by columns (transpose of the row-wise storage the callback fills)
    int*    Cp;         // [n+1]
    int*    Crow;       // [nnzJ] measurement index
    int*    Cpos;       // [nnzJ] position of the entry in the row-wise arrays
    // fill-reducing ordering: perm[k] = original index of the k-th pivot; computed once
    int     have_perm;
    int*    perm;
    int*    pinv;
    // upper triangle of P (JtJ) P' by columns: row index < = column index
    int*    Up;
    int*    Ui;
    double* Ux;
    int*    Usrc_col;   // unused for numeric; kept for clarity
    // factor L by columns (diagonal first)
    int*    parent;
    int*    Lp;
    int*    Li;
    double* Lx;
    int*    Lnz;        // fill pointer per column during the numeric phase
    // work
    double* w;
    int*    stack;
    int*    flag;
    int     Lnnz;
} chol_t;

static void chol_free_symbolic(chol_t* c)
{
    free(c->Jp_saved); free(c->Ji_saved); free(c->Cp); free(c->Crow); free(c->Cpos);
    free(c->Up); free(c->Ui); free(c->Ux); free(c->parent); free(c->Lp); free(c->Li); free(c->Lx); free(c->Lnz);
    c->Jp_saved = c->Ji_saved = c->Cp = c->Crow = c->Cpos = c->Up = c->Ui = c->parent = c->Lp = c->Li = c->Lnz = NULL;
    c->Ux = c->Lx = NULL;
}
static void chol_free(chol_t* c)
{
    chol_free_symbolic(c);
    free(c->perm); free(c->pinv); free(c->w); free(c->stack); free(c->flag);
    memset(c, 0, sizeof(*c));
}

// Exact minimum degree on the elimination graph, adjacency kept as bitsets.
// adj: n x W words, symmetric, no diagonal. Ties: lowest index.
static void minimum_degree(int n, uint64_t* adj, int W, int* perm)
{
    int*  deg  = (int*)malloc((size_t)n * sizeof(int));
    char* done = (char*)calloc((size_t)n, 1);
    int*  nb   = (int*)malloc((size_t)n * sizeof(int));
    for(int i = 0; i < n; i++)
    {
        int d = 0;
        const uint64_t* a = adj + (size_t)i * W;
        for(int k = 0; k < W; k++) d += __builtin_popcountll(a[k]);
        deg[i] = d;
    }
    for(int step = 0; step < n; step++)
    {
        int p = -1, best = n + 1;
        for(int i = 0; i < n; i++)
            if(!done[i] && deg[i] < best) { best = deg[i]; p = i; }
        perm[step] = p;
        done[p] = 1;
        uint64_t* ap = adj + (size_t)p * W;
        int nnb = 0;
        for(int k = 0; k < W; k++)
        {
            uint64_t m = ap[k];
            while(m) { const int b = __builtin_ctzll(m); m &= m - 1; nb[nnb++] = 64 * k + b; }
        }
        // the neighbours of the pivot become a clique; the pivot leaves the graph
        for(int t = 0; t < nnb; t++)
        {
            const int u = nb[t];
            uint64_t* au = adj + (size_t)u * W;
            int d = 0;
            for(int k = 0; k < W; k++) { au[k] |= ap[k]; }
            au[u >> 6] &= ~(1ull << (u & 63));
            au[p >> 6] &= ~(1ull << (p & 63));
            for(int k = 0; k < W; k++) d += __builtin_popcountll(au[k]);
            deg[u] = d;
        }
    }
    free(deg); free(done); free(nb);
}

// (Re)build everything that depends on the sparsity pattern of J
static int chol_symbolic(chol_t* c, const cholmod_sparse* Jt)
{
    const int n = (int)Jt->nrow, Nmeas = (int)Jt->ncol;
    const int* Jp = (const int*)Jt->p;
    const int* Ji = (const int*)Jt->i;
    const int nnzJ = Jp[Nmeas];
    chol_free_symbolic(c);
    c->n = n; c->Nmeas = Nmeas; c->nnzJ = nnzJ;
    c->Jp_saved = (int*)malloc((size_t)(Nmeas + 1) * sizeof(int));
    c->Ji_saved = (int*)malloc((size_t)(nnzJ > 0 ? nnzJ : 1) * sizeof(int));
    memcpy(c->Jp_saved, Jp, (size_t)(Nmeas + 1) * sizeof(int));
    memcpy(c->Ji_saved, Ji, (size_t)nnzJ * sizeof(int));
    if(!c->w)
    {
        c->w     = (double*)calloc((size_t)n, sizeof(double));
        c->stack = (int*)malloc((size_t)n * sizeof(int));
        c->flag  = (int*)malloc((size_t)n * sizeof(int));
    }

    // J by columns
    c->Cp = (int*)calloc((size_t)n + 1, sizeof(int));
    c->Crow = (int*)malloc((size_t)(nnzJ > 0 ? nnzJ : 1) * sizeof(int));
    c->Cpos = (int*)malloc((size_t)(nnzJ > 0 ? nnzJ : 1) * sizeof(int));
    for(int e = 0; e < nnzJ; e++) c->Cp[Ji[e] + 1]++;
    for(int j = 0; j < n; j++) c->Cp[j + 1] += c->Cp[j];
    {
        int* fill = (int*)malloc((size_t)n * sizeof(int));
        memcpy(fill, c->Cp, (size_t)n * sizeof(int));
        for(int m = 0; m < Nmeas; m++)
            for(int e = Jp[m]; e < Jp[m + 1]; e++)
            {
                const int q = fill[Ji[e]]++;
                c->Crow[q] = m;
                c->Cpos[q] = e;
            }
        free(fill);
    }

    // pattern of JtJ as bitsets (needed for the ordering the first time; reused to build the
    // permuted upper triangle)
    const int W = (n + 63) / 64;
    uint64_t* adj = (uint64_t*)calloc((size_t)n * W, sizeof(uint64_t));
    if(!adj) { SAY("out of memory for the %d x %d adjacency bitset", n, n); return 0; }
    for(int j = 0; j < n; j++)
    {
        uint64_t* aj = adj + (size_t)j * W;
        for(int q = c->Cp[j]; q < c->Cp[j + 1]; q++)
        {
            const int m = c->Crow[q];
            for(int e = Jp[m]; e < Jp[m + 1]; e++) aj[Ji[e] >> 6] |= 1ull << (Ji[e] & 63);
        }
        aj[j >> 6] &= ~(1ull << (j & 63));
    }
    if(!c->have_perm)
    {
        c->perm = (int*)malloc((size_t)n * sizeof(int));
        c->pinv = (int*)malloc((size_t)n * sizeof(int));
        uint64_t* work = (uint64_t*)malloc((size_t)n * W * sizeof(uint64_t));
        memcpy(work, adj, (size_t)n * W * sizeof(uint64_t));
        minimum_degree(n, work, W, c->perm);
        free(work);
        for(int k = 0; k < n; k++) c->pinv[c->perm[k]] = k;
        c->have_perm = 1;
    }

    // upper triangle (incl. diagonal) of the permuted matrix, by columns, rows sorted
    c->Up = (int*)calloc((size_t)n + 1, sizeof(int));
    for(int j = 0; j < n; j++)
    {
        const uint64_t* aj = adj + (size_t)j * W;
        const int pj = c->pinv[j];
        c->Up[pj + 1]++;   // diagonal
        for(int k = 0; k < W; k++)
        {
            uint64_t m = aj[k];
            while(m)
            {
                const int i = 64 * k + __builtin_ctzll(m);
                m &= m - 1;
                if(c->pinv[i] < pj) c->Up[pj + 1]++;
            }
        }
    }
    for(int j = 0; j < n; j++) c->Up[j + 1] += c->Up[j];
    const int nnzU = c->Up[n];
    c->Ui = (int*)malloc((size_t)nnzU * sizeof(int));
    c->Ux = (double*)malloc((size_t)nnzU * sizeof(double));
    {
        // rows of column pj, sorted increasing: mark in a byte map over permuted indices
        char* mark = (char*)calloc((size_t)n, 1);
        for(int j = 0; j < n; j++)
        {
            const uint64_t* aj = adj + (size_t)j * W;
            const int pj = c->pinv[j];
            int lo = pj, cnt = 0;
            mark[pj] = 1;
            for(int k = 0; k < W; k++)
            {
                uint64_t m = aj[k];
                while(m)
                {
                    const int i = 64 * k + __builtin_ctzll(m);
                    m &= m - 1;
                    const int pi = c->pinv[i];
                    if(pi < pj) { mark[pi] = 1; if(pi < lo) lo = pi; }
                }
            }
            int q = c->Up[pj];
            for(int r = lo; r <= pj; r++)
                if(mark[r]) { c->Ui[q++] = r; mark[r] = 0; cnt++; }
            (void)cnt;
        }
        free(mark);
    }
    free(adj);

    // elimination tree of the permuted matrix (Liu's algorithm with path compression)
    c->parent = (int*)malloc((size_t)n * sizeof(int));
    {
        int* anc = (int*)malloc((size_t)n * sizeof(int));
        for(int k = 0; k < n; k++)
        {
            c->parent[k] = -1;
            anc[k] = -1;
            for(int q = c->Up[k]; q < c->Up[k + 1]; q++)
            {
                int i = c->Ui[q];
                while(i != -1 && i < k)
                {
                    const int inext = anc[i];
                    anc[i] = k;
                    if(inext == -1) c->parent[i] = k;
                    i = inext;
                }
            }
        }
        free(anc);
    }
    // column counts of L: row k of L is the reach of column k of U in the tree
    c->Lp = (int*)calloc((size_t)n + 1, sizeof(int));
    for(int k = 0; k < n; k++) c->flag[k] = -1;
    for(int k = 0; k < n; k++)
    {
        c->flag[k] = k;
        c->Lp[k + 1]++;   // diagonal
        for(int q = c->Up[k]; q < c->Up[k + 1]; q++)
            for(int i = c->Ui[q]; i != -1 && i < k && c->flag[i] != k; i = c->parent[i])
            {
                c->flag[i] = k;
                c->Lp[i + 1]++;
            }
    }
    for(int k = 0; k < n; k++) c->Lp[k + 1] += c->Lp[k];
    c->Lnnz = c->Lp[n];
    c->Li = (int*)malloc((size_t)c->Lnnz * sizeof(int));
    c->Lx = (double*)malloc((size_t)c->Lnnz * sizeof(double));
    c->Lnz = (int*)malloc((size_t)n * sizeof(int));
    if(!c->Li || !c->Lx) { SAY("out of memory for L (%d entries)", c->Lnnz); return 0; }
    return 1;
}

static int chol_pattern_matches(const chol_t* c, const cholmod_sparse* Jt)
{
    if(!c->Jp_saved || c->n != (int)Jt->nrow || c->Nmeas != (int)Jt->ncol) return 0;
    const int* Jp = (const int*)Jt->p;
    if(Jp[c->Nmeas] != c->nnzJ) return 0;
    return memcmp(c->Jp_saved, Jp, (size_t)(c->Nmeas + 1) * sizeof(int)) == 0 &&
           memcmp(c->Ji_saved, Jt->i, (size_t)c->nnzJ * sizeof(int)) == 0;
}

// Numeric: values of the permuted upper triangle of JtJ + lambda I, then up-looking LL'.
// Returns 1, or 0 if a pivot is not positive
static int chol_numeric(chol_t* c, const cholmod_sparse* Jt, double lambda)
{
    const int n = c->n;
    const int* Jp = (const int*)Jt->p;
    const int* Ji = (const int*)Jt->i;
    const double* Jx = (const double*)Jt->x;
    double* w = c->w;
    // column pj of U = entries (pinv[i], pj) of JtJ with pinv[i] <= pj
    for(int j = 0; j < n; j++)
    {
        const int pj = c->pinv[j];
        for(int q = c->Cp[j]; q < c->Cp[j + 1]; q++)
        {
            const int m = c->Crow[q];
            const double vj = Jx[c->Cpos[q]];
            if(vj == 0.) continue;
            for(int e = Jp[m]; e < Jp[m + 1]; e++) w[c->pinv[Ji[e]]] += vj * Jx[e];
        }
        for(int q = c->Up[pj]; q < c->Up[pj + 1]; q++) c->Ux[q] = w[c->Ui[q]];
        // clear what was touched
        for(int q = c->Cp[j]; q < c->Cp[j + 1]; q++)
        {
            const int m = c->Crow[q];
            for(int e = Jp[m]; e < Jp[m + 1]; e++) w[c->pinv[Ji[e]]] = 0.;
        }
        c->Ux[c->Up[pj + 1] - 1] += lambda;   // the diagonal is the last entry of the column
    }

    for(int k = 0; k < n; k++) { c->Lnz[k] = c->Lp[k]; c->flag[k] = -1; }
    int* s = c->stack;
    for(int k = 0; k < n; k++)
    {
        // nonzero pattern of row k of L: reach of column k of U, in topological order on the stack top..n
        int top = n;
        c->flag[k] = k;
        for(int q = c->Up[k]; q < c->Up[k + 1]; q++)
        {
            int i = c->Ui[q];
            w[i] = c->Ux[q];
            int len = 0;
            for(; i < k && c->flag[i] != k; i = c->parent[i]) { s[len++] = i; c->flag[i] = k; }
            while(len > 0) s[--top] = s[--len];
        }
        double d = w[k];
        w[k] = 0.;
        for(; top < n; top++)
        {
            const int i = s[top];
            const double lki = w[i] / c->Lx[c->Lp[i]];
            w[i] = 0.;
            const int p1 = c->Lnz[i];
            for(int p = c->Lp[i] + 1; p < p1; p++) w[c->Li[p]] -= c->Lx[p] * lki;
            d -= lki * lki;
            c->Li[p1] = k;
            c->Lx[p1] = lki;
            c->Lnz[i] = p1 + 1;
        }
        if(!(d > 0.)) return 0;   // not positive definite (also catches NaN)
        const int p = c->Lnz[k]++;
        c->Li[p] = k;
        c->Lx[p] = sqrt(d);
    }
    return 1;
}

// x <- (JtJ + lambda I)^-1 b
static void chol_solve(const chol_t* c, const double* b, double* x)
{
    const int n = c->n;
    double* y = (double*)malloc((size_t)n * sizeof(double));
    for(int k = 0; k < n; k++) y[k] = b[c->perm[k]];
    for(int j = 0; j < n; j++)
    {
        y[j] /= c->Lx[c->Lp[j]];
        for(int p = c->Lp[j] + 1; p < c->Lnz[j]; p++) y[c->Li[p]] -= c->Lx[p] * y[j];
    }
    for(int j = n - 1; j >= 0; j--)
    {
        for(int p = c->Lp[j] + 1; p < c->Lnz[j]; p++) y[j] -= c->Lx[p] * y[c->Li[p]];
        y[j] /= c->Lx[c->Lp[j]];
    }
    for(int k = 0; k < n; k++) x[c->perm[k]] = y[k];
    free(y);
}

////////////////////////////////////////////////////////////////////////////////
// The operating point and the context
////////////////////////////////////////////////////////////////////////////////
struct dogleg_port_point
{
    // must start with what mrcal.c reads: beforeStep->p, ->x (mrcal.c:6471,6517,6615-6617)
    dogleg_operatingPoint_t pub;
    cholmod_sparse Jt;          // sparse: row-wise J, as the callback fills it
    double* Jdense;             // dense: [Nmeas][Nstate]
    double* Jt_x;
    double* updateCauchy;       double updateCauchy_lensq;  int updateCauchy_valid;
    double* updateGN;           double updateGN_lensq;      int updateGN_valid;
    int     didStepToEdgeOfTrustRegion;
};
typedef struct dogleg_port_point point_t;

typedef struct
{
    dogleg_solverContext_t pub;   // first: the pointer mrcal.c holds is a pointer to this
    point_t* before;
    point_t* after;
    int Nstate, Nmeas, NJnnz, is_sparse;
    dogleg_callback_t* f;
    dogleg_callback_dense_t* f_dense;
    void* cookie;
    dogleg_parameters2_t par;
    double lambda;
    chol_t chol;
    // statistics
    int Nevaluations, Nfactorizations, Niterations, Nsymbolic;
    double t_callback, t_factor, t_products, t_total;
} ctx_t;

// the last finished solve, for the test driver (oracle/ref.py)
static struct
{
    int Nevaluations, Nfactorizations, Niterations, Nsymbolic, Lnnz, Npasses;
    double t_callback, t_factor, t_products, t_total, lambda, norm2_x;
} g_last, g_accum;

void dogleg_port_get_stats(double* out16)
{
    out16[0] = g_accum.Niterations; out16[1] = g_accum.Nevaluations; out16[2] = g_accum.Nfactorizations;
    out16[3] = g_accum.Nsymbolic;   out16[4] = g_last.Lnnz;          out16[5] = g_accum.Npasses;
    out16[6] = g_accum.t_callback;  out16[7] = g_accum.t_factor;     out16[8] = g_accum.t_products;
    out16[9] = g_accum.t_total;     out16[10] = g_last.lambda;       out16[11] = g_last.norm2_x;
    out16[12] = g_last.Niterations; out16[13] = 0; out16[14] = 0; out16[15] = 0;
}
void dogleg_port_reset_stats(void) { memset(&g_accum, 0, sizeof(g_accum)); memset(&g_last, 0, sizeof(g_last)); }

// iteration cap for bounded benchmark samples: <= 0 means "use the caller's max_iterations"
static int g_iteration_cap = 0;
void dogleg_port_set_iteration_cap(int n) { g_iteration_cap = n; }

static point_t* point_alloc(int Nstate, int Nmeas, int NJnnz, int is_sparse)
{
    point_t* pt = (point_t*)calloc(1, sizeof(point_t));
    pt->pub.p        = (double*)calloc((size_t)Nstate, sizeof(double));
    pt->pub.x        = (double*)calloc((size_t)Nmeas, sizeof(double));
    pt->Jt_x         = (double*)calloc((size_t)Nstate, sizeof(double));
    pt->updateCauchy = (double*)calloc((size_t)Nstate, sizeof(double));
    pt->updateGN     = (double*)calloc((size_t)Nstate, sizeof(double));
    if(is_sparse)
    {
        pt->Jt.nrow = (size_t)Nstate; pt->Jt.ncol = (size_t)Nmeas; pt->Jt.nzmax = (size_t)NJnnz;
        pt->Jt.p = calloc((size_t)Nmeas + 1, sizeof(int));
        pt->Jt.i = calloc((size_t)(NJnnz > 0 ? NJnnz : 1), sizeof(int));
        pt->Jt.x = calloc((size_t)(NJnnz > 0 ? NJnnz : 1), sizeof(double));
        pt->Jt.sorted = 1; pt->Jt.packed = 1;
    }
    else
        pt->Jdense = (double*)calloc((size_t)Nmeas * Nstate, sizeof(double));
    return pt;
}
static void point_free(point_t* pt)
{
    if(!pt) return;
    free(pt->pub.p); free(pt->pub.x); free(pt->Jt_x); free(pt->updateCauchy); free(pt->updateGN);
    free(pt->Jt.p); free(pt->Jt.i); free(pt->Jt.x); free(pt->Jdense);
    free(pt);
}

static double norm2(const double* v, int n) { double s = 0.; for(int i = 0; i < n; i++) s += v[i] * v[i]; return s; }
static double inner(const double* a, const double* b, int n) { double s = 0.; for(int i = 0; i < n; i++) s += a[i] * b[i]; return s; }

// |J v|^2 and (optionally) x . J v
static double norm2_J_times(const ctx_t* c, const point_t* pt, const double* v, double* x_dot_Jv)
{
    double s = 0., sx = 0.;
    if(c->is_sparse)
    {
        const int* Jp = (const int*)pt->Jt.p;
        const int* Ji = (const int*)pt->Jt.i;
        const double* Jx = (const double*)pt->Jt.x;
        for(int m = 0; m < c->Nmeas; m++)
        {
            double a = 0.;
            for(int e = Jp[m]; e < Jp[m + 1]; e++) a += Jx[e] * v[Ji[e]];
            s += a * a;
            sx += a * pt->pub.x[m];
        }
    }
    else
        for(int m = 0; m < c->Nmeas; m++)
        {
            const double a = inner(pt->Jdense + (size_t)m * c->Nstate, v, c->Nstate);
            s += a * a;
            sx += a * pt->pub.x[m];
        }
    if(x_dot_Jv) *x_dot_Jv = sx;
    return s;
}

// Evaluate the callback at pt->p. Returns 1 if the gradient is below Jt_x_threshold everywhere
static int compute_operating_point(ctx_t* c, point_t* pt)
{
    double t0 = now_s();
    if(c->is_sparse) c->f(pt->pub.p, pt->pub.x, &pt->Jt, c->cookie);
    else             c->f_dense(pt->pub.p, pt->pub.x, pt->Jdense, c->cookie);
    c->t_callback += now_s() - t0;
    c->Nevaluations++;
    t0 = now_s();
    pt->pub.norm2_x = norm2(pt->pub.x, c->Nmeas);
    memset(pt->Jt_x, 0, (size_t)c->Nstate * sizeof(double));
    if(c->is_sparse)
    {
        const int* Jp = (const int*)pt->Jt.p;
        const int* Ji = (const int*)pt->Jt.i;
        const double* Jx = (const double*)pt->Jt.x;
        for(int m = 0; m < c->Nmeas; m++)
        {
            const double xm = pt->pub.x[m];
            for(int e = Jp[m]; e < Jp[m + 1]; e++) pt->Jt_x[Ji[e]] += Jx[e] * xm;
        }
    }
    else
        for(int m = 0; m < c->Nmeas; m++)
            for(int i = 0; i < c->Nstate; i++) pt->Jt_x[i] += pt->Jdense[(size_t)m * c->Nstate + i] * pt->pub.x[m];
    c->t_products += now_s() - t0;
    pt->updateCauchy_valid = 0;
    pt->updateGN_valid = 0;
    for(int i = 0; i < c->Nstate; i++)
        if(fabs(pt->Jt_x[i]) > c->par.Jt_x_threshold) return 0;
    return 1;
}

static void compute_cauchy(ctx_t* c, point_t* pt)
{
    if(pt->updateCauchy_valid) return;
    pt->updateCauchy_valid = 1;
    // steepest descent direction -Jt x, length to the minimum of the quadratic model along it:
    // k = -|Jt x|^2 / |J Jt x|^2
    const double t0 = now_s();
    const double g2 = norm2(pt->Jt_x, c->Nstate);
    const double Jg2 = norm2_J_times(c, pt, pt->Jt_x, NULL);
    const double k = -g2 / Jg2;
    pt->updateCauchy_lensq = k * k * g2;
    for(int i = 0; i < c->Nstate; i++) pt->updateCauchy[i] = k * pt->Jt_x[i];
    c->t_products += now_s() - t0;
}

// dense (tiny) problems: Cholesky of JtJ + lambda I in place
static int dense_factor_solve(ctx_t* c, point_t* pt, double lambda)
{
    const int n = c->Nstate;
    double* A = (double*)calloc((size_t)n * n, sizeof(double));
    for(int m = 0; m < c->Nmeas; m++)
    {
        const double* r = pt->Jdense + (size_t)m * n;
        for(int i = 0; i < n; i++)
            for(int j = 0; j <= i; j++) A[i * n + j] += r[i] * r[j];
    }
    for(int i = 0; i < n; i++) A[i * n + i] += lambda;
    int ok = 1;
    for(int j = 0; j < n && ok; j++)
    {
        double d = A[j * n + j];
        for(int k = 0; k < j; k++) d -= A[j * n + k] * A[j * n + k];
        if(!(d > 0.)) { ok = 0; break; }
        A[j * n + j] = sqrt(d);
        for(int i = j + 1; i < n; i++)
        {
            double t = A[i * n + j];
            for(int k = 0; k < j; k++) t -= A[i * n + k] * A[j * n + k];
            A[i * n + j] = t / A[j * n + j];
        }
    }
    if(ok)
    {
        double* y = pt->updateGN;
        for(int i = 0; i < n; i++)
        {
            double t = pt->Jt_x[i];
            for(int k = 0; k < i; k++) t -= A[i * n + k] * y[k];
            y[i] = t / A[i * n + i];
        }
        for(int i = n - 1; i >= 0; i--)
        {
            double t = y[i];
            for(int k = i + 1; k < n; k++) t -= A[k * n + i] * y[k];
            y[i] = t / A[i * n + i];
        }
    }
    free(A);
    return ok;
}

static int compute_gauss_newton(ctx_t* c, point_t* pt)
{
    if(pt->updateGN_valid) return 1;
    const double t0 = now_s();
    while(1)
    {
        int ok;
        if(c->is_sparse)
        {
            if(!chol_pattern_matches(&c->chol, &pt->Jt))
            {
                if(!chol_symbolic(&c->chol, &pt->Jt)) return 0;
                c->Nsymbolic++;
            }
            ok = chol_numeric(&c->chol, &pt->Jt, c->lambda);
            if(ok) chol_solve(&c->chol, pt->Jt_x, pt->updateGN);
        }
        else
            ok = dense_factor_solve(c, pt, c->lambda);
        c->Nfactorizations++;
        if(ok) break;
        // singular JtJ: raise lambda and go again; lambda stays for the rest of this solve
        if(c->lambda == 0.) c->lambda = 1e-10;
        else                c->lambda *= 10.;
        if(!isfinite(c->lambda) || c->lambda > 1e30) { SAY("JtJ stays singular with lambda=%g", c->lambda); return 0; }
        if(c->par.dogleg_debug) SAY("singular JtJ. Adding %g I from now on", c->lambda);
    }
    for(int i = 0; i < c->Nstate; i++) pt->updateGN[i] = -pt->updateGN[i];
    pt->updateGN_lensq = norm2(pt->updateGN, c->Nstate);
    pt->updateGN_valid = 1;
    c->t_factor += now_s() - t0;
    return 1;
}

// Step from `from` within the trust region; writes p_new. Returns the expected improvement
// |x|^2 - |x + J step|^2, or NAN on failure
static double take_step_from(ctx_t* c, point_t* from, double* p_new, double* step, double* step_len_sq, double trustregion)
{
    const int n = c->Nstate;
    compute_cauchy(c, from);
    if(from->updateCauchy_lensq >= trustregion * trustregion)
    {
        // Cauchy point outside the trust region: go along the gradient to the edge
        const double k = trustregion / sqrt(from->updateCauchy_lensq);
        for(int i = 0; i < n; i++) step[i] = k * from->updateCauchy[i];
        *step_len_sq = trustregion * trustregion;
        from->didStepToEdgeOfTrustRegion = 1;
    }
    else
    {
        if(!compute_gauss_newton(c, from)) return NAN;
        if(from->updateGN_lensq <= trustregion * trustregion)
        {
            memcpy(step, from->updateGN, (size_t)n * sizeof(double));
            *step_len_sq = from->updateGN_lensq;
            from->didStepToEdgeOfTrustRegion = 0;
        }
        else
        {
            // dogleg: a + k (b-a) on the boundary, a = Cauchy, b = Gauss-Newton, k in [0,1]
            const double* a = from->updateCauchy;
            const double* b = from->updateGN;
            const double dsq = trustregion * trustregion;
            double l2 = 0., neg_c = 0.;
            for(int i = 0; i < n; i++) { const double d = a[i] - b[i]; l2 += d * d; neg_c += d * a[i]; }
            double disc = neg_c * neg_c - l2 * (from->updateCauchy_lensq - dsq);
            if(disc < 0.) disc = 0.;
            const double k = (neg_c + sqrt(disc)) / l2;
            double lensq = 0.;
            for(int i = 0; i < n; i++) { step[i] = a[i] + k * (b[i] - a[i]); lensq += step[i] * step[i]; }
            *step_len_sq = lensq;
            from->didStepToEdgeOfTrustRegion = 1;
        }
    }
    for(int i = 0; i < n; i++) p_new[i] = from->pub.p[i] + step[i];
    const double t0 = now_s();
    double x_Js = 0.;
    const double Js2 = norm2_J_times(c, from, step, &x_Js);
    c->t_products += now_s() - t0;
    return -2. * x_Js - Js2;
}

static double run_optimizer(ctx_t* c)
{
    const double t_start = now_s();
    double trustregion = c->par.trustregion0;
    int stepCount = 0;
    int max_iterations = c->par.max_iterations;
    if(g_iteration_cap > 0 && g_iteration_cap < max_iterations) max_iterations = g_iteration_cap;
    double* step = (double*)malloc((size_t)c->Nstate * sizeof(double));
    int done = compute_operating_point(c, c->before);
    if(c->par.dogleg_debug) SAY("Initial operating point has norm2_x %.12g", c->before->pub.norm2_x);
    while(!done && stepCount < max_iterations)
    {
        while(1)
        {
            double step_len_sq;
            const double expected = take_step_from(c, c->before, c->after->pub.p, step, &step_len_sq, trustregion);
            if(isnan(expected)) { free(step); return -1.; }
            // libdogleg compares the SQUARED step length with update_threshold
            if(step_len_sq < c->par.update_threshold)
            {
                if(c->par.dogleg_debug) SAY("update small enough (%g). Done iterating", step_len_sq);
                done = 1;
                break;
            }
            const int after_zero_gradient = compute_operating_point(c, c->after);
            const double observed = c->before->pub.norm2_x - c->after->pub.norm2_x;
            const double rho = observed / expected;
            if(c->par.dogleg_debug)
                SAY("step %d: norm2_x %.12g -> %.12g |step| %.4g rho %.4g trustregion %.4g lambda %g", stepCount,
                    c->before->pub.norm2_x, c->after->pub.norm2_x, sqrt(step_len_sq), rho, trustregion, c->lambda);
            if(rho < c->par.trustregion_decrease_threshold)
                trustregion *= c->par.trustregion_decrease_factor;
            else if(rho > c->par.trustregion_increase_threshold && c->before->didStepToEdgeOfTrustRegion)
                trustregion *= c->par.trustregion_increase_factor;
            if(rho > 0.)
            {
                point_t* t = c->before; c->before = c->after; c->after = t;
                c->pub.beforeStep = &c->before->pub;
                if(after_zero_gradient) done = 1;
                break;
            }
            // rejected: same operating point, smaller trust region
            if(trustregion < c->par.trustregion_threshold) { done = 1; break; }
        }
        if(done) break;
        stepCount++;
    }
    free(step);
    c->Niterations = stepCount;
    c->t_total = now_s() - t_start;
    return c->before->pub.norm2_x;
}

static double optimize_common(double* p, int Nstate, int Nmeas, int NJnnz, int is_sparse,
                              dogleg_callback_t* f, dogleg_callback_dense_t* fd, void* cookie,
                              const dogleg_parameters2_t* parameters, dogleg_solverContext_t** returnContext)
{
    ctx_t* c = (ctx_t*)calloc(1, sizeof(ctx_t));
    c->Nstate = Nstate; c->Nmeas = Nmeas; c->NJnnz = NJnnz; c->is_sparse = is_sparse;
    c->f = f; c->f_dense = fd; c->cookie = cookie;
    c->par = *parameters;
    if(getenv("DOGLEG_PORT_VERBOSE")) c->par.dogleg_debug = 1;
    c->before = point_alloc(Nstate, Nmeas, NJnnz, is_sparse);
    c->after  = point_alloc(Nstate, Nmeas, NJnnz, is_sparse);
    c->pub.beforeStep = &c->before->pub;
    memcpy(c->before->pub.p, p, (size_t)Nstate * sizeof(double));
    const double norm2_x = run_optimizer(c);
    memcpy(p, c->before->pub.p, (size_t)Nstate * sizeof(double));

    g_last.Niterations = c->Niterations; g_last.Nevaluations = c->Nevaluations;
    g_last.Nfactorizations = c->Nfactorizations; g_last.Nsymbolic = c->Nsymbolic;
    g_last.Lnnz = c->chol.Lnnz; g_last.lambda = c->lambda; g_last.norm2_x = norm2_x;
    g_last.t_callback = c->t_callback; g_last.t_factor = c->t_factor; g_last.t_products = c->t_products; g_last.t_total = c->t_total;
    if(is_sparse)
    {
        g_accum.Niterations += c->Niterations; g_accum.Nevaluations += c->Nevaluations;
        g_accum.Nfactorizations += c->Nfactorizations; g_accum.Nsymbolic += c->Nsymbolic; g_accum.Npasses++;
        g_accum.t_callback += c->t_callback; g_accum.t_factor += c->t_factor; g_accum.t_products += c->t_products; g_accum.t_total += c->t_total;
    }
    if(returnContext) *returnContext = &c->pub;
    else { dogleg_solverContext_t* pc = &c->pub; dogleg_freeContext(&pc); }
    return norm2_x;
}

////////////////////////////////////////////////////////////////////////////////
// libdogleg's interface, as far as mrcal.c uses it
////////////////////////////////////////////////////////////////////////////////
void dogleg_getDefaultParameters(dogleg_parameters2_t* p)
{
    memset(p, 0, sizeof(*p));
    p->max_iterations = 100;
    p->dogleg_debug = 0;
    p->trustregion0 = 1.0e3;
    p->trustregion_decrease_factor = 0.1;
    p->trustregion_decrease_threshold = 0.25;
    p->trustregion_increase_factor = 2.0;
    p->trustregion_increase_threshold = 0.75;
    p->Jt_x_threshold = 1e-8;
    p->update_threshold = 1e-8;
    p->trustregion_threshold = 1e-8;
}

double dogleg_optimize2(double* p, unsigned int Nstate, unsigned int Nmeas, unsigned int NJnnz,
                        dogleg_callback_t* f, void* cookie,
                        const dogleg_parameters2_t* parameters, dogleg_solverContext_t** returnContext)
{
    return optimize_common(p, (int)Nstate, (int)Nmeas, (int)NJnnz, 1, f, NULL, cookie, parameters, returnContext);
}

double dogleg_optimize_dense2(double* p, unsigned int Nstate, unsigned int Nmeas,
                              dogleg_callback_dense_t* f, void* cookie,
                              const dogleg_parameters2_t* parameters, dogleg_solverContext_t** returnContext)
{
    return optimize_common(p, (int)Nstate, (int)Nmeas, 0, 0, NULL, f, cookie, parameters, returnContext);
}

void dogleg_freeContext(dogleg_solverContext_t** pub)
{
    if(!pub || !*pub) return;
    ctx_t* c = (ctx_t*)*pub;   // pub is the first member
    point_free(c->before);
    point_free(c->after);
    chol_free(&c->chol);
    free(c);
    *pub = NULL;
}

// Finite-difference check of one column of the Jacobian, reported as a vnlog the way
// libdogleg's dogleg_testGradient() does: one line per measurement with a nonzero reported or
// observed gradient
void dogleg_testGradient(unsigned int var, const double* p0,
                         unsigned int Nstate, unsigned int Nmeas, unsigned int NJnnz,
                         dogleg_callback_t* f, void* cookie)
{
    const double delta = 1e-6;
    point_t* a = point_alloc((int)Nstate, (int)Nmeas, (int)NJnnz, 1);
    point_t* b = point_alloc((int)Nstate, (int)Nmeas, (int)NJnnz, 1);
    memcpy(a->pub.p, p0, Nstate * sizeof(double));
    memcpy(b->pub.p, p0, Nstate * sizeof(double));
    b->pub.p[var] += delta;
    f(a->pub.p, a->pub.x, &a->Jt, cookie);
    f(b->pub.p, b->pub.x, &b->Jt, cookie);
    if(var == 0) printf("# ivar imeasurement gradient_reported gradient_observed error error_relative\n");
    const int* Jp = (const int*)a->Jt.p;
    const int* Ji = (const int*)a->Jt.i;
    const double* Jx = (const double*)a->Jt.x;
    for(unsigned int m = 0; m < Nmeas; m++)
    {
        double rep = 0.;
        int have = 0;
        for(int e = Jp[m]; e < Jp[m + 1]; e++) if(Ji[e] == (int)var) { rep += Jx[e]; have = 1; }
        const double obs = (b->pub.x[m] - a->pub.x[m]) / delta;
        if(!have && obs == 0.) continue;
        const double err = rep - obs;
        const double den = (fabs(rep) + fabs(obs)) / 2.;
        printf("%u %u %.6g %.6g %.6g %.6g\n", var, m, rep, obs, err, den > 0. ? fabs(err) / den : 0.);
    }
    point_free(a);
    point_free(b);
}
