// sep_kernels.cuh -- "line-coefficient" form of an assembled staggered-grid operator (single GPU).
//
// What b200ls_staggered_analyze (staggered.cpp) reads out of the matrix given to setMatrix: up to three
// staggered fields stored one after the other ([u | v | w] of the velocity system A = I/dt - c nu L,
// navierstokes.cpp:342-344 / createlaplacian.cpp:134-159; or the pressure block of IBPM's modified Poisson
// system, ibpm.cpp:100-203), each a 5-/7-point stencil whose off-diagonal coefficient in direction d depends
// only on the index along d (two 1-D arrays per field and direction), the diagonal as a vector, and a CSR
// remainder for everything behind the stencil blocks (IBPM's Lagrangian coupling); on a stretched grid the pressure
// block of IBPM's system carries face areas on top of the 1-D face arrays (SepField::w).  A row costs
// 8 (x) + 8 (diag) + 8 (y) bytes of HBM traffic instead of ~100 for the assembled CSR row; the terms are added
// in ascending column order without FMA contraction, so the result is bit-identical to MatMult_SeqAIJ on the
// assembled matrix (the remainder columns all lie behind the stencil columns of their row).
//
// The Krylov kernels are the ones of csr_kernels.cuh with the row product swapped: same reductions, same
// device-side KSP logic (finalize_scalars / finalize_scalars_csr), same pointwise kernels.
#pragma once
#include <stdint.h>

#include "csr_kernels.cuh"

namespace b200 {

struct SepField
{
    int n0, n1, n2;            // extents of the field (n2 = 1 in 2-D)
    int per0, per1, per2;      // periodic wrap (only set when the extent is >= 3)
    long long off;             // first row of the field in the packed vector
    const double *cm[3];       // coefficient of the minus neighbour in direction d at index s (0: no entry)
    const double *cp[3];       // plus neighbour
    // non-null: the block is the pressure operator D (dt I) G of a stretched grid (IBPM's modified Poisson system): the
    // coefficient in direction d is (product of the two OTHER cell widths) * cm/cp, with cm/cp the face arrays dt/h --
    // the grouping of the assembled matrix (createdivergence.cpp:142,146,150)
    const double *w[3];
};

struct SepDev
{
    int nf;                    // stencil blocks (fields)
    long long nsep;            // rows covered by the stencil blocks
    long long nrows;
    SepField f[3];
    const double *diag;        // [nsep]
    const int64_t *rem_rowptr; // remainder CSR over all rows, or null when it is empty
    const int32_t *rem_col;
    const double *rem_val;
};

// one row of A times a vector given by its accessor; terms in ascending column order
template <class Fetch>
__device__ __forceinline__ double sep_row(const SepDev &A, long long i, Fetch fetch)
{
    double t = 0.0;
    if (i < A.nsep)
    {
        int fi = 0;
        if (A.nf > 1 && i >= A.f[1].off) fi = 1;
        if (A.nf > 2 && i >= A.f[2].off) fi = 2;
        const SepField &f = A.f[fi];
        const unsigned int l = (unsigned int)(i - f.off);  // the system has fewer than 2^31 rows (checked on the host)
        const int n0 = f.n0, n1 = f.n1, n2 = f.n2;
        const unsigned int row1 = l / (unsigned int)n0;
        const int i0 = (int)(l - row1 * (unsigned int)n0);
        const int i2 = (int)(row1 / (unsigned int)n1);
        const int i1 = (int)(row1 - (unsigned int)i2 * (unsigned int)n1);
        const long long s1 = n0, s2 = (long long)n0 * n1;
        double cxm = f.cm[0][i0], cxp = f.cp[0][i0];
        double cym = f.cm[1][i1], cyp = f.cp[1][i1];
        double czm = f.cm[2][i2], czp = f.cp[2][i2];
        if (f.w[0])
        {
            const double dxi = f.w[0][i0], dyj = f.w[1][i1], dzk = f.w[2][i2];
            const double ayz = __dmul_rn(dyj, dzk), axz = __dmul_rn(dxi, dzk), axy = __dmul_rn(dxi, dyj);
            cxm = __dmul_rn(ayz, cxm);
            cxp = __dmul_rn(ayz, cxp);
            cym = __dmul_rn(axz, cym);
            cyp = __dmul_rn(axz, cyp);
            czm = __dmul_rn(axy, czm);
            czp = __dmul_rn(axy, czp);
        }
        const bool xlo = f.per0 && i0 == 0, xhi = f.per0 && i0 == n0 - 1;
        const bool ylo = f.per1 && i1 == 0, yhi = f.per1 && i1 == n1 - 1;
        const bool zlo = f.per2 && i2 == 0, zhi = f.per2 && i2 == n2 - 1;
        // columns of the six neighbours (a wrapped neighbour lands on the far side of the sorted row)
        const long long jxm = xlo ? i + (n0 - 1) : i - 1, jxp = xhi ? i - (n0 - 1) : i + 1;
        const long long jym = ylo ? i + (n1 - 1) * s1 : i - s1, jyp = yhi ? i - (n1 - 1) * s1 : i + s1;
        const long long jzm = zlo ? i + (n2 - 1) * s2 : i - s2, jzp = zhi ? i - (n2 - 1) * s2 : i + s2;
        // a zero coefficient stands for "no entry in the assembled row" (wall-side neighbour): nothing is fetched
#define B200_SEP_TERM(c, j) \
    if ((c) != 0.0) t = __dadd_rn(t, __dmul_rn((c), fetch(j)))
        // sorted columns: zp(wrapped) < zm < yp(w) < ym < xp(w) < xm < diag < xp < xm(w) < yp < ym(w) < zp < zm(w)
        if (zhi) B200_SEP_TERM(czp, jzp);
        if (!zlo) B200_SEP_TERM(czm, jzm);
        if (yhi) B200_SEP_TERM(cyp, jyp);
        if (!ylo) B200_SEP_TERM(cym, jym);
        if (xhi) B200_SEP_TERM(cxp, jxp);
        if (!xlo) B200_SEP_TERM(cxm, jxm);
        t = __dadd_rn(t, __dmul_rn(A.diag[i], fetch(i)));
        if (!xhi) B200_SEP_TERM(cxp, jxp);
        if (xlo) B200_SEP_TERM(cxm, jxm);
        if (!yhi) B200_SEP_TERM(cyp, jyp);
        if (ylo) B200_SEP_TERM(cym, jym);
        if (!zhi) B200_SEP_TERM(czp, jzp);
        if (zlo) B200_SEP_TERM(czm, jzm);
#undef B200_SEP_TERM
    }
    if (A.rem_rowptr)
        for (int64_t q = A.rem_rowptr[i]; q < A.rem_rowptr[i + 1]; ++q)
            t = __dadd_rn(t, __dmul_rn(A.rem_val[q], fetch((long long)A.rem_col[q])));
    return t;
}

// plain y = A x (b200ls_apply)
__global__ void __launch_bounds__(256) k_sep_apply(SepDev A, const double *x, double *y)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < A.nrows; i += (long long)gridDim.x * blockDim.x)
        y[i] = sep_row(A, i, [&](long long j) { return x[j]; });
}

// CG class 0:  x += a' p' ; p = z + b p' ; w = A p ; dpi = p.w   (p rebuilt on the fly per column; k_csr_cg_spmv)
template <bool JACOBI, int NULLMODE>
__global__ void __launch_bounds__(256) k_sep_cg_spmv(SepDev A, CsrVecs v, ReduceWs ws, DevState *st, SolveConsts kc, double *hist)
{
    if (st->done) return;
    const double shift = st->shift, bcoef = st->b, aprev = st->a;
    const bool xupd = st->pending != 0;
    double acc[1] = {0.0};
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < A.nrows; i += (long long)gridDim.x * blockDim.x)
    {
        const double pold = v.p_in[i];
        if (xupd) v.x[i] = __dadd_rn(v.x[i], __dmul_rn(aprev, pold));
        const double pi = __dadd_rn(csr_z<JACOBI, NULLMODE>(v, i, shift), __dmul_rn(bcoef, pold));
        v.p_out[i] = pi;
        const double t = sep_row(A, i, [&](long long j) {
            return __dadd_rn(csr_z<JACOBI, NULLMODE>(v, j, shift), __dmul_rn(bcoef, v.p_in[j]));
        });
        v.w[i] = t;
        acc[0] = fma(pi, t, acc[0]);
    }
    csr_reduce_finalize<1>(acc, FIN_SPMV, ws, st, kc, hist);
}

// BiCGStab: v = B A p ; d1 = v.rp   (k_bcgs_spmv1)
template <bool JACOBI>
__global__ void __launch_bounds__(256) k_sep_bcgs_spmv1(SepDev A, const double *p, const double *dinv, const double *rp,
                                                        double *vv, ReduceWs ws, DevState *st, SolveConsts kc, double *hist)
{
    if (st->done) return;
    double acc[1] = {0.0};
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < A.nrows; i += (long long)gridDim.x * blockDim.x)
    {
        double t = sep_row(A, i, [&](long long j) { return p[j]; });
        if (JACOBI) t = __dmul_rn(t, dinv[i]);
        vv[i] = t;
        acc[0] = fma(t, rp[i], acc[0]);
    }
    csr_reduce_finalize<1>(acc, FIN_BCGS_D1, ws, st, kc, hist);
}

// BiCGStab: s = r - alpha v (rebuilt on the fly per column) ; t = B A s ; sums {s.t, t.t, s.s}   (k_bcgs_spmv2)
template <bool JACOBI>
__global__ void __launch_bounds__(256) k_sep_bcgs_spmv2(SepDev A, const double *r, const double *vv, const double *dinv,
                                                        double *s, double *t_out, ReduceWs ws, DevState *st, SolveConsts kc,
                                                        double *hist)
{
    if (st->done) return;
    const double malpha = -st->alpha;
    double acc[3] = {0, 0, 0};
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < A.nrows; i += (long long)gridDim.x * blockDim.x)
    {
        const double si = __dadd_rn(__dmul_rn(malpha, vv[i]), r[i]);  // VecWAXPY(S, -alpha, V, R)
        s[i] = si;
        double t = sep_row(A, i, [&](long long j) { return __dadd_rn(__dmul_rn(malpha, vv[j]), r[j]); });
        if (JACOBI) t = __dmul_rn(t, dinv[i]);
        t_out[i] = t;
        acc[0] = fma(si, t, acc[0]);
        acc[1] = fma(t, t, acc[1]);
        acc[2] = fma(si, si, acc[2]);
    }
    csr_reduce_finalize<3>(acc, FIN_BCGS_OMEGA, ws, st, kc, hist);
}

// ---- preconditioned CG with an explicit z0 = M^-1 r (pc_type mg on the hybrid operator: multigrid on the pressure block,
// diagonal scaling on the rows behind it).  k_sep_cg_spmv is used unchanged with r := z0 and JACOBI = false; the update
// kernel of the CSR path is split into k_mg_rupdate (r -= a w), the preconditioner, and these sums.

// the rows behind the stencil block: z0 = D^-1 r
__global__ void __launch_bounds__(256) k_sep_tail_pc(long long nsep, long long nrows, const double *r, const double *dinv,
                                                     double *z0, const DevState *st)
{
    if (st->done) return;
    for (long long i = nsep + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nrows; i += (long long)gridDim.x * blockDim.x)
        z0[i] = __dmul_rn(r[i], dinv[i]);
}

// the six sums of k_csr_cg_update for z0 given as a vector
template <int NULLMODE>
__global__ void __launch_bounds__(256) k_sep_zsums(long long n, const double *z0v, const double *r, const double *nv, int fin_kind,
                                                   ReduceWs ws, DevState *st, SolveConsts kc, double *hist)
{
    if (st->done) return;
    const double c = st->c;
    double acc[6] = {0, 0, 0, 0, 0, 0};
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    {
        const double z0 = z0v[i], rn = r[i];
        if (NULLMODE == 2)
        {
            const double nvi = nv[i];
            acc[0] = fma(z0, nvi, acc[0]);
            acc[1] = fma(z0, z0, acc[1]);
            acc[2] = fma(z0, rn, acc[2]);
            acc[3] = fma(nvi, rn, acc[3]);
            acc[4] = fma(nvi, nvi, acc[4]);
            acc[5] = fma(rn, rn, acc[5]);
        }
        else
        {
            const double d0 = z0 - c;
            acc[0] += z0;
            acc[1] += d0;
            acc[2] = fma(d0, d0, acc[2]);
            acc[3] = fma(d0, rn, acc[3]);
            acc[4] += rn;
            acc[5] = fma(rn, rn, acc[5]);
        }
    }
    csr_reduce_finalize<6>(acc, fin_kind, ws, st, kc, hist);
}

}  // namespace b200
