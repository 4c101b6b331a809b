// csr_kernels.cuh -- general assembled-operator path (single GPU).
//
// For operators the separable stencil cannot express -- PetIBM's IBPM modified Poisson system
// [D;E] BN [G,-H] (applications/ibpm/ibpm.cpp:100-203), the velocity system A = I/dt - c nu L
// (navierstokes.cpp:342-344, solved with bcgs + jacobi in the shipped configs) and BN order > 1 -- the
// matrix given to setMatrix is kept as CSR on the device.  Same two-kernel CG structure and device-side
// KSP logic as the stencil path; MatMult_SeqAIJ's serial per-row summation order is kept (one thread per
// row, no FMA contraction), so the SpMV is bit-identical to the reference's assembled MatMult.
// BiCGStab follows KSPSolve_BCGS (bcgs.c) with left preconditioning: 4 kernels and 3 grid reductions per
// iteration (PETSc: 2 MatMult, 2 PCApply, 6 vector passes, 4 reductions).
#pragma once
#include <stdint.h>

#include "kernels.cuh"

namespace b200 {

struct CsrDev
{
    int64_t nrows;
    const int64_t *rowptr;
    const int32_t *col;
    const double *val;
};

// additional reduction kinds (finalize_scalars_csr)
enum
{
    FIN_CSR_UPDATE = 10,   // CG with one explicit null-space vector
    FIN_CSR_INIT = 11,
    FIN_BCGS_INIT = 20,
    FIN_BCGS_D1 = 21,
    FIN_BCGS_OMEGA = 22,
    FIN_BCGS_UPD = 23
};

struct CsrVecs
{
    const double *r;
    const double *p_in;
    double *p_out;
    double *w;
    double *x;
    const double *dinv;  // Jacobi or null
    const double *nv;    // explicit orthonormal null-space vector or null
};

// z_j = B r_j with the null space removed, recomputed wherever it is needed
template <bool JACOBI, int NULLMODE>
__device__ __forceinline__ double csr_z(const CsrVecs &v, int64_t j, double shift)
{
    double z = v.r[j];
    if (JACOBI) z = __dmul_rn(z, v.dinv[j]);
    if (NULLMODE == 1) z = __dadd_rn(z, shift);
    if (NULLMODE == 2) z = __dadd_rn(z, __dmul_rn(shift, v.nv[j]));
    return z;
}

// scalar logic for the reductions that only the CSR path has
__device__ inline void finalize_scalars_csr(int kind, const double *S, DevState &s, const SolveConsts &k, double *hist)
{
    if (kind == FIN_CSR_UPDATE || kind == FIN_CSR_INIT)
    {
        // S: {z0.nv, z0.z0, z0.r, nv.r, nv.nv, r.r}; MatNullSpaceRemove: z += (-(z0.nv)) nv
        const double sv = -S[0];
        const double zz = S[1] + 2.0 * sv * S[0] + sv * sv * S[4];
        const double zr = S[2] + sv * S[3];
        const double T[6] = {0.0, 0.0, zz, zr, 0.0, S[5]};
        // reuse the CG logic with shift := sv and centre 0 (d = z0)
        SolveConsts kk = k;
        kk.has_const = 0;
        s.c = 0.0;
        finalize_scalars(kind == FIN_CSR_INIT ? FIN_INIT : FIN_UPDATE, T, s, kk, hist);
        s.shift = sv;
        s.c = 0.0;
        return;
    }
    if (kind == FIN_BCGS_INIT)
    {
        // dp = ||B b||; rp = r so rho = (r, rp) = ||r||^2 as a separately accumulated sum
        const double dp = (k.norm_type != 0) ? sqrt(S[0]) : 0.0;
        s.dp = dp;
        if (s.nhist < k.hist_cap) hist[s.nhist] = dp;
        s.nhist++;
        int reason = converged_default(s, k, 0, dp);
        s.i = 0;
        s.its = 0;
        s.rhoold = 1.0;
        s.alpha = 1.0;
        s.omegaold = 1.0;
        s.rho = S[1];
        if (!reason && k.max_it <= 0) reason = -3;
        if (!reason)
        {
            s.beta = (s.rho / s.rhoold) * (s.alpha / s.omegaold);
            s.b = -s.omegaold * s.beta;  // coefficient of v in VecAXPBYPCZ(P, 1, -omegaold*beta, beta, R, V)
        }
        else
        {
            s.reason = reason;
            s.done = 1;
        }
        return;
    }
    if (kind == FIN_BCGS_D1)
    {
        const double d1 = S[0];
        if (isnan(d1) || isinf(d1))
        {
            s.reason = -9;
            s.done = 1;
        }
        else if (d1 == 0.0)
        {
            s.reason = -5;  // KSP_DIVERGED_BREAKDOWN
            s.done = 1;
        }
        else
            s.alpha = s.rho / d1;
        return;
    }
    if (kind == FIN_BCGS_OMEGA)
    {
        const double d1 = S[0], d2 = S[1], ss = S[2];
        if (d2 == 0.0)
        {
            if (ss != 0.0)
            {
                s.reason = -5;
                s.done = 1;
                return;
            }
            // t = s = 0: x += alpha p is the exact answer
            s.pending = 2;
            s.its = s.i + 1;
            s.dp = 0.0;
            if (s.nhist < k.hist_cap) hist[s.nhist] = 0.0;
            s.nhist++;
            s.reason = 2;
            s.done = 1;
            return;
        }
        s.omega = d1 / d2;
        return;
    }
    if (kind == FIN_BCGS_UPD)
    {
        const double dp = (k.norm_type != 0) ? sqrt(S[0]) : 0.0;
        s.dp = dp;
        s.its = s.i + 1;
        if (s.nhist < k.hist_cap) hist[s.nhist] = dp;
        s.nhist++;
        int reason = converged_default(s, k, s.i + 1, dp);
        if (!reason && s.rho == 0.0) reason = -5;
        if (!reason)
        {
            s.i = s.i + 1;
            if (s.i >= k.max_it) reason = -3;
        }
        if (!reason)
        {
            s.rhoold = s.rho;
            s.omegaold = s.omega;
            s.rho = S[1];
            s.beta = (s.rho / s.rhoold) * (s.alpha / s.omegaold);
            s.b = -s.omegaold * s.beta;
        }
        else
        {
            s.reason = reason;
            s.done = 1;
        }
    }
}

// block + grid reduction for the CSR kernels (single GPU): same deterministic two-stage scheme
template <int NS>
__device__ __forceinline__ void csr_reduce_finalize(double (&acc)[NS], int kind, const ReduceWs &ws, DevState *st,
                                                    const SolveConsts &k, double *hist)
{
    if (kind < 10)
    {
        CommDev cm;
        cm.mode = 0;
        cm.rank = 0;
        cm.nranks = 1;
        cm.r_ghost_dn = cm.r_ghost_up = nullptr;
        grid_reduce_finalize<NS>(acc, kind, ws, cm, st, k, hist, false);
        return;
    }
    __shared__ double s_red[32][NS];
    __shared__ bool s_last;
    const int tid = threadIdx.x;
    const int lane = tid & 31, wid = tid >> 5, nw = (blockDim.x + 31) >> 5;
    const unsigned int nblocks = gridDim.x, bid = blockIdx.x;
#pragma unroll
    for (int q = 0; q < NS; ++q)
    {
        const double v = warp_sum(acc[q]);
        if (lane == 0) s_red[wid][q] = v;
    }
    __syncthreads();
    if (wid == 0)
    {
#pragma unroll
        for (int q = 0; q < NS; ++q)
        {
            double v = (lane < nw) ? s_red[lane][q] : 0.0;
            v = warp_sum(v);
            if (lane == 0) ws.partials[(size_t)bid * B200_NSUM + q] = v;
        }
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(ws.counter, 1u) == nblocks - 1);
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    double tot[NS];
#pragma unroll
    for (int q = 0; q < NS; ++q) tot[q] = 0.0;
    for (unsigned int b = tid; b < nblocks; b += blockDim.x)
#pragma unroll
        for (int q = 0; q < NS; ++q) tot[q] += __ldcg(&ws.partials[(size_t)b * B200_NSUM + q]);
    __syncthreads();
#pragma unroll
    for (int q = 0; q < NS; ++q)
    {
        const double v = warp_sum(tot[q]);
        if (lane == 0) s_red[wid][q] = v;
    }
    __syncthreads();
    if (wid != 0) return;
    double S[B200_NSUM];
#pragma unroll
    for (int q = 0; q < B200_NSUM; ++q) S[q] = 0.0;
#pragma unroll
    for (int q = 0; q < NS; ++q)
    {
        double v = (lane < nw) ? s_red[lane][q] : 0.0;
        S[q] = warp_sum(v);
    }
    if (lane == 0)
    {
        *ws.counter = 0u;
        finalize_scalars_csr(kind, S, *st, k, hist);
    }
}

// ------------------------------------------------------------------------------------------
// plain y = A x (b200ls_apply)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_csr_apply(CsrDev A, const double *x, double *y)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < A.nrows; i += (int64_t)gridDim.x * blockDim.x)
    {
        double t = 0.0;
        for (int64_t q = A.rowptr[i]; q < A.rowptr[i + 1]; ++q) t = __dadd_rn(t, __dmul_rn(A.val[q], x[A.col[q]]));
        y[i] = t;
    }
}

// ------------------------------------------------------------------------------------------
// CG class 0 on CSR:  x += a' p' ; p = z + b p' ; w = A p ; dpi = p.w   (p built on the fly per column)
// ------------------------------------------------------------------------------------------
template <bool JACOBI, int NULLMODE>
__global__ void __launch_bounds__(256) k_csr_cg_spmv(CsrDev A, CsrVecs v, ReduceWs ws, DevState *st, SolveConsts kc,
                                                     double *hist)
{
    if (st->done) return;
    const double shift = st->shift, bcoef = st->b, aprev = st->a;
    const bool xupd = st->pending != 0;
    double acc[1] = {0.0};
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < A.nrows; i += (int64_t)gridDim.x * blockDim.x)
    {
        const double pold = v.p_in[i];
        if (xupd) v.x[i] = __dadd_rn(v.x[i], __dmul_rn(aprev, pold));
        const double pi = __dadd_rn(csr_z<JACOBI, NULLMODE>(v, i, shift), __dmul_rn(bcoef, pold));
        v.p_out[i] = pi;
        double t = 0.0;
        for (int64_t q = A.rowptr[i]; q < A.rowptr[i + 1]; ++q)
        {
            const int64_t j = A.col[q];
            const double pj = __dadd_rn(csr_z<JACOBI, NULLMODE>(v, j, shift), __dmul_rn(bcoef, v.p_in[j]));
            t = __dadd_rn(t, __dmul_rn(A.val[q], pj));
        }
        v.w[i] = t;
        acc[0] = fma(pi, t, acc[0]);
    }
    csr_reduce_finalize<1>(acc, FIN_SPMV, ws, st, kc, hist);
}

// ------------------------------------------------------------------------------------------
// CG class 1 on CSR:  r -= a w ; sums ; KSP logic.  NULLMODE 0/1 share the stencil path's sums.
// ------------------------------------------------------------------------------------------
template <bool JACOBI, int NULLMODE, bool INIT>
__global__ void __launch_bounds__(256) k_csr_cg_update(int64_t n, double *r, const double *w, const double *dinv,
                                                       const double *nv, int fin_kind, ReduceWs ws, DevState *st,
                                                       SolveConsts kc, double *hist)
{
    if (st->done) return;
    const double ma = INIT ? 0.0 : -st->a;
    const double c = st->c;
    double acc[6] = {0, 0, 0, 0, 0, 0};
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    {
        double rn = r[i];
        if (!INIT)
        {
            rn = __dadd_rn(rn, __dmul_rn(ma, w[i]));
            r[i] = rn;
        }
        const double z0 = JACOBI ? __dmul_rn(rn, dinv[i]) : rn;
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

// ------------------------------------------------------------------------------------------
// BiCGStab (KSPSolve_BCGS, left preconditioning, zero initial guess)
// ------------------------------------------------------------------------------------------
// r = B b ; rp = r ; p = v = 0 ; sums {r.r, r.rp}
template <bool JACOBI>
__global__ void __launch_bounds__(256) k_bcgs_init(int64_t n, const double *b, const double *dinv, double *r, double *rp,
                                                   double *p, double *vv, ReduceWs ws, DevState *st, SolveConsts kc,
                                                   double *hist)
{
    double acc[2] = {0, 0};
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    {
        const double ri = JACOBI ? __dmul_rn(b[i], dinv[i]) : b[i];
        r[i] = ri;
        rp[i] = ri;
        p[i] = 0.0;
        vv[i] = 0.0;
        acc[0] = fma(ri, ri, acc[0]);
        acc[1] = fma(ri, ri, acc[1]);
    }
    csr_reduce_finalize<2>(acc, FIN_BCGS_INIT, ws, st, kc, hist);
}

// p = r + (-omegaold*beta) v + beta p      (VecAXPBYPCZ)
__global__ void __launch_bounds__(256) k_bcgs_p(int64_t n, const double *r, const double *vv, double *p, DevState *st)
{
    if (st->done) return;
    const double bq = st->b, beta = st->beta;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        p[i] = __dadd_rn(__dadd_rn(r[i], __dmul_rn(bq, vv[i])), __dmul_rn(beta, p[i]));
}

// v = B A p ; d1 = v.rp
template <bool JACOBI>
__global__ void __launch_bounds__(256) k_bcgs_spmv1(CsrDev A, const double *p, const double *dinv, const double *rp,
                                                    double *vv, ReduceWs ws, DevState *st, SolveConsts kc, double *hist)
{
    if (st->done) return;
    double acc[1] = {0.0};
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < A.nrows; i += (int64_t)gridDim.x * blockDim.x)
    {
        double t = 0.0;
        for (int64_t q = A.rowptr[i]; q < A.rowptr[i + 1]; ++q) t = __dadd_rn(t, __dmul_rn(A.val[q], p[A.col[q]]));
        if (JACOBI) t = __dmul_rn(t, dinv[i]);
        vv[i] = t;
        acc[0] = fma(t, rp[i], acc[0]);
    }
    csr_reduce_finalize<1>(acc, FIN_BCGS_D1, ws, st, kc, hist);
}

// s = r - alpha v (built on the fly per column, stored for the own row) ; t = B A s ; sums {s.t, t.t, s.s}
template <bool JACOBI>
__global__ void __launch_bounds__(256) k_bcgs_spmv2(CsrDev A, const double *r, const double *vv, const double *dinv,
                                                    double *s, double *t_out, ReduceWs ws, DevState *st, SolveConsts kc,
                                                    double *hist)
{
    if (st->done) return;
    const double malpha = -st->alpha;
    double acc[3] = {0, 0, 0};
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < A.nrows; i += (int64_t)gridDim.x * blockDim.x)
    {
        const double si = __dadd_rn(__dmul_rn(malpha, vv[i]), r[i]);  // VecWAXPY(S, -alpha, V, R)
        s[i] = si;
        double t = 0.0;
        for (int64_t q = A.rowptr[i]; q < A.rowptr[i + 1]; ++q)
        {
            const int64_t j = A.col[q];
            const double sj = __dadd_rn(__dmul_rn(malpha, vv[j]), r[j]);
            t = __dadd_rn(t, __dmul_rn(A.val[q], sj));
        }
        if (JACOBI) t = __dmul_rn(t, dinv[i]);
        t_out[i] = t;
        acc[0] = fma(si, t, acc[0]);
        acc[1] = fma(t, t, acc[1]);
        acc[2] = fma(si, si, acc[2]);
    }
    csr_reduce_finalize<3>(acc, FIN_BCGS_OMEGA, ws, st, kc, hist);
}

// x = alpha p + omega s + x ; r = s - omega t ; sums {r.r, r.rp}
__global__ void __launch_bounds__(256) k_bcgs_upd(int64_t n, const double *p, const double *s, const double *t,
                                                  const double *rp, double *x, double *r, ReduceWs ws, DevState *st,
                                                  SolveConsts kc, double *hist)
{
    if (st->done) return;
    const double alpha = st->alpha, omega = st->omega;
    double acc[2] = {0, 0};
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    {
        const double si = s[i];
        x[i] = __dadd_rn(__dadd_rn(__dmul_rn(alpha, p[i]), __dmul_rn(omega, si)), x[i]);  // VecAXPBYPCZ(X,alpha,omega,1,P,S)
        const double ri = __dadd_rn(__dmul_rn(-omega, t[i]), si);                          // VecWAXPY(R,-omega,T,S)
        r[i] = ri;
        acc[0] = fma(ri, ri, acc[0]);
        acc[1] = fma(ri, rp[i], acc[1]);
    }
    csr_reduce_finalize<2>(acc, FIN_BCGS_UPD, ws, st, kc, hist);
}

// tail of the "t = s = 0" exit: x += alpha p
__global__ void __launch_bounds__(256) k_bcgs_tail(int64_t n, const double *p, double *x, DevState *st)
{
    if (st->pending != 2) return;
    const double alpha = st->alpha;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        x[i] = __dadd_rn(x[i], __dmul_rn(alpha, p[i]));
}

}  // namespace b200
