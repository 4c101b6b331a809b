// mg_kernels.cuh -- geometric multigrid preconditioner for the separable pressure operator (single GPU).
//
// Why: every Poisson configuration PetIBM ships preconditions CG with algebraic multigrid (PETSc GAMG on the CPU,
// AmgX classical AMG on the GPU: examples/*/config/poisson_solver.info, solversPetscOptions.info); without a
// multilevel preconditioner CG needs O(n) iterations per time step.  The operator here
// is not an arbitrary sparse matrix but D (dt I) G of a stretched Cartesian grid, fully described by six 1-D arrays,
// so the hierarchy is GEOMETRIC and matrix-free on every level:
//
//   coarsening    per axis one or two neighbouring cells form a coarse cell, widths add up; WHICH pairs are merged is
//                 decided level by level so that the narrowest cells go first (width-equalising coarsening,
//                 mg_schedule.h: a stretched grid loses its anisotropy on the way down, a uniform one is halved);
//   operator      the same closed form D (dt I) G on the merged cells (rediscretisation; appendix A.1 of SURVEY.md);
//   restriction   sum over the merged cells (the rows are volume-integrated divergences), prolongation = its
//                 transpose (piecewise constant), so the V-cycle is symmetric;
//   smoother      Chebyshev polynomial in D^-1 A on [lambda_max / alpha, lambda_max] with lambda_max = 2 from
//                 Gershgorin (every row of this operator sums to zero), same polynomial before and after the coarse
//                 correction (symmetric), a longer polynomial on the coarsest grid -- a FIXED linear operator, as
//                 PCG requires;
//   null space    the constant is removed from z after the cycle exactly like PETSc does after any PCApply
//                 (MatNullSpaceRemove: the shift computed by finalize_scalars and applied inside k_spmv2).
//
// This is an extension with PETSc's option names (-<name>_pc_type mg, -<name>_pc_mg_levels,
// -<name>_mg_levels_ksp_max_it, -<name>_mg_coarse_ksp_max_it); the reference has no geometric multigrid to be
// bit-compared with -- the checker is an independent numpy/scipy restatement of this algorithm in
// tests/mg_reference.py (assembled level operators), and the converged solutions agree with plain CG.
//
// Kernels: one thread per cell, 7-point rows rebuilt from the 1-D arrays, neighbour values through L1/L2.  A Chebyshev
// step is ONE pass (x + d formed on the fly at the seven points, residual, new d) with ping-pong buffers; the
// prolongation is folded into the first post-smoothing step and the residual into the restriction.
#pragma once
#include <stdint.h>

#include "kernels.cuh"
#include "mg_schedule.h"

namespace b200 {

struct MgLevel
{
    int nx, ny, nz;           // cells of this level
    int px;                   // row pitch of this level's vectors
    long long plane, base;    // plane stride and offset of cell (0,0,0): index = base + i + px*j + plane*k
    int perx, pery, perz;     // periodic axes (index wrap)
    // transition to the next level (null on the coarsest): m?[i] = coarse cell of fine cell i, s?[I] = first fine cell of
    // coarse cell I (n_coarse + 1 entries); one or two fine cells per coarse cell and axis
    const int *mx, *my, *mz;
    const int *sx, *sy, *sz;
    const double *dx, *dy, *dz;  // cell widths
    const double *gx, *gy, *gz;  // face coefficients dt/h, n+1 entries, 0 at walls, wrap value at both ends if periodic
};

// one row of the level operator: slots 0..6 = z-, y-, x-, centre, x+, y+, z+ (ascending column order)
struct MgRow
{
    double c[7];
    long long j[7];            // vector indices of the seven points
    int im, ip, jm, jp, km, kp;  // neighbour coordinates (wrapped / clamped)
};

__device__ __forceinline__ MgRow mg_row(const MgLevel &L, int i, int j, int k)
{
    MgRow r;
    const double dxi = L.dx[i], dyj = L.dy[j], dzk = L.dz[k];
    const double ayz = dyj * dzk, axz = dxi * dzk, axy = dxi * dyj;
    r.c[2] = ayz * L.gx[i];
    r.c[4] = ayz * L.gx[i + 1];
    r.c[1] = axz * L.gy[j];
    r.c[5] = axz * L.gy[j + 1];
    r.c[0] = axy * L.gz[k];
    r.c[6] = axy * L.gz[k + 1];
    r.c[3] = -(((((r.c[2] + r.c[4]) + r.c[1]) + r.c[5]) + r.c[0]) + r.c[6]);
    // neighbours: wrap on periodic axes; at a wall the coefficient is zero and the index stays on the cell itself
    r.im = i > 0 ? i - 1 : (L.perx ? L.nx - 1 : i);
    r.ip = i < L.nx - 1 ? i + 1 : (L.perx ? 0 : i);
    r.jm = j > 0 ? j - 1 : (L.pery ? L.ny - 1 : j);
    r.jp = j < L.ny - 1 ? j + 1 : (L.pery ? 0 : j);
    r.km = k > 0 ? k - 1 : (L.perz ? L.nz - 1 : k);
    r.kp = k < L.nz - 1 ? k + 1 : (L.perz ? 0 : k);
    const long long rowj = L.base + (long long)L.px * j, pk = L.plane * k;
    r.j[3] = rowj + pk + i;
    r.j[2] = rowj + pk + r.im;
    r.j[4] = rowj + pk + r.ip;
    r.j[1] = L.base + (long long)L.px * r.jm + pk + i;
    r.j[5] = L.base + (long long)L.px * r.jp + pk + i;
    r.j[0] = rowj + L.plane * r.km + i;
    r.j[6] = rowj + L.plane * r.kp + i;
    return r;
}

// (A v)(cell); v(q) = value at slot q
template <class V>
__device__ __forceinline__ double mg_apply(const MgRow &r, V v)
{
    double t = r.c[0] * v(0);
#pragma unroll
    for (int q = 1; q < 7; ++q) t += r.c[q] * v(q);
    return t;
}

__device__ __forceinline__ void mg_cell(const MgLevel &L, long long t, int &i, int &j, int &k)
{
    const unsigned int tt = (unsigned int)t;   // a level never has more than 2^31 cells (checked on the host)
    const unsigned int row = tt / (unsigned int)L.nx;
    i = (int)(tt - row * (unsigned int)L.nx);
    k = (int)(row / (unsigned int)L.ny);
    j = (int)(row - (unsigned int)k * (unsigned int)L.ny);
}

// index in the NEXT level's vectors of the coarse cell that holds fine cell (i, j, k)
__device__ __forceinline__ long long mg_coarse_index(const MgLevel &L, const MgLevel &Lc, int i, int j, int k)
{
    return Lc.base + L.mx[i] + (long long)Lc.px * L.my[j] + Lc.plane * L.mz[k];
}

// ---- bodies: (t0, stride) = the calling thread's first cell and the number of cooperating threads, so that the same
// code runs as a grid-stride kernel of its own or as one step of the single-CTA tail below

// first Chebyshev step from a zero guess: d = (1/theta) D^-1 b   (x stays implicit zero)
__device__ __forceinline__ void mg_first_body(const MgLevel &L, const double *b, double *dout, double inv_theta, long long t0,
                                              long long stride)
{
    const long long n = (long long)L.nx * L.ny * L.nz;
    for (long long t = t0; t < n; t += stride)
    {
        int i, j, k;
        mg_cell(L, t, i, j, k);
        const MgRow r = mg_row(L, i, j, k);
        dout[r.j[3]] = r.c[3] != 0.0 ? inv_theta * (b[r.j[3]] / r.c[3]) : 0.0;
    }
}

// one Chebyshev step: the iterate is v = x + d (+ P e_c);  r = b - A v;  d' = c1 d + c2 D^-1 r;
//   x' = v          (d' still to be added by the next pass), or
//   x' = v + d'     when LAST (nothing follows).
// XZERO / DZERO: x / d is implicitly zero (not read).  PROLONG: the coarse-grid correction e_c of the next level is
// added on the fly (piecewise-constant prolongation folded into the first post-smoothing step).
struct MgNoHook
{
    __device__ __forceinline__ void operator()(long long, double) const {}
};

template <bool XZERO, bool DZERO, bool PROLONG, bool LAST, class Hook = MgNoHook>
__device__ __forceinline__ void mg_step_body(const MgLevel &L, const MgLevel &Lc, const double *b, const double *xin,
                                             const double *din, const double *ec, double *xout, double *dout, double c1, double c2,
                                             long long t0, long long stride, Hook hook = Hook())
{
    const long long n = (long long)L.nx * L.ny * L.nz;
    for (long long t = t0; t < n; t += stride)
    {
        int i, j, k;
        mg_cell(L, t, i, j, k);
        const MgRow r = mg_row(L, i, j, k);
        long long cj[7];
        if (PROLONG)
        {
            cj[3] = mg_coarse_index(L, Lc, i, j, k);
            cj[2] = mg_coarse_index(L, Lc, r.im, j, k);
            cj[4] = mg_coarse_index(L, Lc, r.ip, j, k);
            cj[1] = mg_coarse_index(L, Lc, i, r.jm, k);
            cj[5] = mg_coarse_index(L, Lc, i, r.jp, k);
            cj[0] = mg_coarse_index(L, Lc, i, j, r.km);
            cj[6] = mg_coarse_index(L, Lc, i, j, r.kp);
        }
        auto val = [&](int q) {
            double v = 0.0;
            if (!XZERO) v = xin[r.j[q]];
            if (!DZERO) v += din[r.j[q]];
            if (PROLONG) v += ec[cj[q]];
            return v;
        };
        const double vc = val(3);
        const double res = b[r.j[3]] - mg_apply(r, val);
        const double dold = DZERO ? 0.0 : din[r.j[3]];
        const double dnew = r.c[3] != 0.0 ? c1 * dold + c2 * (res / r.c[3]) : 0.0;
        const double xnew = LAST ? vc + dnew : vc;
        xout[r.j[3]] = xnew;
        if (!LAST) dout[r.j[3]] = dnew;
        hook(r.j[3], xnew);
    }
}

// residual + restriction: v = x + d is materialised (xsum), b_c(I,J,K) = sum over the merged fine cells of (b - A v).
// One thread per COARSE cell.
template <bool XZERO>
__device__ __forceinline__ void mg_restrict_body(const MgLevel &L, const MgLevel &Lc, const double *b, const double *xin,
                                                 const double *din, double *xsum, double *bc, long long t0, long long stride)
{
    const long long nc = (long long)Lc.nx * Lc.ny * Lc.nz;
    for (long long t = t0; t < nc; t += stride)
    {
        int I, J, K;
        mg_cell(Lc, t, I, J, K);
        const int i0 = L.sx[I], i1 = L.sx[I + 1] - 1;
        const int j0 = L.sy[J], j1 = L.sy[J + 1] - 1;
        const int k0 = L.sz[K], k1 = L.sz[K + 1] - 1;
        double acc = 0.0;
        for (int k = k0; k <= k1; ++k)
            for (int j = j0; j <= j1; ++j)
                for (int i = i0; i <= i1; ++i)
                {
                    const MgRow r = mg_row(L, i, j, k);
                    auto val = [&](int q) { return XZERO ? din[r.j[q]] : xin[r.j[q]] + din[r.j[q]]; };
                    xsum[r.j[3]] = val(3);
                    acc += b[r.j[3]] - mg_apply(r, val);
                }
        bc[Lc.base + I + (long long)Lc.px * J + Lc.plane * K] = acc;
    }
}

// ---- one kernel per step (levels with enough cells to fill the machine)
__global__ void __launch_bounds__(256) k_mg_cheb_first(MgLevel L, const double *b, double *dout, double inv_theta,
                                                       const DevState *st)
{
    if (st->done) return;
    mg_first_body(L, b, dout, inv_theta, blockIdx.x * (long long)blockDim.x + threadIdx.x, (long long)gridDim.x * blockDim.x);
}

template <bool XZERO, bool DZERO, bool PROLONG, bool LAST>
__global__ void __launch_bounds__(256) k_mg_cheb_step(MgLevel L, MgLevel Lc, const double *b, const double *xin,
                                                      const double *din, const double *ec, double *xout, double *dout,
                                                      double c1, double c2, const DevState *st)
{
    if (st->done) return;
    mg_step_body<XZERO, DZERO, PROLONG, LAST>(L, Lc, b, xin, din, ec, xout, dout, c1, c2,
                                              blockIdx.x * (long long)blockDim.x + threadIdx.x, (long long)gridDim.x * blockDim.x);
}

template <bool XZERO>
__global__ void __launch_bounds__(256) k_mg_restrict(MgLevel L, MgLevel Lc, const double *b, const double *xin, const double *din,
                                                     double *xsum, double *bc, const DevState *st)
{
    if (st->done) return;
    mg_restrict_body<XZERO>(L, Lc, b, xin, din, xsum, bc, blockIdx.x * (long long)blockDim.x + threadIdx.x,
                            (long long)gridDim.x * blockDim.x);
}

// ---- the coarse tail of the cycle as ONE launch.  Below a few thousand cells every step above is a launch that costs
// more than its work (a V(2,2) cycle on 256^3 has ~36 of them under 16^3).  The steps of the levels >= tail_level are
// recorded once, in schedule order (mg_schedule.h: the recording launcher), and interpreted here by a single CTA with a
// block barrier between steps; the bodies are the ones above, so the numbers are the same bit for bit.
__global__ void __launch_bounds__(512) k_mg_tail(const MgLevel *levels, const MgOp *ops, int nops, const DevState *st)
{
    if (st->done) return;
    const long long t0 = threadIdx.x, stride = blockDim.x;
    for (int q = 0; q < nops; ++q)
    {
        const MgOp op = ops[q];
        const MgLevel &L = levels[op.l];
        if (op.kind == 0) mg_first_body(L, op.b, op.dout, op.c2, t0, stride);
        else if (op.kind == 2)
        {
            const MgLevel &Lc = levels[op.l + 1];
            if (op.xzero) mg_restrict_body<true>(L, Lc, op.b, op.xin, op.din, op.xout, op.dout, t0, stride);
            else mg_restrict_body<false>(L, Lc, op.b, op.xin, op.din, op.xout, op.dout, t0, stride);
        }
        else
        {
            const MgLevel &Lc = levels[op.prolong ? op.l + 1 : op.l];
#define B200_MG_TAIL(XZ, DZ, PR, LA) \
    mg_step_body<XZ, DZ, PR, LA>(L, Lc, op.b, op.xin, op.din, op.ec, op.xout, op.dout, op.c1, op.c2, t0, stride)
            if (op.prolong)
            {
                if (op.last) B200_MG_TAIL(false, true, true, true);
                else B200_MG_TAIL(false, true, true, false);
            }
            else if (op.xzero)
            {
                if (op.last) B200_MG_TAIL(true, false, false, true);
                else B200_MG_TAIL(true, false, false, false);
            }
            else
            {
                if (op.last) B200_MG_TAIL(false, false, false, true);
                else B200_MG_TAIL(false, false, false, false);
            }
#undef B200_MG_TAIL
        }
        __syncthreads();  // the next step reads what this one wrote (global memory, same block)
    }
}

// ---- tuning "mg_fuse": the two CG passes around the cycle folded into its first and last fine-level steps
// first step + residual update: r <- r - a w (VecAXPY(R, -a, W)), d = (1/theta) D^-1 r      (k_mg_rupdate + k_mg_cheb_first)
__global__ void __launch_bounds__(256) k_mg_first_rupd(MgLevel L, double *r, const double *w, double *dout, double inv_theta,
                                                       const DevState *st)
{
    if (st->done) return;
    const double ma = -st->a;
    const long long n = (long long)L.nx * L.ny * L.nz;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x)
    {
        int i, j, k;
        mg_cell(L, t, i, j, k);
        const MgRow row = mg_row(L, i, j, k);
        const double rn = __dadd_rn(r[row.j[3]], __dmul_rn(ma, w[row.j[3]]));
        r[row.j[3]] = rn;
        dout[row.j[3]] = row.c[3] != 0.0 ? inv_theta * (rn / row.c[3]) : 0.0;
    }
}

// last step + the six sums of k_mg_zsums for the z it produces (b is the residual r)      (last step + k_mg_zsums)
template <bool PROLONG>
__global__ void __launch_bounds__(256) k_mg_last_sums(MgLevel L, MgLevel Lc, const double *b, const double *xin, const double *din,
                                                      const double *ec, double *xout, double c1, double c2, int fin_kind,
                                                      ReduceWs ws, CommDev cm, DevState *st, SolveConsts kc, double *hist)
{
    if (st->done) return;
    const double c = st->c;
    double acc[6] = {0, 0, 0, 0, 0, 0};
    auto hook = [&](long long idx, double z0) {
        const double rn = b[idx];
        const double d0 = z0 - c;
        acc[0] += z0;
        acc[1] += d0;
        acc[2] = fma(d0, d0, acc[2]);
        acc[3] = fma(d0, rn, acc[3]);
        acc[4] += rn;
        acc[5] = fma(rn, rn, acc[5]);
    };
    const long long t0 = blockIdx.x * (long long)blockDim.x + threadIdx.x, stride = (long long)gridDim.x * blockDim.x;
    if (PROLONG) mg_step_body<false, true, true, true>(L, Lc, b, xin, din, ec, xout, nullptr, c1, c2, t0, stride, hook);
    else mg_step_body<false, false, false, true>(L, Lc, b, xin, din, ec, xout, nullptr, c1, c2, t0, stride, hook);
    grid_reduce_finalize<6>(acc, fin_kind, ws, cm, st, kc, hist, false);
}

// PCG with an explicit z = M^-1 r: the residual update on its own (the sums need z, which the cycle produces next)
__global__ void __launch_bounds__(256) k_mg_rupdate(long long n, double *r, const double *w, const DevState *st)
{
    if (st->done) return;
    const double ma = -st->a;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x)
        r[t] = __dadd_rn(r[t], __dmul_rn(ma, w[t]));  // VecAXPY(R, -a, W)
}

// the sums of k_update2 for z given as a vector: {sum z, sum d, sum d^2, sum d r, sum r, sum r^2}, d = z - c
__global__ void __launch_bounds__(256) k_mg_zsums(MgLevel L, const double *z, const double *r, int fin_kind, ReduceWs ws, CommDev cm,
                                                  DevState *st, SolveConsts kc, double *hist)
{
    if (st->done) return;
    const double c = st->c;
    double acc[6] = {0, 0, 0, 0, 0, 0};
    const long long n = (long long)L.nx * L.ny * L.nz;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x)
    {
        int i, j, k;
        mg_cell(L, t, i, j, k);
        const long long idx = L.base + i + (long long)L.px * j + L.plane * k;
        const double z0 = z[idx], rn = r[idx];
        const double d0 = z0 - c;
        acc[0] += z0;
        acc[1] += d0;
        acc[2] = fma(d0, d0, acc[2]);
        acc[3] = fma(d0, rn, acc[3]);
        acc[4] += rn;
        acc[5] = fma(rn, rn, acc[5]);
    }
    grid_reduce_finalize<6>(acc, fin_kind, ws, cm, st, kc, hist, false);
}

}  // namespace b200
