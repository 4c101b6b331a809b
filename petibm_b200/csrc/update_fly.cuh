// update_fly.cuh -- k_update2 with the Jacobi diagonal rebuilt on the fly (round-2 candidate, tuning upd_variant = 2).
//
// With -pc_type jacobi the update kernel streams 1/diag next to r and w (32 instead of 24 B/row; DESIGN.md section 10,
// item 4) although the diagonal of the separable operator is a function of six 1-D arrays.  k_update2 issues in 14 % of
// its cycles (profiles/r01_ncu_full_k_spmv2_k_update2_256.txt), so the ~60 extra instructions per 16 bytes should hide
// behind the memory stream.  Same accumulation order as k_jacobi_setup (PCSetUp_Jacobi: MatGetDiagonal + VecReciprocal),
// so z = r / diag is bit-identical to the stored-reciprocal path; everything else is k_update2 verbatim.
#pragma once
#include "kernels.cuh"

namespace b200 {

// 1/diag of the two cells of double2 number idx of the owned range (pad columns: 0, like the stored vector)
__device__ __forceinline__ double2 dinv_pair(const GridDev &g, unsigned int idx)
{
    const unsigned int pxh = (unsigned int)g.px >> 1;
    const unsigned int row = idx / pxh;
    const int i = (int)(2u * (idx - row * pxh));
    const unsigned int kl = row / (unsigned int)g.ny;
    const int j = (int)(row - kl * (unsigned int)g.ny);
    const int kg = g.kz0 + (int)kl;
    const double dyj = g.dy[j], dzk = g.dz[kg];
    const double ayz = __dmul_rn(dyj, dzk);
    const double gya = g.gy[j], gyb = g.gy[j + 1], gza = g.gz[kg], gzb = g.gz[kg + 1];
    const bool wy = g.pery && j == 0;
    const bool wz = (g.gz[0] != 0.0) && kg == 0;
    double out[2];
#pragma unroll
    for (int q = 0; q < 2; ++q)
    {
        const int ii = i + q;
        if (ii >= g.nx)
        {
            out[q] = 0.0;
            continue;
        }
        const double dxi = g.dx[ii];
        const double axz = __dmul_rn(dxi, dzk), axy = __dmul_rn(dxi, dyj);
        const double cxm = __dmul_rn(ayz, g.gx[ii]), cxp = __dmul_rn(ayz, g.gx[ii + 1]);
        const double cym = __dmul_rn(axz, gya), cyp = __dmul_rn(axz, gyb);
        const double czm = __dmul_rn(axy, gza), czp = __dmul_rn(axy, gzb);
        double dg = __dadd_rn(cxm, cxp);
        dg = __dadd_rn(dg, wy ? cyp : cym);
        dg = __dadd_rn(dg, wy ? cym : cyp);
        dg = __dadd_rn(dg, wz ? czp : czm);
        dg = __dadd_rn(dg, wz ? czm : czp);
        dg = -dg;
        out[q] = (dg != 0.0) ? __ddiv_rn(1.0, dg) : 1.0;
    }
    return make_double2(out[0], out[1]);
}

template <bool INIT, bool PADDED, bool PUSH, int U>
__global__ void __launch_bounds__(256) k_update2f(GridDev g, UpdVecs v, int fin_kind, ReduceWs ws, CommDev cm, DevState *st,
                                                  SolveConsts kc, double *hist)
{
    pdl_sync();
    if (st->done) return;
    trace_kernel_start(ws);
    const double ma = INIT ? 0.0 : -st->a;
    const double c = st->c;
    const unsigned int plane2 = (unsigned int)(g.plane >> 1);
    const unsigned int n2 = plane2 * (unsigned int)g.nzl;
    const unsigned int last0 = plane2 * (unsigned int)(g.nzl - 1);
    double2 *const __restrict__ r2 = reinterpret_cast<double2 *>(v.r + g.plane);
    const double2 *const __restrict__ w2 = reinterpret_cast<const double2 *>(v.w + g.plane);
    double2 *const gdn = reinterpret_cast<double2 *>(cm.r_ghost_dn);
    double2 *const gup = reinterpret_cast<double2 *>(cm.r_ghost_up);
    const unsigned int stride = gridDim.x * blockDim.x;
    const unsigned int pxh = (unsigned int)g.px >> 1;
    const unsigned int nx = (unsigned int)g.nx;
    double acc[6] = {0, 0, 0, 0, 0, 0};
    unsigned int i0 = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned int top = n2 - 1u;
    const bool rev = v.reverse != 0;
    while (i0 < n2 && n2 - i0 > (unsigned int)(U - 1) * stride)
    {
        double2 rr[U], wr[U];
        unsigned int idx[U];
#pragma unroll
        for (int u = 0; u < U; ++u)
        {
            const unsigned int j = i0 + u * stride;
            idx[u] = rev ? top - j : j;
            rr[u] = r2[idx[u]];
            if (!INIT) wr[u] = w2[idx[u]];
        }
#pragma unroll
        for (int u = 0; u < U; ++u)  // the diagonal arithmetic of item u runs while the loads of the later items are in flight
            UpdItem<true, INIT, PADDED, PUSH>::run(idx[u], rr[u], wr[u], dinv_pair(g, idx[u]), ma, c, r2, gdn, gup, plane2, last0, pxh,
                                                   nx, acc);
        if (n2 - i0 <= (unsigned int)U * stride)
        {
            i0 = n2;
            break;
        }
        i0 += U * stride;
    }
    for (; i0 < n2; i0 = (n2 - i0 > stride) ? i0 + stride : n2)
    {
        const unsigned int i = rev ? top - i0 : i0;
        double2 rr = r2[i], wr = make_double2(0, 0);
        if (!INIT) wr = w2[i];
        UpdItem<true, INIT, PADDED, PUSH>::run(i, rr, wr, dinv_pair(g, i), ma, c, r2, gdn, gup, plane2, last0, pxh, nx, acc);
    }
    grid_reduce_finalize<6>(acc, fin_kind, ws, cm, st, kc, hist, PUSH);
}

}  // namespace b200
