// spmv5.cuh -- class-0 kernel for CACHE-RESIDENT slabs: the same fused
//     x <- x + a' p' ; p <- z + b p' ; w <- A p ; dpi <- p.w
// as k_spmv2 / k_spmv4, without any shared-memory staging.
//
// On 8 GPUs the 256^3 problem leaves a 256 x 256 x 32 slab per GPU: five vectors of 16.8 MB, all of it resident in the
// 126 MB L2.  There the marching kernels are bound by their own latency chain (a CTA walks 32 planes through a 3-4 stage
// ring, one L2 round trip per step, 7-14 warps per SM: profiles/r01_ncu_full_slab_256x256x32.txt), not by bandwidth.
// This kernel trades bytes for parallelism: one thread owns two x-neighbouring points and a short run of KB planes, reads
// r and p' of its points and of their x/y neighbours straight from global memory (L1/L2 hits: the lines are shared with
// the neighbouring rows of the same CTA), rebuilds p = z + b p' at every point it needs (pointwise, 3 flops) and keeps
// the z column p(k-1), p(k), p(k+1) in registers.  No barriers, no pipeline fill, tens of warps per SM.
// Arithmetic (operand order, no FMA contraction) is identical to k_spmv / k_spmv2 / k_spmv4: bit-exact against the
// assembled D*(dt*G) MatMult.  Non-periodic grids only (wall neighbours are clamped onto a valid point; their
// coefficient is zero).
#pragma once
#include "kernels.cuh"

namespace b200 {

template <int TY, int KB, int MINB, bool JACOBI>
__global__ void __launch_bounds__(32 * TY, MINB)
    k_spmv5(GridDev g, VecSet v, ReduceWs ws, CommDev cm, DevState *st, SolveConsts kc, double *hist, int ghost_store)
{
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int i = blockIdx.x * 64 + 2 * tx, j = blockIdx.y * TY + ty;
    const int k0 = blockIdx.z * KB;
    const int k1 = min(k0 + KB, g.nzl);
    const bool st0 = (j < g.ny) && (i < g.nx);
    const bool st1 = (j < g.ny) && (i + 1 < g.nx);

    // thread-constant coefficients and neighbour offsets (launch-invariant data only)
    double dx0 = 0, dx1 = 0, gxa = 0, gxb = 0, gxc = 0, dyj = 0, gya = 0, gyb = 0;
    long long o_c = 0, o_ym = 0, o_yp = 0, o_xm = 0, o_xp = 0;
    if (st0)
    {
        dyj = g.dy[j];
        gya = g.gy[j];
        gyb = g.gy[j + 1];
        dx0 = g.dx[i];
        gxa = g.gx[i];
        gxb = g.gx[i + 1];
        if (st1)
        {
            dx1 = g.dx[i + 1];
            gxc = g.gx[i + 2];
        }
        const int jm = j > 0 ? j - 1 : 0, jp = j + 1 < g.ny ? j + 1 : g.ny - 1;
        const int im = i > 0 ? i - 1 : 0, ip = i + 2 < g.nx ? i + 2 : g.nx - 1;
        o_c = (long long)j * g.px + i;
        o_ym = (long long)jm * g.px + i;
        o_yp = (long long)jp * g.px + i;
        o_xm = (long long)j * g.px + im;
        o_xp = (long long)j * g.px + ip;
    }
    const double axy0 = __dmul_rn(dx0, dyj), axy1 = __dmul_rn(dx1, dyj);

    // ---- everything above reads only launch-invariant data: from here on the predecessor must be done
    pdl_sync();
    if (st->done) return;
    trace_kernel_start(ws);
    const double shift = st->shift, bcoef = st->b, aprev = st->a;
    const bool xupd = st->pending != 0;
    halo_wait_cta(cm, st->seq, k0 == 0, k1 == g.nzl);  // ghost planes of r from the neighbour GPUs

    auto pval = [&](double r, double d, double pp) -> double {
        double z = r;
        if (JACOBI) z = __dmul_rn(z, d);
        z = __dadd_rn(z, shift);
        return __dadd_rn(z, __dmul_rn(bcoef, pp));
    };
    auto ld2 = [](const double *a, long long o) { return *reinterpret_cast<const double2 *>(a + o); };
    // p at the thread's own pair on storage plane sp (0 .. nzl+1); pp_out receives p' there
    auto own = [&](int sp, double2 &pp_out) -> double2 {
        const long long o = (long long)sp * g.plane + o_c;
        const double2 r = ld2(v.r, o);
        pp_out = ld2(v.p_in, o);
        double2 d = make_double2(0, 0);
        if (JACOBI) d = ld2(v.dinv, o);
        return make_double2(pval(r.x, d.x, pp_out.x), pval(r.y, d.y, pp_out.y));
    };

    double acc0 = 0.0;
    if (st0)
    {
        double2 ppm, ppc, ppn;
        double2 pm = own(k0, ppm);       // plane k0-1 (storage k0)
        double2 pc = own(k0 + 1, ppc);   // plane k0
        if (ghost_store && k0 == 0)
        {
            if (st1) *reinterpret_cast<double2 *>(v.p_out + o_c) = pm;
            else v.p_out[o_c] = pm.x;
        }
        double gz_lo = g.gz[g.kz0 + k0];
#pragma unroll 1
        for (int kk = k0; kk < k1; ++kk)
        {
            const long long base = (long long)(kk + 1) * g.plane;
            double2 pn = own(kk + 2, ppn);  // plane kk+1 (a ghost plane above the last owned one)
            // ---- p on plane kk at the x/y neighbours
            const double2 rym = ld2(v.r, base + o_ym), pym = ld2(v.p_in, base + o_ym);
            const double2 ryp = ld2(v.r, base + o_yp), pyp = ld2(v.p_in, base + o_yp);
            const double rxm = v.r[base + o_xm], pxm = v.p_in[base + o_xm];
            const double rxp = v.r[base + o_xp], pxp = v.p_in[base + o_xp];
            double2 dym = make_double2(0, 0), dyp = make_double2(0, 0);
            double dxm = 0, dxp = 0;
            if (JACOBI)
            {
                dym = ld2(v.dinv, base + o_ym);
                dyp = ld2(v.dinv, base + o_yp);
                dxm = v.dinv[base + o_xm];
                dxp = v.dinv[base + o_xp];
            }
            double2 xv = make_double2(0, 0);
            if (xupd) xv = ld2(v.x, base + o_c);
            const double dz_cur = g.dz[g.kz0 + kk];
            const double gz_hi = g.gz[g.kz0 + kk + 1];
            const double vym0 = pval(rym.x, dym.x, pym.x), vym1 = pval(rym.y, dym.y, pym.y);
            const double vyp0 = pval(ryp.x, dyp.x, pyp.x), vyp1 = pval(ryp.y, dyp.y, pyp.y);
            const double vxm = pval(rxm, dxm, pxm), vxp = pval(rxp, dxp, pxp);
            const double ayz = __dmul_rn(dyj, dz_cur);
            const double cx0 = __dmul_rn(ayz, gxa), cx1 = __dmul_rn(ayz, gxb), cx2 = __dmul_rn(ayz, gxc);
            double wout[2];
#pragma unroll
            for (int q = 0; q < 2; ++q)
            {
                const double dxi = q ? dx1 : dx0;
                const double cxm = q ? cx1 : cx0;
                const double cxp = q ? cx2 : cx1;
                const double axz = __dmul_rn(dxi, dz_cur);
                const double cym = __dmul_rn(axz, gya), cyp = __dmul_rn(axz, gyb);
                const double czm = __dmul_rn(q ? axy1 : axy0, gz_lo);
                const double czp = __dmul_rn(q ? axy1 : axy0, gz_hi);
                const double x0 = q ? pc.y : pc.x;
                const double xm = q ? pc.x : vxm;
                const double xp = q ? vxp : pc.y;
                const double vym = q ? vym1 : vym0, vyp = q ? vyp1 : vyp0;
                const double vzm = q ? pm.y : pm.x, vzp = q ? pn.y : pn.x;
                // diagonal: MatMatMult accumulation over the D row u(i-1),u(i),v(j-1),v(j),w(k-1),w(k)
                double dg = __dadd_rn(cxm, cxp);
                dg = __dadd_rn(dg, cym);
                dg = __dadd_rn(dg, cyp);
                dg = __dadd_rn(dg, czm);
                dg = __dadd_rn(dg, czp);
                dg = -dg;
                // MatMult_SeqAIJ in ascending column order: k-1, j-1, i-1, diag, i+1, j+1, k+1
                double s_ = __dmul_rn(czm, vzm);
                s_ = __dadd_rn(s_, __dmul_rn(cym, vym));
                s_ = __dadd_rn(s_, __dmul_rn(cxm, xm));
                s_ = __dadd_rn(s_, __dmul_rn(dg, x0));
                s_ = __dadd_rn(s_, __dmul_rn(cxp, xp));
                s_ = __dadd_rn(s_, __dmul_rn(cyp, vyp));
                s_ = __dadd_rn(s_, __dmul_rn(czp, vzp));
                wout[q] = s_;
                if (q == 0 || st1) acc0 = fma(x0, s_, acc0);
            }
            // ---- stores on plane kk: w, the new search direction, the deferred VecAXPY(X, a', P')
            if (st1)
            {
                *reinterpret_cast<double2 *>(v.w + base + o_c) = make_double2(wout[0], wout[1]);
                *reinterpret_cast<double2 *>(v.p_out + base + o_c) = pc;
                if (xupd)
                    *reinterpret_cast<double2 *>(v.x + base + o_c) =
                        make_double2(__dadd_rn(xv.x, __dmul_rn(aprev, ppc.x)), __dadd_rn(xv.y, __dmul_rn(aprev, ppc.y)));
            }
            else
            {
                v.w[base + o_c] = wout[0];
                v.p_out[base + o_c] = pc.x;
                if (xupd) v.x[base + o_c] = __dadd_rn(xv.x, __dmul_rn(aprev, ppc.x));
            }
            pm = pc;
            pc = pn;
            ppc = ppn;
            gz_lo = gz_hi;
        }
        // ghost plane above the slab (pc holds plane k1 after the rotation)
        if (ghost_store && k1 == g.nzl)
        {
            const long long o = (long long)(g.nzl + 1) * g.plane + o_c;
            if (st1) *reinterpret_cast<double2 *>(v.p_out + o) = pc;
            else v.p_out[o] = pc.x;
        }
    }
    double acc[1] = {acc0};
    grid_reduce_finalize<1>(acc, FIN_SPMV, ws, cm, st, kc, hist, false);
}

}  // namespace b200
