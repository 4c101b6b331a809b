// spmv4.cuh -- third-generation class-0 kernel: the same fused
//     x <- x + a' p' ; p <- z + b p' ; w <- A p ; dpi <- p.w
// as k_spmv2 (spmv2.cuh), rebuilt around the Blackwell copy engine:
//   * TMA: per plane ONE elected thread of a producer warp issues one cp.async.bulk.tensor 3-D box per input array
//     (r, p' [, 1/diag] with their x/y halo: (64+4) x (TY+2) x 1; x without halo) into an S-stage shared-memory ring.
//     The box is clipped by the tensor map (extent nx x ny x (nzl+2)): everything outside the grid -- wall halos, the
//     pad columns of the row pitch, ragged last tiles -- arrives as zeros, so there are no halo lanes, no per-thread
//     copy instructions, no address arithmetic and no ragged-edge predicates on the load side;
//   * mbarrier pipeline: a `full` barrier per stage (expect-tx bytes of the boxes) and an `empty` barrier (one arrival
//     per compute warp).  There is NO CTA-wide barrier in the plane loop: a warp waits for the bytes of its plane,
//     computes, releases the stage;
//   * no shared p tile: p = z + b p' is pointwise, so a thread rebuilds p of its four x/y neighbours from the staged
//     r / p' instead of publishing p through shared memory and synchronising (3 extra dependent flops per neighbour);
//   * the row sum of plane k is split where MatMult's column order allows it: everything up to the (k, j+1) term is
//     formed when plane k is staged, the last term (k+1) is added one step later -- one stage is live per step.
// Arithmetic (operand order, no FMA contraction) is identical to k_spmv / k_spmv2: bit-exact against the assembled
// D*(dt*G) MatMult.  Non-periodic grids only (periodic wrap rows re-sort their columns: k_spmv2<PER> keeps those).
#pragma once
#include "kernels.cuh"

namespace b200 {

template <int TY, int S, bool JACOBI>
struct Spmv4Smem
{
    static constexpr int BX = 64, BW = BX + 4, BH = TY + 2;
    static constexpr unsigned HBOX_BYTES = 8u * BW * BH;                      // bytes one halo box transfers
    static constexpr unsigned HBOX_PITCH = (HBOX_BYTES + 127u) & ~127u;       // TMA destinations are 128-byte aligned
    static constexpr unsigned XBOX_BYTES = 8u * BX * TY;
    static constexpr int NH = 2 + (JACOBI ? 1 : 0);                           // r, p' [, dinv]
    static constexpr unsigned X_OFF = NH * HBOX_PITCH;
    static constexpr unsigned STAGE_BYTES = X_OFF + XBOX_BYTES;
    static constexpr unsigned ring_off = 0;
    static constexpr unsigned bar_off = STAGE_BYTES * S;                      // full[S], empty[S]
    static constexpr unsigned coef_off = bar_off + 16u * S;                   // dz[kz+2], gz[kz+2] of the chunk
    static size_t total(int table_planes) { return 128u + coef_off + 16u * (size_t)(table_planes + 2); }  // BAL: the whole slab
};

struct Spmv4Maps
{
    TmaMap r, p, x, d;  // halo boxes of r, p_in, dinv; plain box of x
};

// BAL = false: one CTA per (x/y tile, z chunk of kz_chunk planes), grid = (x tiles, y tiles, chunks).
// BAL = true ("balanced split"): a 1-D grid of resident CTAs; the (tile, plane) space is linearised tile-major and cut into
// equal ranges of kz_chunk planes, so a CTA finishes the tail of one tile column and continues with the head of the next.
// The producer keeps the ring full across the switch (no drain / fill bubble between work items, no partial last wave);
// every segment costs its two halo planes again.  MEASURED: 40 % slower than the regular grid (206 against 147 us at 256^3) --
// in the regular grid neighbouring tiles march through the same planes at the same time, so their shared halo rows (59 % of
// the r / p' bytes of a 64 x 4 tile) are L2 hits; here every CTA is at a different plane and they come from DRAM again.
// Kept as tile ids 46 / 47 for the record.
template <int TY, int S, int MINB, bool JACOBI, bool BAL>
__global__ void __launch_bounds__(32 * (TY + 1), MINB)
    k_spmv4(const __grid_constant__ Spmv4Maps maps, GridDev g, VecSet v, int kz_chunk, ReduceWs ws, CommDev cm, DevState *st,
            SolveConsts kc, double *hist, int ghost_store)
{
    using L = Spmv4Smem<TY, S, JACOBI>;
    constexpr int BX = L::BX, BW = L::BW;
    constexpr unsigned A_R = 0, A_P = L::HBOX_PITCH, A_D = 2 * L::HBOX_PITCH, A_X = L::X_OFF;
    B200_DYNAMIC_SMEM(smem_raw);
    const unsigned int smem_base = (smem_u32(smem_raw) + 127u) & ~127u;
    const unsigned int bar_full = smem_base + L::bar_off, bar_empty = bar_full + 8u * S;

    const int tx = threadIdx.x, ty = threadIdx.y;
    const bool producer = (ty == TY);
    const int ntx = (g.nx + BX - 1) / BX;
    // the range of this CTA in the linearised (tile, plane) space, or its single (tile, chunk)
    long long lin0 = 0, lin1 = 0;
    int tb0;  // plane whose table index is 1: the coefficient table holds planes tb0-1 .. tb0+tlen-2
    int tlen;
    if (BAL)
    {
        const long long total = (long long)ntx * ((g.ny + TY - 1) / TY) * g.nzl;
        lin0 = min((long long)blockIdx.x * kz_chunk, total);
        lin1 = min(lin0 + kz_chunk, total);
        tb0 = 0;
        tlen = g.nzl + 2;
    }
    else
    {
        tb0 = blockIdx.z * kz_chunk;
        tlen = min(kz_chunk, g.nzl - tb0) + 2;
    }

    if (tx == 0 && ty == 0)
    {
        for (int s = 0; s < S; ++s)
        {
            mbar_init(bar_full + 8u * s, 1);     // the producer's arrive.expect_tx
            mbar_init(bar_empty + 8u * s, TY);   // one arrival per compute warp
        }
        mbar_init_fence();
    }
    if (producer && tx == 0)
    {
        tma_prefetch_desc(&maps.r);
        tma_prefetch_desc(&maps.p);
        tma_prefetch_desc(&maps.x);
        if (JACOBI) tma_prefetch_desc(&maps.d);
    }
    // per-plane 1-D coefficients: cdz[u] = dz of plane kk = tb0-1+u (owned planes only), cgz[u] = plus-face z coefficient of
    // plane kk (u = 0: the minus face of plane tb0)
    const unsigned int coef_dz = smem_base + L::coef_off, coef_gz = coef_dz + 8u * (unsigned)tlen;
    for (int u = ty * 32 + tx; u < tlen - 1; u += 32 * (TY + 1))
    {
        if (u >= 1) sts64(coef_dz + 8u * (unsigned)u, g.dz[g.kz0 + tb0 - 1 + u]);
        sts64(coef_gz + 8u * (unsigned)u, g.gz[g.kz0 + tb0 + u]);
    }
    // thread-constant coefficients of a tile column (launch-invariant data); those of the first segment are loaded here,
    // before the predecessor has finished
    double dx0 = 0, dx1 = 0, gxa = 0, gxb = 0, gxc = 0, dyj = 0, gya = 0, gyb = 0;
    auto load_consts = [&](int ti0, int tj0) {
        const int i = ti0 + 2 * tx, j = tj0 + ty;
        dx0 = dx1 = gxa = gxb = gxc = dyj = gya = gyb = 0.0;
        if (!producer && j < g.ny)
        {
            dyj = g.dy[j];
            gya = g.gy[j];
            gyb = g.gy[j + 1];
            if (i < g.nx)
            {
                dx0 = g.dx[i];
                gxa = g.gx[i];
                gxb = g.gx[i + 1];
            }
            if (i + 1 < g.nx)
            {
                dx1 = g.dx[i + 1];
                gxc = g.gx[i + 2];
            }
        }
    };
    {
        const long long tile0 = BAL ? lin0 / g.nzl : 0;
        load_consts(BAL ? (int)(tile0 % ntx) * BX : (int)blockIdx.x * BX, BAL ? (int)(tile0 / ntx) * TY : (int)blockIdx.y * TY);
    }
    __syncthreads();  // barriers initialised, coefficient table staged

    // ---- everything above reads only launch-invariant data: from here on the predecessor must be done
    pdl_sync();
    if (st->done) return;
    trace_kernel_start(ws);
    const double shift = st->shift, bcoef = st->b, aprev = st->a;
    const bool xupd = st->pending != 0;

    // segment iterator (identical in every thread): tile column, first and one-past-last owned plane
    long long lin = lin0;
    bool first = true;
    auto next_segment = [&](int &i0, int &j0, int &k0, int &k1) -> bool {
        if (BAL)
        {
            if (lin >= lin1) return false;
            const long long tile = lin / g.nzl;
            k0 = (int)(lin - tile * g.nzl);
            k1 = (int)min((long long)g.nzl, (long long)k0 + (lin1 - lin));
            i0 = (int)(tile % ntx) * BX;
            j0 = (int)(tile / ntx) * TY;
            lin += k1 - k0;
            return true;
        }
        if (!first) return false;
        first = false;
        i0 = blockIdx.x * BX;
        j0 = blockIdx.y * TY;
        k0 = blockIdx.z * kz_chunk;
        k1 = min(k0 + kz_chunk, g.nzl);
        return true;
    };

    double acc0 = 0.0;
    int i0, j0, k0, k1;
    if (producer)
    {
        if (tx == 0)
        {
            const unsigned long long seq = st->seq;
            unsigned int stage = 0, par = 1;  // parity to wait for on `empty`: the previous use of the stage
            long long issued = 0;
            while (next_segment(i0, j0, k0, k1))
            {
                const int nplanes = k1 - k0 + 2;  // t = 0 .. nplanes-1  <->  plane kk = k0-1+t, storage plane k0+t
                for (int t = 0; t < nplanes; ++t, ++issued)
                {
                    // a ghost plane of r filled by a neighbour GPU: its hand-shake first (only this thread reads it, through TMA)
                    if (t == 0 && k0 == 0) halo_wait_thread(cm, seq, true, false);
                    if (t == nplanes - 1 && k1 == g.nzl) halo_wait_thread(cm, seq, false, true);
                    if (cm.mode == 1 && (t == 0 || t == nplanes - 1)) proxy_async_fence();
                    if (issued >= S) mbar_wait(bar_empty + 8u * stage, par);
                    const bool own = (t >= 1) && (t < nplanes - 1);
                    const bool wantx = own && xupd;
                    int sp = k0 + t;  // storage plane of kk = k0-1+t
                    if (g.perz_wrap)
                    {
                        if (sp == 0) sp = g.nzl;
                        else if (sp == g.nzl + 1) sp = 1;
                    }
                    const unsigned int fb = bar_full + 8u * stage;
                    const unsigned int sb = smem_base + L::ring_off + stage * L::STAGE_BYTES;
                    mbar_arrive_expect_tx(fb, L::NH * L::HBOX_BYTES + (wantx ? L::XBOX_BYTES : 0u));
                    tma_load_3d(sb + A_R, &maps.r, fb, i0 - 2, j0 - 1, sp);
                    tma_load_3d(sb + A_P, &maps.p, fb, i0 - 2, j0 - 1, sp);
                    if (JACOBI) tma_load_3d(sb + A_D, &maps.d, fb, i0 - 2, j0 - 1, sp);
                    if (wantx) tma_load_3d(sb + A_X, &maps.x, fb, i0, j0, k0 + t);
                    if (++stage == S)
                    {
                        stage = 0;
                        par ^= 1u;
                    }
                }
            }
        }
    }
    else
    {
        const long long planeB = g.plane * 8;
        char *const bx = reinterpret_cast<char *>(v.x);
        char *const bpo = reinterpret_cast<char *>(v.p_out);
        char *const bw = reinterpret_cast<char *>(v.w);
        const unsigned int own_off = 8u * (unsigned)((ty + 1) * BW + 2 + 2 * tx);
        const unsigned int x_off = A_X + 8u * (unsigned)(ty * BX + 2 * tx);

        // p = (B r + shift) + b p' at one point
        auto pval = [&](double r, double d, double pp) -> double {
            double z = r;
            if (JACOBI) z = __dmul_rn(z, d);
            z = __dadd_rn(z, shift);
            return __dadd_rn(z, __dmul_rn(bcoef, pp));
        };

        unsigned int stage = 0, par = 0;
        int nseg = 0;
        while (next_segment(i0, j0, k0, k1))
        {
            const int nplanes = k1 - k0 + 2;
            const int i = i0 + 2 * tx, j = j0 + ty;
            const bool st0 = (j < g.ny) && (i < g.nx);
            const bool st1 = (j < g.ny) && (i + 1 < g.nx);
            if (nseg++ > 0) load_consts(i0, j0);  // the first segment's were loaded before the dependency wait
            const double axy0 = __dmul_rn(dx0, dyj), axy1 = __dmul_rn(dx1, dyj);
            long long so = (long long)k0 * planeB + ((long long)j * g.px + i) * 8;  // store offset of plane kk (storage k0+t), t = 0
            const unsigned int tu = (unsigned)(k0 - tb0);  // table index of t = 0

            double2 pcen = make_double2(0, 0);                    // p on plane kk-1
            double part0 = 0, part1 = 0;                          // row sums of plane kk-1 up to the (j+1) term
            double czp0 = 0, czp1 = 0;                            // plus-face z coefficients of plane kk-1
#pragma unroll 1
            for (int t = 0; t < nplanes; ++t)
            {
                const unsigned int sb = smem_base + L::ring_off + stage * L::STAGE_BYTES;
                mbar_wait(bar_full + 8u * stage, par);
                const bool own = (t >= 1) && (t < nplanes - 1);
                const double2 rc = lds128(sb + A_R + own_off);
                const double2 pc = lds128(sb + A_P + own_off);
                double2 dc = make_double2(0, 0);
                if (JACOBI) dc = lds128(sb + A_D + own_off);
                double2 pnew;
                pnew.x = pval(rc.x, dc.x, pc.x);
                pnew.y = pval(rc.y, dc.y, pc.y);
                // ---- finish plane kk-1: the (k+1) term closes the row sum
                if (t >= 2 && st0)
                {
                    const double w0 = __dadd_rn(part0, __dmul_rn(czp0, pnew.x));
                    const double w1 = __dadd_rn(part1, __dmul_rn(czp1, pnew.y));
                    acc0 = fma(pcen.x, w0, acc0);
                    if (st1)
                    {
                        acc0 = fma(pcen.y, w1, acc0);
                        *reinterpret_cast<double2 *>(bw + so - planeB) = make_double2(w0, w1);
                    }
                    else *reinterpret_cast<double *>(bw + so - planeB) = w0;
                }
                if (own)
                {
                    // ---- p at the four x/y neighbours of the thread's two points, rebuilt from the staged r / p'
                    const double2 rym = lds128(sb + A_R + own_off - 8u * BW), pym = lds128(sb + A_P + own_off - 8u * BW);
                    const double2 ryp = lds128(sb + A_R + own_off + 8u * BW), pyp = lds128(sb + A_P + own_off + 8u * BW);
                    const double rxm = lds64(sb + A_R + own_off - 8u), pxm = lds64(sb + A_P + own_off - 8u);
                    const double rxp = lds64(sb + A_R + own_off + 16u), pxp = lds64(sb + A_P + own_off + 16u);
                    double2 dym = make_double2(0, 0), dyp = make_double2(0, 0);
                    double dxm = 0, dxp = 0;
                    if (JACOBI)
                    {
                        dym = lds128(sb + A_D + own_off - 8u * BW);
                        dyp = lds128(sb + A_D + own_off + 8u * BW);
                        dxm = lds64(sb + A_D + own_off - 8u);
                        dxp = lds64(sb + A_D + own_off + 16u);
                    }
                    double2 xv = make_double2(0, 0);
                    if (xupd) xv = lds128(sb + x_off);
                    const double dz_cur = lds64(coef_dz + 8u * (tu + (unsigned)t));
                    const double gz_lo = lds64(coef_gz + 8u * (tu + (unsigned)(t - 1)));
                    const double gz_hi = lds64(coef_gz + 8u * (tu + (unsigned)t));
                    // the stage is in registers: hand it back to the producer
                    __syncwarp();
                    if (tx == 0) mbar_arrive(bar_empty + 8u * stage);

                    if (st0)
                    {
                        const double vym0 = pval(rym.x, dym.x, pym.x), vym1 = pval(rym.y, dym.y, pym.y);
                        const double vyp0 = pval(ryp.x, dyp.x, pyp.x), vyp1 = pval(ryp.y, dyp.y, pyp.y);
                        const double vxm = pval(rxm, dxm, pxm), vxp = pval(rxp, dxp, pxp);
                        const double ayz = __dmul_rn(dyj, dz_cur);
                        const double cx0 = __dmul_rn(ayz, gxa), cx1 = __dmul_rn(ayz, gxb), cx2 = __dmul_rn(ayz, gxc);
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
                            const double x0 = q ? pnew.y : pnew.x;
                            const double xm = q ? pnew.x : vxm;
                            const double xp = q ? vxp : pnew.y;
                            const double vym = q ? vym1 : vym0, vyp = q ? vyp1 : vyp0;
                            const double vzm = q ? pcen.y : pcen.x;
                            // diagonal: MatMatMult accumulation over the D row u(i-1),u(i),v(j-1),v(j),w(k-1),w(k)
                            double dg = __dadd_rn(cxm, cxp);
                            dg = __dadd_rn(dg, cym);
                            dg = __dadd_rn(dg, cyp);
                            dg = __dadd_rn(dg, czm);
                            dg = __dadd_rn(dg, czp);
                            dg = -dg;
                            // MatMult_SeqAIJ in ascending column order: k-1, j-1, i-1, diag, i+1, j+1 (k+1 follows next step)
                            double s_ = __dmul_rn(czm, vzm);
                            s_ = __dadd_rn(s_, __dmul_rn(cym, vym));
                            s_ = __dadd_rn(s_, __dmul_rn(cxm, xm));
                            s_ = __dadd_rn(s_, __dmul_rn(dg, x0));
                            s_ = __dadd_rn(s_, __dmul_rn(cxp, xp));
                            s_ = __dadd_rn(s_, __dmul_rn(cyp, vyp));
                            if (q) { part1 = s_; czp1 = czp; }
                            else { part0 = s_; czp0 = czp; }
                        }
                        // ---- deferred VecAXPY(X, a', P') and the new search direction on the owned plane
                        if (xupd)
                        {
                            double2 xn;
                            xn.x = __dadd_rn(xv.x, __dmul_rn(aprev, pc.x));
                            xn.y = __dadd_rn(xv.y, __dmul_rn(aprev, pc.y));
                            if (st1) *reinterpret_cast<double2 *>(bx + so) = xn;
                            else *reinterpret_cast<double *>(bx + so) = xn.x;
                        }
                        if (st1) *reinterpret_cast<double2 *>(bpo + so) = pnew;
                        else *reinterpret_cast<double *>(bpo + so) = pnew.x;
                    }
                }
                else
                {
                    __syncwarp();
                    if (tx == 0) mbar_arrive(bar_empty + 8u * stage);
                    // ghost planes of the new search direction when a neighbour GPU exists
                    if (ghost_store && (k0 + t == 0 || k0 + t == g.nzl + 1))
                    {
                        if (st1) *reinterpret_cast<double2 *>(bpo + so) = pnew;
                        else if (st0) *reinterpret_cast<double *>(bpo + so) = pnew.x;
                    }
                }
                pcen = pnew;
                so += planeB;
                if (++stage == S)
                {
                    stage = 0;
                    par ^= 1u;
                }
            }
        }
    }
    double acc[1] = {acc0};
    grid_reduce_finalize<1>(acc, FIN_SPMV, ws, cm, st, kc, hist, false);
}

}  // namespace b200
