// spmv2.cuh -- second-generation class-0 kernel: the same fused
//     x <- x + a' p' ; p <- z + b p' ; w <- A p ; dpi <- p.w
// as k_spmv (kernels.cuh), restructured for the B200 memory system and issue budget:
//   * every global load is a cp.async (LDGSTS) into an S-stage shared-memory ring, so the bytes in
//     flight per SM are set by shared memory (S-1 planes x 48 B/thread) and not by registers;
//   * addresses are ONE loop-carried byte offset per thread plus block-uniform bases; the plane loop
//     is unrolled by three so that the 3-plane tile ring and the (k-1, k, k+1) register rotation are
//     static (the first version spent ~60 % of its issue slots re-deriving addresses and moving
//     registers, see profiles/);
//   * dead ghost planes are simply zero-filled storage, periodic-wrap summation order lives in a
//     separate instantiation (PER) so the common non-periodic kernel carries no wrap code;
//   * per-plane 1-D coefficients (dz, gz) are prefetched one plane ahead.
// Arithmetic (operand order, no FMA contraction) is identical to k_spmv: bit-exact against the
// assembled D*(dt*G) MatMult.
#pragma once
#include <type_traits>

#include "kernels.cuh"

namespace b200 {


template <int TXT, int TYT, int S, bool JACOBI, bool APPLY>
struct Spmv2Smem
{
    static constexpr int BX = 2 * TXT;
    static constexpr int SROW = BX + 4;  // [1] left halo, [2..BX+1] tile, [BX+2] right halo
    static constexpr int NARR = (APPLY ? 1 : 3) + (JACOBI ? 1 : 0);  // r [, p, x] [, dinv]
    static constexpr int NHAL = (APPLY ? 1 : 2) + (JACOBI ? 1 : 0);  // r [, p] [, dinv]
    static constexpr unsigned ARR_BYTES = 16u * TYT * TXT;           // one array of one stage
    static constexpr unsigned STAGE_BYTES = ARR_BYTES * NARR;
    static constexpr unsigned HARR_BYTES = 8u * TYT * 2;
    static constexpr unsigned HSTAGE_BYTES = HARR_BYTES * NHAL;
    static constexpr unsigned TILE_BYTES = 8u * TYT * SROW;
    static constexpr unsigned ring_off = 0;
    static constexpr unsigned halo_off = STAGE_BYTES * S;
    static constexpr unsigned tile_off = halo_off + HSTAGE_BYTES * S;
    static constexpr unsigned coef_off = tile_off + 3u * TILE_BYTES;  // dz[nplanes], gz[nplanes] of the chunk
    static constexpr size_t total_fixed = coef_off;
    static size_t total(int kz_chunk) { return total_fixed + 16u * (size_t)(kz_chunk + 2); }
};

template <int TXT, int TYT, int S, int MINB, bool JACOBI, bool APPLY, bool PER>
__global__ void __launch_bounds__(TXT *TYT, MINB)
    k_spmv2(GridDev g, VecSet v, int kz_chunk, ReduceWs ws, CommDev cm, DevState *st, SolveConsts kc, double *hist,
            int ghost_store)
{
    static_assert(TXT == 32, "one warp per tile row");
    using L = Spmv2Smem<TXT, TYT, S, JACOBI, APPLY>;
    constexpr int BX = L::BX, TY = TYT - 2, SROW = L::SROW;
    constexpr unsigned A_R = 0, A_P = L::ARR_BYTES, A_X = 2 * L::ARR_BYTES, A_D = (APPLY ? 1 : 3) * L::ARR_BYTES;
    constexpr unsigned H_R = 0, H_P = L::HARR_BYTES, H_D = (APPLY ? 1 : 2) * L::HARR_BYTES;
    B200_DYNAMIC_SMEM(smem_raw);
    const unsigned int smem_base = smem_u32(smem_raw);

    const int tx = threadIdx.x, ty = threadIdx.y;
    const int i0 = blockIdx.x * BX, j0 = blockIdx.y * TY;
    const int k0 = blockIdx.z * kz_chunk;
    const int k1 = min(k0 + kz_chunk, g.nzl);
    const int nplanes = k1 - k0 + 2;  // t = 0 .. nplanes-1  <->  plane kk = k0-1+t
    const int i = i0 + 2 * tx;
    const int jr = j0 - 1 + ty;
    int jm = jr;
    if (jr < 0) jm = g.pery ? g.ny - 1 : 0;
    else if (jr >= g.ny) jm = g.pery ? jr - g.ny : g.ny - 1;
    if (jm >= g.ny) jm = g.ny - 1;
    const bool vec_ok = (i + 1 < g.nx);
    int c0 = i, c1 = i + 1;
    if (c0 >= g.nx) c0 = g.perx ? min(c0 - g.nx, g.nx - 1) : g.nx - 1;
    if (c1 >= g.nx) c1 = g.perx ? min(c1 - g.nx, g.nx - 1) : g.nx - 1;
    const bool interior_row = (ty >= 1) && (ty <= TY) && (jr < g.ny);
    const bool st0 = interior_row && (i < g.nx);
    const bool st1 = interior_row && (i + 1 < g.nx);
    const bool edgeL = (tx == 0), edgeR = (tx == TXT - 1);
    int ch = 0;
    if (edgeL) ch = (i0 == 0) ? (g.perx ? g.nx - 1 : 0) : i0 - 1;
    if (edgeR)
    {
        ch = i0 + BX;
        if (ch >= g.nx) ch = g.perx ? min(ch - g.nx, g.nx - 1) : g.nx - 1;
    }
    const bool edge = edgeL || edgeR;

    // ---- per-thread invariant addresses
    const long long planeB = g.plane * 8;                       // plane stride in bytes
    const long long rowB = (long long)jm * g.px * 8;
    const long long tB = rowB + (long long)i * 8;               // this thread's double2 inside a plane
    const long long hB = rowB + (long long)ch * 8;              // its x-halo element (edge lanes)
    const long long dB0 = rowB + (long long)c0 * 8 - tB;        // scalar fallbacks relative to tB
    const long long dB1 = rowB + (long long)c1 * 8 - tB;
    const unsigned int ring_thr = smem_base + L::ring_off + 16u * (unsigned)(ty * TXT + tx);
    const unsigned int halo_thr = smem_base + L::halo_off + 8u * (unsigned)(ty * 2 + (edgeR ? 1 : 0));
    const unsigned int tile_thr = smem_base + L::tile_off + 8u * (unsigned)(ty * SROW + 2 + 2 * tx);
    const unsigned int tile_hal = smem_base + L::tile_off + 8u * (unsigned)(ty * SROW + (edgeL ? 1 : BX + 2));
    const char *const br = reinterpret_cast<const char *>(v.r);
    const char *const bp = reinterpret_cast<const char *>(v.p_in);
    const char *const bd = reinterpret_cast<const char *>(v.dinv);
    char *const bx = reinterpret_cast<char *>(v.x);
    char *const bpo = reinterpret_cast<char *>(v.p_out);
    char *const bw = reinterpret_cast<char *>(v.w);

    // thread-constant coefficients
    double dx0 = 0, dx1 = 0, gxa = 0, gxb = 0, gxc = 0, dyj = 0, gya = 0, gyb = 0;
    if (interior_row)
    {
        dyj = g.dy[jr];
        gya = g.gy[jr];
        gyb = g.gy[jr + 1];
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
    const double axy0 = __dmul_rn(dx0, dyj), axy1 = __dmul_rn(dx1, dyj);
    // per-plane 1-D coefficients of the chunk, staged once: coef[t] = dz of plane kk = k0-1+t,
    // coef[kz_chunk+2+t] = plus-face z coefficient of plane kk (t = 1 .. nplanes-2 are the owned planes)
    const unsigned int coef_dz = smem_base + L::coef_off, coef_gz = coef_dz + 8u * (unsigned)(kz_chunk + 2);
    {
        double *cf = reinterpret_cast<double *>(smem_raw + L::coef_off);
        for (int t = ty * TXT + tx + 1; t < nplanes - 1; t += TXT * TYT)
        {
            cf[t] = g.dz[g.kz0 + k0 - 1 + t];
            cf[kz_chunk + 2 + t] = g.gz[g.kz0 + k0 + t];
        }
        if (tx == 0 && ty == 0) cf[kz_chunk + 2] = g.gz[g.kz0 + k0];  // minus face of plane k0
    }

    // ---- everything above reads only launch-invariant data: from here on the predecessor must be done
    pdl_sync();
    double shift = 0.0, bcoef = 0.0, aprev = 0.0;
    bool xupd = false;
    if (!APPLY)
    {
        if (st->done) return;
        trace_kernel_start(ws);
        shift = st->shift;
        bcoef = st->b;
        aprev = st->a;
        xupd = st->pending != 0;
        // ghost planes of r come from the neighbour GPUs: wait for their hand-shake (no-op on one GPU / NCCL transport)
        halo_wait_cta(cm, st->seq, k0 == 0, k1 == g.nzl);
    }

    // storage offset (bytes) of plane kk = k0-1+t: (kk+1)*plane, except the single-GPU periodic wrap
    auto plane_off = [&](int t) -> long long {
        int sp_ = k0 + t;  // kk + 1
        if (g.perz_wrap)
        {
            if (sp_ == 0) sp_ = g.nzl;
            else if (sp_ == g.nzl + 1) sp_ = 1;
        }
        return (long long)sp_ * planeB;
    };

    // ---- producer: one cp.async group per plane; ld_stage = ring offset of the stage being filled
    unsigned int ld_stage = 0, ld_hstage = 0;
    auto issue = [&](int t) {
        if (t < nplanes)
        {
            const long long o = plane_off(t) + tB;
            const unsigned int sd = ring_thr + ld_stage;
            if (vec_ok)
            {
                cp_async16(sd + A_R, br + o);
                if (!APPLY) cp_async16(sd + A_P, bp + o);
                if (JACOBI) cp_async16(sd + A_D, bd + o);
            }
            else
            {
                cp_async8(sd + A_R, br + o + dB0);
                cp_async8(sd + A_R + 8, br + o + dB1);
                if (!APPLY)
                {
                    cp_async8(sd + A_P, bp + o + dB0);
                    cp_async8(sd + A_P + 8, bp + o + dB1);
                }
                if (JACOBI)
                {
                    cp_async8(sd + A_D, bd + o + dB0);
                    cp_async8(sd + A_D + 8, bd + o + dB1);
                }
            }
            if (edge)
            {
                const long long oh = o - tB + hB;
                const unsigned int hd = halo_thr + ld_hstage;
                cp_async8(hd + H_R, br + oh);
                if (!APPLY) cp_async8(hd + H_P, bp + oh);
                if (JACOBI) cp_async8(hd + H_D, bd + oh);
            }
            if (!APPLY && xupd && t >= 1 && t < nplanes - 1)
            {
                const long long xo = (long long)(k0 + t) * planeB + tB;
                if (st1) cp_async16(sd + A_X, bx + xo);
                else if (st0) cp_async8(sd + A_X, bx + xo);
            }
        }
        cp_async_commit();
        ld_stage += L::STAGE_BYTES;
        ld_hstage += L::HSTAGE_BYTES;
        if (ld_stage == L::STAGE_BYTES * S)
        {
            ld_stage = 0;
            ld_hstage = 0;
        }
    };

    double2 pq0 = make_double2(0, 0), pq1 = make_double2(0, 0), pq2 = make_double2(0, 0);
    double czm0 = 0, czm1 = 0, acc0 = 0;
    unsigned int cs_stage = 0, cs_hstage = 0;
    long long so = (long long)k0 * planeB + tB;  // store offset of plane kk (= (kk+1)*plane + thread), t = 0

    // one plane: PH = t mod 3 (static); pnew/pcen/pmin = registers of planes kk, kk-1, kk-2
    auto step = [&](int t, auto ph, double2 &pnew, const double2 &pcen, const double2 &pmin) {
        constexpr int PH = decltype(ph)::value;
        constexpr unsigned TB_NEW = PH * L::TILE_BYTES, TB_CEN = ((PH + 2) % 3) * L::TILE_BYTES;
        issue(t + S - 1);
        const bool own = (t >= 1) && (t < nplanes - 1);
        cp_async_wait<S - 1>();
        // ---- build p on plane kk
        const unsigned int rs = ring_thr + cs_stage;
        const double2 rv = lds128(rs + A_R);
        double2 pv = make_double2(0, 0);
        if (!APPLY) pv = lds128(rs + A_P);
        double z0 = rv.x, z1 = rv.y;
        if (JACOBI)
        {
            const double2 dv = lds128(rs + A_D);
            z0 = __dmul_rn(z0, dv.x);
            z1 = __dmul_rn(z1, dv.y);
        }
        if (!APPLY)
        {
            z0 = __dadd_rn(z0, shift);
            z1 = __dadd_rn(z1, shift);
            pnew.x = __dadd_rn(z0, __dmul_rn(bcoef, pv.x));
            pnew.y = __dadd_rn(z1, __dmul_rn(bcoef, pv.y));
        }
        else
        {
            pnew.x = z0;
            pnew.y = z1;
        }
        sts128(tile_thr + TB_NEW, pnew);
        if (edge)
        {
            const unsigned int hs = halo_thr + cs_hstage;
            double zh = lds64(hs + H_R);
            if (JACOBI) zh = __dmul_rn(zh, lds64(hs + H_D));
            if (!APPLY)
            {
                zh = __dadd_rn(zh, shift);
                zh = __dadd_rn(zh, __dmul_rn(bcoef, lds64(hs + H_P)));
            }
            sts64(tile_hal + TB_NEW, zh);
        }
        if (!APPLY)
        {
            // ---- deferred VecAXPY(X, a', P') on the owned planes
            if (xupd && own && st0)
            {
                const double2 xv = lds128(rs + A_X);
                double2 xn;
                xn.x = __dadd_rn(xv.x, __dmul_rn(aprev, pv.x));
                xn.y = __dadd_rn(xv.y, __dmul_rn(aprev, pv.y));
                if (st1) *reinterpret_cast<double2 *>(bx + so) = xn;
                else *reinterpret_cast<double *>(bx + so) = xn.x;
            }
            // ---- the new search direction (ghost planes too when a neighbour GPU exists)
            if (own || (ghost_store && (k0 + t == 0 || k0 + t == g.nzl + 1)))
            {
                if (st1) *reinterpret_cast<double2 *>(bpo + so) = pnew;
                else if (st0) *reinterpret_cast<double *>(bpo + so) = pnew.x;
            }
        }
        __syncthreads();
        // ---- w on plane k = kk-1 (needs p on kk-2, kk-1, kk)
        if (t >= 2)
        {
            if (st0)
            {
                const unsigned int tc = tile_thr + TB_CEN;
                const double xm0 = lds64(tc - 8);
                const double xp1 = lds64(tc + 16);
                const double2 ym = lds128(tc - 8u * SROW);
                const double2 yp = lds128(tc + 8u * SROW);
                const double dz_cur = lds64(coef_dz + 8u * (unsigned)(t - 1));
                const double gz_cur = lds64(coef_gz + 8u * (unsigned)(t - 1));
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
                    const double czm = q ? czm1 : czm0;
                    const double czp = __dmul_rn(q ? axy1 : axy0, gz_cur);
                    const double x0 = q ? pcen.y : pcen.x;
                    const double xm = q ? pcen.x : xm0;
                    const double xp = q ? xp1 : pcen.y;
                    const double vym = q ? ym.y : ym.x, vyp = q ? yp.y : yp.x;
                    const double vzm = q ? pmin.y : pmin.x, vzp = q ? pnew.y : pnew.x;
                    double tsum = 0.0;
                    bool fast = true;
                    if (PER)
                    {
                        const int k = k0 + t - 2;
                        const int iq = i + q;
                        const bool wrapx = g.perx && (iq == 0 || iq == g.nx - 1);
                        const bool wrapy_lo = g.pery && jr == 0, wrapy_hi = g.pery && jr == g.ny - 1;
                        const bool wrapz_lo = g.wrapz_lo && k == 0, wrapz_hi = g.wrapz_hi && k == g.nzl - 1;
                        fast = !(wrapx || wrapy_lo || wrapy_hi || wrapz_lo || wrapz_hi);
                        if (!fast)
                        {
                            // periodic wrap rows: the wrapped neighbour changes its place in the sorted row
                            const bool xlo = g.perx && iq == 0, xhi = g.perx && iq == g.nx - 1;
                            double dg = __dadd_rn(cxm, cxp);  // first pair: commutative
                            dg = __dadd_rn(dg, wrapy_lo ? cyp : cym);
                            dg = __dadd_rn(dg, wrapy_lo ? cym : cyp);
                            dg = __dadd_rn(dg, wrapz_lo ? czp : czm);
                            dg = __dadd_rn(dg, wrapz_lo ? czm : czp);
                            dg = -dg;
                            const double tzm = __dmul_rn(czm, vzm), tzp = __dmul_rn(czp, vzp);
                            const double tym = __dmul_rn(cym, vym), typ = __dmul_rn(cyp, vyp);
                            const double txm = __dmul_rn(cxm, xm), txp = __dmul_rn(cxp, xp);
                            const double td = __dmul_rn(dg, x0);
                            // sorted columns: zp(w) < zm < yp(w) < ym < xp(w) < xm < d < xp < xm(w) < yp < ym(w) < zp < zm(w)
                            double s_ = 0.0;
                            if (wrapz_hi) s_ = __dadd_rn(s_, tzp);
                            if (!wrapz_lo) s_ = __dadd_rn(s_, tzm);
                            if (wrapy_hi) s_ = __dadd_rn(s_, typ);
                            if (!wrapy_lo) s_ = __dadd_rn(s_, tym);
                            if (xhi) s_ = __dadd_rn(s_, txp);
                            if (!xlo) s_ = __dadd_rn(s_, txm);
                            s_ = __dadd_rn(s_, td);
                            if (!xhi) s_ = __dadd_rn(s_, txp);
                            if (xlo) s_ = __dadd_rn(s_, txm);
                            if (!wrapy_hi) s_ = __dadd_rn(s_, typ);
                            if (wrapy_lo) s_ = __dadd_rn(s_, tym);
                            if (!wrapz_hi) s_ = __dadd_rn(s_, tzp);
                            if (wrapz_lo) s_ = __dadd_rn(s_, tzm);
                            tsum = s_;
                        }
                    }
                    if (fast)
                    {
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
                        tsum = s_;
                    }
                    wout[q] = tsum;
                    if (q == 0 || st1) acc0 = fma(x0, tsum, acc0);
                    if (q) czm1 = czp;
                    else czm0 = czp;
                }
                // storage plane of k = kk-1 is kk: one plane below the store offset of plane kk
                if (st1) *reinterpret_cast<double2 *>(bw + so - planeB) = make_double2(wout[0], wout[1]);
                else *reinterpret_cast<double *>(bw + so - planeB) = wout[0];
            }
        }
        else if (t == 1)
        {
            // entering the chunk: minus-face z coefficient of plane k0
            const double gza = lds64(coef_gz);
            czm0 = __dmul_rn(axy0, gza);
            czm1 = __dmul_rn(axy1, gza);
        }
        so += planeB;
        cs_stage += L::STAGE_BYTES;
        cs_hstage += L::HSTAGE_BYTES;
        if (cs_stage == L::STAGE_BYTES * S)
        {
            cs_stage = 0;
            cs_hstage = 0;
        }
    };

#pragma unroll
    for (int t = 0; t < S - 1; ++t) issue(t);

#pragma unroll 1
    for (int t = 0; t < nplanes; t += 3)
    {
        step(t, std::integral_constant<int, 0>{}, pq0, pq2, pq1);
        if (t + 1 < nplanes) step(t + 1, std::integral_constant<int, 1>{}, pq1, pq0, pq2);
        if (t + 2 < nplanes) step(t + 2, std::integral_constant<int, 2>{}, pq2, pq1, pq0);
    }
    cp_async_wait<0>();
    double acc[1] = {acc0};
    if (!APPLY) grid_reduce_finalize<1>(acc, FIN_SPMV, ws, cm, st, kc, hist, false);
}

}  // namespace b200
