// sep_tile.cuh -- tiled, plane-marching kernels for the line-coefficient operator of sep_kernels.cuh.
//
// The row-per-thread kernels (k_sep_*) spend ~350 warp instructions per warp-row: two integer divisions for (i, j, k), a
// branch per stencil term and one gather per neighbour (two for BiCGStab's s = r - alpha v) -- issue- and latency-bound
// at 2.0-2.3 TB/s (profiles/r02_velocity_bench.log).  Here a CTA owns a (32 XR) x 8 tile of one field's (x, y) plane and
// marches it through a chunk of z: the operand of a cell is built ONCE per plane (registers carry it across k - 1, k,
// k + 1; a double-buffered shared plane with a one-cell rim hands it to the x/y neighbours), the x/y coefficients live
// in registers for the whole march, and nothing is divided per row.  One __syncthreads per plane.
//
// Everything a thread reads from HBM goes through a PRIVATE cp.async queue in shared memory (R stages, one plane each): the
// operand arrays of its own cells and of the rim cells it is responsible for, the diagonal, and whatever the result needs
// (1/diag, r', x).  A slot is written and read by the same thread only, so the queue needs no barrier -- just
// cp.async.wait_group -- and R - 2 planes are in flight behind the one being consumed without costing a register.  (The first
// version loaded plane k + 1 into registers at the top of step k: 18-20 warp-stalls on the long scoreboard per issued
// instruction, 2 640 iterations/s against 2 733 for the row-per-thread kernels, profiles/r02_velocity_bench.log.)
//
// Bit-identity with MatMult_SeqAIJ on the assembled matrix is kept exactly as in sep_row: terms in ascending column order
// (z-, y-, x-, diagonal, x+, y+, z+, remainder), no FMA contraction, a zero coefficient = "no entry" = nothing added.  A cell
// with a WRAPPED periodic neighbour has a different column order; those cells (a surface) and the rows behind the stencil
// blocks (IBPM's Lagrangian rows) go through sep_row itself, inside the same kernel, so one launch still carries the
// whole product and its reduction.
//
// Reference: the velocity solve vSolver->solve (navierstokes.cpp:524-537) on A = I/dt - c nu L (createlaplacian.cpp:134-159)
// and IBPM's modified Poisson solve (ibpm.cpp:164-194); PETSc's MatMult_SeqAIJ + VecDot/VecWAXPY inside KSPSolve_BCGS/_CG.
#pragma once
#include "sep_kernels.cuh"

namespace b200 {

constexpr int SEP_TY = 8;  // tile lines in y (one warp per line, 256 threads)

struct SepTileField
{
    int tiles_x, tiles;    // tiles along x, tiles per plane
    int nchunk, zchunk;    // z chunks and planes per chunk
    int block0;            // first CTA of the field
    int surf0;             // first candidate of the field in the list of wrapped cells
};

struct SepTilePlan
{
    SepTileField f[3];
    int stencil_blocks;    // CTAs marching tiles
    int surf_blocks;       // CTAs on the cells with a wrapped periodic neighbour (row per thread)
    int tail_blocks;       // CTAs on the rows behind the stencil blocks (row per thread)
    int surf_cells;        // candidates in the wrapped-cell list (two faces per periodic axis and field)
    __host__ __device__ int blocks() const { return stencil_blocks + surf_blocks + tail_blocks; }
};

// launch geometry: every field cut into (32 xr) x 8 tiles and as many z chunks as FIT into `target_blocks` CTAs (the
// resident slots of the device: a grid slightly above one wave leaves most SMs idle during its second wave -- 480 CTAs on
// 444 slots: 3 320 against 3 700 iterations/s with 384; at least 4 planes per chunk: each chunk re-reads the plane below
// it), or chunks of `zchunk_req` planes when given
inline SepTilePlan sep_tile_plan(const SepDev &A, int xr, int zchunk_req, int target_blocks)
{
    SepTilePlan T{};
    const int TX = 32 * xr;
    long long tiles_all = 0;
    for (int q = 0; q < A.nf; ++q)
    {
        const SepField &F = A.f[q];
        T.f[q].tiles_x = (F.n0 + TX - 1) / TX;
        T.f[q].tiles = T.f[q].tiles_x * ((F.n1 + SEP_TY - 1) / SEP_TY);
        tiles_all += T.f[q].tiles;
    }
    int b0 = 0;
    long long surf = 0;
    for (int q = 0; q < A.nf; ++q)
    {
        const SepField &F = A.f[q];
        const int n2 = F.n2;
        int nch;
        if (zchunk_req > 0) nch = (n2 + zchunk_req - 1) / zchunk_req;
        else
        {
            const long long want = tiles_all > 0 ? target_blocks / tiles_all : 1;
            const long long cap = n2 / 4 > 1 ? n2 / 4 : 1;
            nch = (int)(want < 1 ? 1 : (want > cap ? cap : want));
        }
        const int zc = (n2 + nch - 1) / nch;
        T.f[q].zchunk = zc;
        T.f[q].nchunk = (n2 + zc - 1) / zc;
        T.f[q].block0 = b0;
        b0 += T.f[q].tiles * T.f[q].nchunk;
        T.f[q].surf0 = (int)surf;
        if (F.per0) surf += 2LL * F.n1 * F.n2;
        if (F.per1) surf += 2LL * F.n0 * F.n2;
        if (F.per2) surf += 2LL * F.n0 * F.n1;
    }
    T.stencil_blocks = b0;
    T.surf_cells = (int)surf;
    T.surf_blocks = (int)((surf + 255) / 256 < 128 ? (surf + 255) / 256 : 128);
    const long long tail = A.nrows - A.nsep;
    T.tail_blocks = tail > 0 ? (int)((tail + 255) / 256 < 64 ? (tail + 255) / 256 : 64) : 0;
    return T;
}

// bytes of dynamic shared memory: two operand planes with their rim + R stages of the per-thread queue
template <int XR, int R, int NRIM, int NOWN>
struct SepTileSmem
{
    static constexpr int TX = 32 * XR, LD = TX + 2;
    static constexpr unsigned int PLANE = 8u * (SEP_TY + 2) * LD;
    static constexpr unsigned int OWN = 8u * NOWN * XR * 256, RIM = 8u * NRIM * XR * 96;  // rim slots: warps 0-2 only
    static constexpr unsigned int STAGE = OWN + RIM;
    static constexpr size_t bytes = 2u * PLANE + (size_t)R * STAGE;
};

// Op (one of the policies below):
//   NRIM operand arrays src[0 .. NRIM) -- operand(s) builds the vector entry the product multiplies from their values;
//   NEX extra arrays ex[0 .. NEX) read at the row itself; emit(i, own, t, raw, e) consumes row i's product t
//   (own = operand of the row, raw = its operand-array values, e = its extra-array values).
template <int XR, bool HYB, int R, class Op>
__device__ __forceinline__ void sep_tile_rows(const SepDev &A, const SepTilePlan &T, Op &op)
{
    constexpr int TX = 32 * XR, TY = SEP_TY, LD = TX + 2, NRIM = Op::NRIM, NEX = Op::NEX, NOWN = NRIM + 1 + NEX;
    using SM = SepTileSmem<XR, R, NRIM, NOWN>;
    B200_DYNAMIC_SMEM(smem_raw);
    const unsigned int s_base = smem_u32(smem_raw);
    const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
    int b = blockIdx.x;
    if (b >= T.stencil_blocks)
    {
        // the two row-per-thread parts (sep_row, any column order): cells with a wrapped neighbour, rows behind the blocks
        auto fetch = [&](long long j) { return op.val(j); };
        b -= T.stencil_blocks;
        if (b >= T.surf_blocks)
        {
            for (long long i = A.nsep + (long long)(b - T.surf_blocks) * blockDim.x + tid; i < A.nrows;
                 i += (long long)T.tail_blocks * blockDim.x)
                op.emit_global(i, sep_row(A, i, fetch));
            return;
        }
        for (int q = b * (int)blockDim.x + tid; q < T.surf_cells; q += T.surf_blocks * (int)blockDim.x)
        {
            int fi = 0;
            if (A.nf > 1 && q >= T.f[1].surf0) fi = 1;
            if (A.nf > 2 && q >= T.f[2].surf0) fi = 2;
            const SepField &f = A.f[fi];
            const int n0 = f.n0, n1 = f.n1, n2 = f.n2;
            int l = q - T.f[fi].surf0, i0, i1, i2;
            const int sx = f.per0 ? 2 * n1 * n2 : 0, sy = f.per1 ? 2 * n0 * n2 : 0;
            if (l < sx)
            {
                const int side = l / (n1 * n2), rem = l - side * (n1 * n2);
                i2 = rem / n1;
                i1 = rem - i2 * n1;
                i0 = side ? n0 - 1 : 0;
            }
            else if (l < sx + sy)
            {
                l -= sx;
                const int side = l / (n0 * n2), rem = l - side * (n0 * n2);
                i2 = rem / n0;
                i0 = rem - i2 * n0;
                i1 = side ? n1 - 1 : 0;
                if (f.per0 && (i0 == 0 || i0 == n0 - 1)) continue;  // listed with the x faces
            }
            else
            {
                l -= sx + sy;
                const int side = l / (n0 * n1), rem = l - side * (n0 * n1);
                i1 = rem / n0;
                i0 = rem - i1 * n0;
                i2 = side ? n2 - 1 : 0;
                if ((f.per0 && (i0 == 0 || i0 == n0 - 1)) || (f.per1 && (i1 == 0 || i1 == n1 - 1))) continue;
            }
            const long long i = f.off + ((long long)i2 * n1 + i1) * n0 + i0;
            op.emit_global(i, sep_row(A, i, fetch));
        }
        return;
    }
    int fi = 0;
    if (A.nf > 1 && b >= T.f[1].block0) fi = 1;
    if (A.nf > 2 && b >= T.f[2].block0) fi = 2;
    const SepField &f = A.f[fi];
    const SepTileField &g = T.f[fi];
    b -= g.block0;
    const int ch = b / g.tiles, tl = b - ch * g.tiles;
    const int tyi = tl / g.tiles_x, txi = tl - tyi * g.tiles_x;
    const int n0 = f.n0, n1 = f.n1, n2 = f.n2;
    const int x0 = txi * TX, y0 = tyi * TY;
    const int kb = ch * g.zchunk, ke = min(n2, kb + g.zchunk);
    const int i1 = y0 + ty, xb = x0 + tx;
    const bool in_y = i1 < n1;

    // own cells r = 0 .. XR-1 at i0 = xb + 32 r: row index in plane 0 = own0 + 32 r; `live` bit r: inside the field;
    // `wrap` bit r: a periodic neighbour wraps in x or y (those cells belong to the surface CTAs above)
    const int own0 = (int)(f.off + (long long)i1 * n0 + xb);
    unsigned int xin = 0, live = 0, wrap = 0;
    double cxm[XR], cxp[XR], dxi[XR];
#pragma unroll
    for (int r = 0; r < XR; ++r)
    {
        const int i0 = xb + 32 * r;
        const bool in = i0 < n0;
        if (in) xin |= 1u << r;
        if (in && in_y) live |= 1u << r;
        if ((f.per0 && (i0 == 0 || i0 == n0 - 1)) || (f.per1 && (i1 == 0 || i1 == n1 - 1))) wrap |= 1u << r;
        cxm[r] = (in && in_y) ? f.cm[0][i0] : 0.0;
        cxp[r] = (in && in_y) ? f.cp[0][i0] : 0.0;
        dxi[r] = (HYB && in && in_y) ? f.w[0][i0] : 0.0;
    }
    const double cym = in_y ? f.cm[1][i1] : 0.0, cyp = in_y ? f.cp[1][i1] : 0.0;
    const double dyj = (HYB && in_y) ? f.w[1][i1] : 0.0;

    // rim of the shared plane: warp 0 the line below the tile, warp 1 the line above (XR cells per lane, slot r at
    // hs0 + 32 r, row index h0 + 32 r), 16 lanes of warp 2 the two columns (one cell per lane)
    int h0 = -1, hs0 = -1;
    unsigned int hlive = 0;  // bit r: rim slot r is written by this thread; the value is 0 outside the field
    if (ty < 2)
    {
        const int jy = ty == 0 ? y0 - 1 : y0 + TY;
        hs0 = (ty == 0 ? 0 : TY + 1) * LD + 1 + tx;
        hlive = (1u << XR) - 1u;
        if (jy >= 0 && jy < n1) h0 = (int)(f.off + (long long)jy * n0 + xb);
        else xin = 0;  // only the rim loads of this thread look at xin from here on
    }
    else if (ty == 2 && tx < 2 * TY)
    {
        const int side = tx / TY, yy = tx - side * TY;
        const int jy = y0 + yy, jx = side ? x0 + TX : x0 - 1;
        hs0 = (yy + 1) * LD + (side ? TX + 1 : 0);
        hlive = 1u;
        xin = (jy < n1 && jx >= 0 && jx < n0) ? 1u : 0u;
        h0 = (int)(f.off + (long long)jy * n0 + jx);
    }
    const unsigned int hload = hlive & xin;  // rim cells inside the field

    // ---- the per-thread queue: stage kk % R holds plane kk.  Row indices are 32-bit here (the host refuses systems of
    // 2^31 rows or more), so an address is one IMAD.WIDE.
    const unsigned int s2 = (unsigned int)n0 * (unsigned int)n1;
    const unsigned int q_base = s_base + 2u * SM::PLANE + 8u * (unsigned)tid;
    auto own_slot = [&](int kk, int a, int r) { return q_base + (unsigned)(kk % R) * SM::STAGE + 8u * 256u * (unsigned)(a * XR + r); };
    auto rim_slot = [&](int kk, int a, int r) { return q_base + (unsigned)(kk % R) * SM::STAGE + SM::OWN + 8u * 96u * (unsigned)(a * XR + r); };
    // planes kb - 1 and ke are only needed as z neighbours: operand arrays of the own cells, nothing else
    auto issue = [&](int kk) {
        if (kk >= 0 && kk < n2 && kk <= ke)
        {
            const unsigned int io = (unsigned int)own0 + (unsigned int)kk * s2, ho = (unsigned int)h0 + (unsigned int)kk * s2;
            const bool full = kk >= kb && kk < ke;
#pragma unroll
            for (int r = 0; r < XR; ++r)
                if (live >> r & 1u)
                {
#pragma unroll
                    for (int a = 0; a < NRIM; ++a) cp_async8(own_slot(kk, a, r), op.src[a] + (io + 32u * r));
                    if (full)
                    {
                        cp_async8(own_slot(kk, NRIM, r), A.diag + (io + 32u * r));
#pragma unroll
                        for (int a = 0; a < NEX; ++a) cp_async8(own_slot(kk, NRIM + 1 + a, r), op.ex[a] + (io + 32u * r));
                    }
                }
            if (full)
            {
#pragma unroll
                for (int r = 0; r < XR; ++r)
                    if (hload >> r & 1u)
                    {
#pragma unroll
                        for (int a = 0; a < NRIM; ++a) cp_async8(rim_slot(kk, a, r), op.src[a] + (ho + 32u * r));
                    }
            }
        }
        cp_async_commit();
    };
    auto own_operands = [&](int kk, double (&v)[XR]) {
#pragma unroll
        for (int r = 0; r < XR; ++r)
        {
            v[r] = 0.0;
            if (live >> r & 1u)
            {
                double sv[NRIM];
#pragma unroll
                for (int a = 0; a < NRIM; ++a) sv[a] = lds64(own_slot(kk, a, r));
                v[r] = op.operand(sv);
            }
        }
    };
    auto rim_operands = [&](int kk, double (&h)[XR]) {
#pragma unroll
        for (int r = 0; r < XR; ++r)
        {
            h[r] = 0.0;
            if (hload >> r & 1u)
            {
                double sv[NRIM];
#pragma unroll
                for (int a = 0; a < NRIM; ++a) sv[a] = lds64(rim_slot(kk, a, r));
                h[r] = op.operand(sv);
            }
        }
    };
    const unsigned int s_own = s_base + 8u * (unsigned)((ty + 1) * LD + 1 + tx), s_rim = s_base + 8u * (unsigned)(hs0 < 0 ? 0 : hs0);
    auto store_plane = [&](int buf, const double (&v)[XR], const double (&h)[XR]) {
#pragma unroll
        for (int r = 0; r < XR; ++r) sts64(s_own + buf * SM::PLANE + 256u * r, v[r]);
#pragma unroll
        for (int r = 0; r < XR; ++r)
            if (hlive >> r & 1u) sts64(s_rim + buf * SM::PLANE + 256u * r, h[r]);
    };

    double vm[XR], vc[XR], vp[XR], hv[XR];
#pragma unroll
    for (int q = -1; q < R - 1; ++q) issue(kb + q);
    cp_async_wait<R - 2>();  // planes kb - 1 and kb have landed
#pragma unroll
    for (int r = 0; r < XR; ++r) vm[r] = 0.0;
    if (kb > 0) own_operands(kb - 1, vm);
    own_operands(kb, vc);
    rim_operands(kb, hv);
    store_plane(kb & 1, vc, hv);
    __syncthreads();
    issue(kb + R - 1);  // into the stage of plane kb - 1

    for (int k = kb; k < ke; ++k)
    {
        const int buf = k & 1;
        cp_async_wait<R - 2>();  // plane k + 1 has landed; k + 2 .. k + R - 1 are in flight
#pragma unroll
        for (int r = 0; r < XR; ++r) vp[r] = hv[r] = 0.0;
        if (k + 1 < n2) own_operands(k + 1, vp);
        if (k + 1 < ke) rim_operands(k + 1, hv);
        const bool wz = f.per2 && (k == 0 || k == n2 - 1);
        const double czm = f.cm[2][k], czp = f.cp[2][k];
        const double dzk = HYB ? f.w[2][k] : 0.0;
        const unsigned int todo = wz ? 0u : (live & ~wrap);
        const unsigned int io = (unsigned int)own0 + (unsigned int)k * s2;
#pragma unroll
        for (int r = 0; r < XR; ++r)
        {
            if (!(todo >> r & 1u)) continue;
            const unsigned int i = io + 32u * r;
            const unsigned int c = s_own + buf * SM::PLANE + 256u * r;
            // the four in-plane neighbours (the rim holds zeros outside the field; a term with a zero coefficient is
            // dropped below whatever its operand)
            const double vym = lds64(c - 8u * LD), vxm = lds64(c - 8u), vxp = lds64(c + 8u), vyp = lds64(c + 8u * LD);
            const double dg = lds64(own_slot(k, NRIM, r));
            double axm = cxm[r], axp = cxp[r], aym = cym, ayp = cyp, azm = czm, azp = czp;
            if (HYB)
            {
                // (product of the two other cell widths) * face array: the grouping of the assembled matrix (sep_row)
                const double ayz = __dmul_rn(dyj, dzk), axz = __dmul_rn(dxi[r], dzk), axy = __dmul_rn(dxi[r], dyj);
                axm = __dmul_rn(ayz, axm);
                axp = __dmul_rn(ayz, axp);
                aym = __dmul_rn(axz, aym);
                ayp = __dmul_rn(axz, ayp);
                azm = __dmul_rn(axy, azm);
                azp = __dmul_rn(axy, azp);
            }
            double t = 0.0;
            // a zero coefficient = no entry in the assembled row: the term is not added (select, no branch)
#define B200_TILE_TERM(a, x) t = ((a) != 0.0) ? __dadd_rn(t, __dmul_rn((a), (x))) : t
            B200_TILE_TERM(azm, vm[r]);
            B200_TILE_TERM(aym, vym);
            B200_TILE_TERM(axm, vxm);
            t = __dadd_rn(t, __dmul_rn(dg, vc[r]));
            B200_TILE_TERM(axp, vxp);
            B200_TILE_TERM(ayp, vyp);
            B200_TILE_TERM(azp, vp[r]);
#undef B200_TILE_TERM
            if (A.rem_rowptr)
            {
#pragma unroll 1
                for (int64_t q = A.rem_rowptr[i]; q < A.rem_rowptr[i + 1u]; ++q)
                    t = __dadd_rn(t, __dmul_rn(A.rem_val[q], op.val((unsigned int)A.rem_col[q])));
            }
            double raw[NRIM], e[NEX + 1];
#pragma unroll
            for (int a = 0; a < NRIM; ++a) raw[a] = Op::NEEDS_RAW ? lds64(own_slot(k, a, r)) : 0.0;
#pragma unroll
            for (int a = 0; a < NEX; ++a) e[a] = lds64(own_slot(k, NRIM + 1 + a, r));
            op.emit(i, vc[r], t, raw, e);
        }
        if (k + 1 < ke) store_plane(buf ^ 1, vp, hv);
        __syncthreads();
        issue(k + R);  // into the stage of plane k: this thread has read everything it queued there
#pragma unroll
        for (int r = 0; r < XR; ++r)
        {
            vm[r] = vc[r];
            vc[r] = vp[r];
        }
    }
    cp_async_wait<0>();
}

// ---- the operand / result policies of the four products -------------------------------------------------------------

// the generic-path halves shared by the policies: the operand and the result of a row straight from global memory
template <class Op>
struct SepOpBase
{
    template <class I>
    __device__ __forceinline__ double val(I j) const
    {
        const Op &o = *static_cast<const Op *>(this);
        double sv[Op::NRIM];
#pragma unroll
        for (int a = 0; a < Op::NRIM; ++a) sv[a] = o.src[a][j];
        return o.operand(sv);
    }
    template <class I>
    __device__ __forceinline__ void emit_global(I i, double t)
    {
        Op &o = *static_cast<Op *>(this);
        double raw[Op::NRIM], e[Op::NEX + 1];
#pragma unroll
        for (int a = 0; a < Op::NRIM; ++a) raw[a] = o.src[a][i];
#pragma unroll
        for (int a = 0; a < Op::NEX; ++a) e[a] = o.ex[a][i];
        o.emit(i, o.operand(raw), t, raw, e);
    }
};

struct SepOpApply : SepOpBase<SepOpApply>
{
    static constexpr int NRIM = 1, NEX = 0;
    static constexpr bool NEEDS_RAW = false;
    const double *src[1];
    const double *ex[1];
    double *y;
    __device__ __forceinline__ double operand(const double (&s)[1]) const { return s[0]; }
    template <class I>
    __device__ __forceinline__ void emit(I i, double, double t, const double (&)[1], const double (&)[1]) { y[i] = t; }
};

// CG class 0:  x += a' p' ; p = z + b p' ; w = A p ; p.w   (k_sep_cg_spmv).  Operand arrays: r, p', [1/diag], [null vector];
// extra: x.
template <bool JACOBI, int NULLMODE>
struct SepOpCg : SepOpBase<SepOpCg<JACOBI, NULLMODE>>
{
    static constexpr int NRIM = 2 + (JACOBI ? 1 : 0) + (NULLMODE == 2 ? 1 : 0), NEX = 1;
    static constexpr bool NEEDS_RAW = true;
    const double *src[NRIM];
    const double *ex[1];
    double *x, *p_out, *w;
    double shift, bcoef, aprev;
    bool xupd;
    double acc[1];
    __device__ __forceinline__ double operand(const double (&s)[NRIM]) const
    {
        double z = s[0];  // csr_z: B r with the null space removed
        if (JACOBI) z = __dmul_rn(z, s[2]);
        if (NULLMODE == 1) z = __dadd_rn(z, shift);
        if (NULLMODE == 2) z = __dadd_rn(z, __dmul_rn(shift, s[2 + (JACOBI ? 1 : 0)]));
        return __dadd_rn(z, __dmul_rn(bcoef, s[1]));
    }
    template <class I>
    __device__ __forceinline__ void emit(I i, double own, double t, const double (&raw)[NRIM], const double (&e)[2])
    {
        if (xupd) x[i] = __dadd_rn(e[0], __dmul_rn(aprev, raw[1]));
        p_out[i] = own;
        w[i] = t;
        acc[0] = fma(own, t, acc[0]);
    }
};

// BiCGStab: v = B A p ; v.rp   (k_sep_bcgs_spmv1).  Operand array: p; extras: r', [1/diag].
template <bool JACOBI>
struct SepOpBcgs1 : SepOpBase<SepOpBcgs1<JACOBI>>
{
    static constexpr int NRIM = 1, NEX = JACOBI ? 2 : 1;
    static constexpr bool NEEDS_RAW = false;
    const double *src[1];
    const double *ex[NEX];
    double *vv;
    double acc[1];
    __device__ __forceinline__ double operand(const double (&s)[1]) const { return s[0]; }
    template <class I>
    __device__ __forceinline__ void emit(I i, double, double t, const double (&)[1], const double (&e)[NEX + 1])
    {
        if (JACOBI) t = __dmul_rn(t, e[1]);
        vv[i] = t;
        acc[0] = fma(t, e[0], acc[0]);
    }
};

// BiCGStab: s = r - alpha v ; t = B A s ; {s.t, t.t, s.s}   (k_sep_bcgs_spmv2).  Operand arrays: r, v; extra: [1/diag].
template <bool JACOBI>
struct SepOpBcgs2 : SepOpBase<SepOpBcgs2<JACOBI>>
{
    static constexpr int NRIM = 2, NEX = JACOBI ? 1 : 0;
    static constexpr bool NEEDS_RAW = false;
    const double *src[2];
    const double *ex[1];
    double *s, *t_out;
    double malpha;
    double acc[3];
    __device__ __forceinline__ double operand(const double (&q)[2]) const { return __dadd_rn(__dmul_rn(malpha, q[1]), q[0]); }  // VecWAXPY(S,-alpha,V,R)
    template <class I>
    __device__ __forceinline__ void emit(I i, double own, double t, const double (&)[2], const double (&e)[NEX + 1])
    {
        s[i] = own;
        if (JACOBI) t = __dmul_rn(t, e[0]);
        t_out[i] = t;
        acc[0] = fma(own, t, acc[0]);
        acc[1] = fma(t, t, acc[1]);
        acc[2] = fma(own, own, acc[2]);
    }
};

template <int XR, int R, class Op>
constexpr size_t sep_tile_smem_bytes()
{
    return SepTileSmem<XR, R, Op::NRIM, Op::NRIM + 1 + Op::NEX>::bytes;
}

template <int XR, bool HYB, int R>
__global__ void __launch_bounds__(256) k_sep_tile_apply(SepDev A, SepTilePlan T, const double *x, double *y)
{
    SepOpApply op;
    op.src[0] = x;
    op.ex[0] = nullptr;
    op.y = y;
    sep_tile_rows<XR, HYB, R>(A, T, op);
}

template <int XR, bool HYB, int R, bool JACOBI, int NULLMODE>
__global__ void __launch_bounds__(256) k_sep_tile_cg_spmv(SepDev A, SepTilePlan T, CsrVecs v, ReduceWs ws, DevState *st, SolveConsts kc,
                                                          double *hist)
{
    if (st->done) return;
    SepOpCg<JACOBI, NULLMODE> op;
    op.src[0] = v.r;
    op.src[1] = v.p_in;
    if (JACOBI) op.src[2] = v.dinv;
    if (NULLMODE == 2) op.src[2 + (JACOBI ? 1 : 0)] = v.nv;
    op.ex[0] = v.x;
    op.x = v.x;
    op.p_out = v.p_out;
    op.w = v.w;
    op.shift = st->shift;
    op.bcoef = st->b;
    op.aprev = st->a;
    op.xupd = st->pending != 0;
    op.acc[0] = 0.0;
    sep_tile_rows<XR, HYB, R>(A, T, op);
    csr_reduce_finalize<1>(op.acc, FIN_SPMV, ws, st, kc, hist);
}

template <int XR, int R, bool JACOBI>
__global__ void __launch_bounds__(256) k_sep_tile_bcgs_spmv1(SepDev A, SepTilePlan T, const double *p, const double *dinv,
                                                             const double *rp, double *vv, ReduceWs ws, DevState *st, SolveConsts kc,
                                                             double *hist)
{
    if (st->done) return;
    SepOpBcgs1<JACOBI> op;
    op.src[0] = p;
    op.ex[0] = rp;
    if (JACOBI) op.ex[JACOBI ? 1 : 0] = dinv;
    op.vv = vv;
    op.acc[0] = 0.0;
    sep_tile_rows<XR, false, R>(A, T, op);
    csr_reduce_finalize<1>(op.acc, FIN_BCGS_D1, ws, st, kc, hist);
}

template <int XR, int R, bool JACOBI>
__global__ void __launch_bounds__(256) k_sep_tile_bcgs_spmv2(SepDev A, SepTilePlan T, const double *r, const double *vv,
                                                             const double *dinv, double *s, double *t_out, ReduceWs ws, DevState *st,
                                                             SolveConsts kc, double *hist)
{
    if (st->done) return;
    SepOpBcgs2<JACOBI> op;
    op.src[0] = r;
    op.src[1] = vv;
    op.ex[0] = dinv;
    op.s = s;
    op.t_out = t_out;
    op.malpha = -st->alpha;
    op.acc[0] = op.acc[1] = op.acc[2] = 0.0;
    sep_tile_rows<XR, false, R>(A, T, op);
    csr_reduce_finalize<3>(op.acc, FIN_BCGS_OMEGA, ws, st, kc, hist);
}

}  // namespace b200
