// sep_tile.cuh -- tiled, plane-marching kernels for the line-coefficient operator of sep_kernels.cuh.
//
// The row-per-thread kernels (k_sep_*) spend ~350 warp instructions per warp-row: two integer divisions for (i, j, k), a
// branch per stencil term and one gather per neighbour (two for BiCGStab's s = r - alpha v) -- issue- and latency-bound
// at 2.0-2.3 TB/s (profiles/r02_velocity_bench.log).  Here a CTA owns a (32 XR) x 8 tile of one field's (x, y) plane and
// marches it through a chunk of z: the operand of a cell is built ONCE per plane (registers carry it across k - 1, k,
// k + 1; a double-buffered shared plane with a one-cell rim hands it to the x/y neighbours), the x/y coefficients live
// in registers for the whole march, and nothing is divided per row.  One __syncthreads per plane.
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

// launch geometry: every field cut into (32 xr) x 8 tiles and as many z chunks as it takes to reach `target_blocks` CTAs
// (at least 4 planes per chunk: each chunk re-reads one plane below it), or chunks of `zchunk_req` planes when given
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
            const long long want = tiles_all > 0 ? (target_blocks + tiles_all - 1) / tiles_all : 1;
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

constexpr size_t sep_tile_smem_bytes(int xr) { return sizeof(double) * 2 * (SEP_TY + 2) * (32 * xr + 2); }

// Op: double val(I j) -- the operand vector at row j, built from global memory (I: long long or unsigned int);
//     void emit(I i, double own, double t) -- consumes row i's product t (own = val(i))
template <int XR, bool HYB, class Op>
__device__ __forceinline__ void sep_tile_rows(const SepDev &A, const SepTilePlan &T, Op &op)
{
    constexpr int TX = 32 * XR, TY = SEP_TY, LD = TX + 2;
    B200_DYNAMIC_SMEM(smem_raw);  // two planes of (TY + 2) x LD doubles: sep_tile_smem_bytes(XR)
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
            {
                const double own = op.val(i);
                op.emit(i, own, sep_row(A, i, fetch));
            }
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
            const double own = op.val(i);
            op.emit(i, own, sep_row(A, i, fetch));
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

    double vm[XR], vc[XR], vp[XR], hv[XR];
    // io / ho: row index of own cell 0 / rim cell 0 in the plane being loaded.  Row indices are 32-bit here (the host
    // refuses systems of 2^31 rows or more), so an address is one IMAD.WIDE; an index that is never dereferenced may wrap.
    auto load_plane = [&](unsigned int io, unsigned int ho, double (&v)[XR], double (&h)[XR]) {
#pragma unroll
        for (int r = 0; r < XR; ++r) v[r] = (live >> r & 1u) ? op.val(io + 32u * r) : 0.0;
#pragma unroll
        for (int r = 0; r < XR; ++r) h[r] = ((hlive & xin) >> r & 1u) ? op.val(ho + 32u * r) : 0.0;
    };
    constexpr unsigned int PLANE = 8u * (TY + 2) * LD;
    const unsigned int s_own = s_base + 8u * (unsigned)((ty + 1) * LD + 1 + tx), s_rim = s_base + 8u * (unsigned)(hs0 < 0 ? 0 : hs0);
    auto store_plane = [&](int buf, const double (&v)[XR], const double (&h)[XR]) {
#pragma unroll
        for (int r = 0; r < XR; ++r) sts64(s_own + buf * PLANE + 256u * r, v[r]);
#pragma unroll
        for (int r = 0; r < XR; ++r)
            if (hlive >> r & 1u) sts64(s_rim + buf * PLANE + 256u * r, h[r]);
    };

    const unsigned int s2 = (unsigned int)n0 * (unsigned int)n1;
    unsigned int io = (unsigned int)own0 + (unsigned int)kb * s2, ho = (unsigned int)h0 + (unsigned int)kb * s2;
#pragma unroll
    for (int r = 0; r < XR; ++r) vm[r] = (kb > 0 && (live >> r & 1u)) ? op.val(io - s2 + 32u * r) : 0.0;
    load_plane(io, ho, vc, hv);
    store_plane(kb & 1, vc, hv);
    __syncthreads();

    for (int k = kb; k < ke; ++k)
    {
        const int buf = k & 1;
        if (k + 1 < n2) load_plane(io + s2, ho + s2, vp, hv);
        else
        {
#pragma unroll
            for (int r = 0; r < XR; ++r) vp[r] = hv[r] = 0.0;
        }
        const bool wz = f.per2 && (k == 0 || k == n2 - 1);
        const double czm = f.cm[2][k], czp = f.cp[2][k];
        const double dzk = HYB ? f.w[2][k] : 0.0;
        const unsigned int todo = wz ? 0u : (live & ~wrap);
#pragma unroll
        for (int r = 0; r < XR; ++r)
        {
            if (!(todo >> r & 1u)) continue;
            const unsigned int i = io + 32u * r;
            const unsigned int c = s_own + buf * PLANE + 256u * r;
            // the four in-plane neighbours (the rim holds zeros outside the field; a term with a zero coefficient is
            // dropped below whatever its operand)
            const double vym = lds64(c - 8u * LD), vxm = lds64(c - 8u), vxp = lds64(c + 8u), vyp = lds64(c + 8u * LD);
            const double dg = A.diag[i];
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
            op.emit(i, vc[r], t);
        }
        if (k + 1 < ke) store_plane(buf ^ 1, vp, hv);
        __syncthreads();
        io += s2;
        ho += s2;
#pragma unroll
        for (int r = 0; r < XR; ++r)
        {
            vm[r] = vc[r];
            vc[r] = vp[r];
        }
    }
}

// ---- the operand / result policies of the four products -------------------------------------------------------------

struct SepOpApply
{
    const double *x;
    double *y;
    template <class I>
    __device__ __forceinline__ double val(I j) const { return x[j]; }
    template <class I>
    __device__ __forceinline__ void emit(I i, double, double t) { y[i] = t; }
};

// CG class 0:  x += a' p' ; p = z + b p' ; w = A p ; p.w   (k_sep_cg_spmv)
template <bool JACOBI, int NULLMODE>
struct SepOpCg
{
    CsrVecs v;
    double shift, bcoef, aprev;
    bool xupd;
    double acc[1];
    template <class I>
    __device__ __forceinline__ double val(I j) const
    {
        return __dadd_rn(csr_z<JACOBI, NULLMODE>(v, j, shift), __dmul_rn(bcoef, v.p_in[j]));
    }
    template <class I>
    __device__ __forceinline__ void emit(I i, double own, double t)
    {
        if (xupd) v.x[i] = __dadd_rn(v.x[i], __dmul_rn(aprev, v.p_in[i]));
        v.p_out[i] = own;
        v.w[i] = t;
        acc[0] = fma(own, t, acc[0]);
    }
};

// BiCGStab: v = B A p ; v.rp   (k_sep_bcgs_spmv1)
template <bool JACOBI>
struct SepOpBcgs1
{
    const double *p, *dinv, *rp;
    double *vv;
    double acc[1];
    template <class I>
    __device__ __forceinline__ double val(I j) const { return p[j]; }
    template <class I>
    __device__ __forceinline__ void emit(I i, double, double t)
    {
        if (JACOBI) t = __dmul_rn(t, dinv[i]);
        vv[i] = t;
        acc[0] = fma(t, rp[i], acc[0]);
    }
};

// BiCGStab: s = r - alpha v ; t = B A s ; {s.t, t.t, s.s}   (k_sep_bcgs_spmv2)
template <bool JACOBI>
struct SepOpBcgs2
{
    const double *r, *vv, *dinv;
    double *s, *t_out;
    double malpha;
    double acc[3];
    template <class I>
    __device__ __forceinline__ double val(I j) const { return __dadd_rn(__dmul_rn(malpha, vv[j]), r[j]); }  // VecWAXPY(S,-alpha,V,R)
    template <class I>
    __device__ __forceinline__ void emit(I i, double own, double t)
    {
        s[i] = own;
        if (JACOBI) t = __dmul_rn(t, dinv[i]);
        t_out[i] = t;
        acc[0] = fma(own, t, acc[0]);
        acc[1] = fma(t, t, acc[1]);
        acc[2] = fma(own, own, acc[2]);
    }
};

template <int XR, bool HYB>
__global__ void __launch_bounds__(256) k_sep_tile_apply(SepDev A, SepTilePlan T, const double *x, double *y)
{
    SepOpApply op{x, y};
    sep_tile_rows<XR, HYB>(A, T, op);
}

template <int XR, bool HYB, bool JACOBI, int NULLMODE>
__global__ void __launch_bounds__(256) k_sep_tile_cg_spmv(SepDev A, SepTilePlan T, CsrVecs v, ReduceWs ws, DevState *st, SolveConsts kc,
                                                          double *hist)
{
    if (st->done) return;
    SepOpCg<JACOBI, NULLMODE> op{v, st->shift, st->b, st->a, st->pending != 0, {0.0}};
    sep_tile_rows<XR, HYB>(A, T, op);
    csr_reduce_finalize<1>(op.acc, FIN_SPMV, ws, st, kc, hist);
}

template <int XR, bool JACOBI>
__global__ void __launch_bounds__(256) k_sep_tile_bcgs_spmv1(SepDev A, SepTilePlan T, const double *p, const double *dinv,
                                                             const double *rp, double *vv, ReduceWs ws, DevState *st, SolveConsts kc,
                                                             double *hist)
{
    if (st->done) return;
    SepOpBcgs1<JACOBI> op{p, dinv, rp, vv, {0.0}};
    sep_tile_rows<XR, false>(A, T, op);
    csr_reduce_finalize<1>(op.acc, FIN_BCGS_D1, ws, st, kc, hist);
}

template <int XR, bool JACOBI>
__global__ void __launch_bounds__(256) k_sep_tile_bcgs_spmv2(SepDev A, SepTilePlan T, const double *r, const double *vv,
                                                             const double *dinv, double *s, double *t_out, ReduceWs ws, DevState *st,
                                                             SolveConsts kc, double *hist)
{
    if (st->done) return;
    SepOpBcgs2<JACOBI> op{r, vv, dinv, s, t_out, -st->alpha, {0.0, 0.0, 0.0}};
    sep_tile_rows<XR, false>(A, T, op);
    csr_reduce_finalize<3>(op.acc, FIN_BCGS_OMEGA, ws, st, kc, hist);
}

}  // namespace b200
