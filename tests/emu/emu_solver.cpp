// emu_solver.cpp -- TEST SCAFFOLDING: drives the UNMODIFIED kernel sources of petibm_b200/csrc (kernels.cuh,
// spmv2.cuh) under the fiber emulation of tests/emu/cuda_emu.h, single rank, with a host loop that mirrors
// solve_stencil_cg of b200ls.cu (state reset, scatter, init passes, {k_spmv2, k_update2} until done, x tail,
// gather).  Built by tests/test_emulated_kernels.py with g++ -DB200_EMULATE -ffp-contract=off; never shipped.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "spmv2.cuh"
#include "spmv3.cuh"
#include "spmv4.cuh"
#include "spmv5.cuh"
#include "update_fly.cuh"
#include "csr_kernels.cuh"
#include "sep_kernels.cuh"
#include "sep_tile.cuh"
#include "mg_kernels.cuh"
#include "ops_kernels.cuh"
#include "dense_kernels.cuh"

using namespace b200;

namespace {

struct Problem
{
    GridDev g{};
    std::vector<double> axes;
    size_t vec_elems = 0;
    int64_t nlocal = 0;
    int per[3] = {0, 0, 0};
};

void build(Problem &P, int dim, const int64_t *n, const int *per, const double *dx, const double *dy, const double *dz,
           double dt)
{
    const int64_t nx = n[0], ny = n[1], nz = dim == 3 ? n[2] : 1;
    std::vector<double> hdx(dx, dx + nx), hdy(dy, dy + ny), hdz;
    if (dim == 3) hdz.assign(dz, dz + nz);
    else hdz.assign(1, 1.0);
    auto faces = [&](const std::vector<double> &d, bool periodic, bool active) {
        const size_t m = d.size();
        std::vector<double> gg(m + 1, 0.0);
        if (!active) return gg;
        for (size_t s = 1; s < m; ++s)
        {
            const double hh = 0.5 * (d[s] + d[s - 1]);
            const double inv = 1.0 / hh;
            gg[s] = dt * inv;
        }
        if (periodic)
        {
            const double hh = 0.5 * (d[0] + d[m - 1]);
            const double inv = 1.0 / hh;
            gg[0] = gg[m] = dt * inv;
        }
        return gg;
    };
    for (int d = 0; d < 3; ++d) P.per[d] = d < dim ? (per[d] != 0) : 0;
    const auto gx = faces(hdx, P.per[0], true), gy = faces(hdy, P.per[1], true), gz = faces(hdz, P.per[2], dim == 3);
    P.axes.clear();
    size_t o[6];
    const std::vector<double> *parts[6] = {&hdx, &hdy, &hdz, &gx, &gy, &gz};
    for (int q = 0; q < 6; ++q)
    {
        o[q] = P.axes.size();
        P.axes.insert(P.axes.end(), parts[q]->begin(), parts[q]->end());
    }
    GridDev &g = P.g;
    g.nx = (int)nx;
    g.ny = (int)ny;
    g.nzl = (int)nz;
    g.px = (int)((nx + 15) / 16 * 16);
    g.plane = (long long)g.px * g.ny;
    g.perx = P.per[0];
    g.pery = P.per[1];
    g.perz_wrap = P.per[2] ? 1 : 0;
    g.kz0 = 0;
    g.nzg = (int)nz;
    g.wrapz_lo = P.per[2] ? 1 : 0;
    g.wrapz_hi = P.per[2] ? 1 : 0;
    g.dx = P.axes.data() + o[0];
    g.dy = P.axes.data() + o[1];
    g.dz = P.axes.data() + o[2];
    g.gx = P.axes.data() + o[3];
    g.gy = P.axes.data() + o[4];
    g.gz = P.axes.data() + o[5];
    P.nlocal = nx * ny * nz;
    P.vec_elems = (size_t)g.plane * (size_t)(g.nzl + 2);
}

struct Ws
{
    std::vector<double> partials;
    unsigned int counter[4] = {0, 0, 0, 0};
    ReduceWs ws{};
    CommDev cm{};
    Ws()
    {
        partials.assign(8 * 65536, 0.0);
        ws.partials = partials.data();
        ws.counter = counter;
        ws.trace = nullptr;
        memset(&cm, 0, sizeof cm);
        cm.nranks = 1;
    }
};

template <int TYT, int S, int MINB, bool JAC, bool APPLY>
void launch_spmv(const Problem &P, const VecSet &v, int kz, Ws &W, DevState *st, const SolveConsts &kc, double *hist)
{
    using L = Spmv2Smem<32, TYT, S, JAC, APPLY>;
    const GridDev &g = P.g;
    dim3 grid((unsigned)((g.nx + 63) / 64), (unsigned)((g.ny + TYT - 3) / (TYT - 2)), (unsigned)((g.nzl + kz - 1) / kz));
    dim3 block(32, TYT);
    const bool periodic = P.per[0] || P.per[1] || P.per[2];
    if (periodic)
        emu::launch(grid, block, L::total(kz), [&] { k_spmv2<32, TYT, S, MINB, JAC, APPLY, true>(g, v, kz, W.ws, W.cm, st, kc, hist, 0); });
    else
        emu::launch(grid, block, L::total(kz), [&] { k_spmv2<32, TYT, S, MINB, JAC, APPLY, false>(g, v, kz, W.ws, W.cm, st, kc, hist, 0); });
}

// k_spmv3 (balanced split): `kz` is reused as the number of CTAs so that tests can force multi-segment ranges
template <int TYT, int S, int MINB, bool JAC, bool SPLIT>
void launch_spmv3(const Problem &P, const VecSet &v, int nctas, Ws &W, DevState *st, const SolveConsts &kc, double *hist)
{
    using L = Spmv3Smem<32, TYT, S, JAC, false, SPLIT ? 4 : 3>;
    const GridDev &g = P.g;
    const long long ntiles = (long long)((g.nx + 63) / 64) * ((g.ny + TYT - 3) / (TYT - 2));
    const long long total = ntiles * g.nzl;
    if (nctas <= 0) nctas = 3;
    const int ppc = (int)((total + nctas - 1) / nctas);
    const int grid = (int)((total + ppc - 1) / ppc) + 1;   // one CTA without work on purpose
    const int max_seg = ppc < g.nzl ? ppc : g.nzl;
    const bool periodic = P.per[0] || P.per[1] || P.per[2];
    if (periodic)
        emu::launch(dim3(grid), dim3(32, TYT), L::total(max_seg), [&] { k_spmv3<32, TYT, S, MINB, JAC, false, true, SPLIT>(g, v, ppc, max_seg, W.ws, W.cm, st, kc, hist, 0); });
    else
        emu::launch(dim3(grid), dim3(32, TYT), L::total(max_seg), [&] { k_spmv3<32, TYT, S, MINB, JAC, false, false, SPLIT>(g, v, ppc, max_seg, W.ws, W.cm, st, kc, hist, 0); });
}

// k_spmv4 (TMA boxes + mbarrier pipeline): the emulated tensor map carries what cuTensorMapEncodeTiled is given in
// b200ls.cu (tma_map_for): extents (nx, ny, nzl+2), strides (1, px, plane), box (bw, bh, 1)
inline TmaMap emu_tma_map(const GridDev &g, const double *vec, int bw, int bh)
{
    TmaMap m;
    m.base = vec;
    m.dim[0] = g.nx; m.dim[1] = g.ny; m.dim[2] = g.nzl + 2;
    m.stride[0] = 1; m.stride[1] = g.px; m.stride[2] = g.plane;
    m.box[0] = bw; m.box[1] = bh; m.box[2] = 1;
    return m;
}
template <int TY, int S, int MINB, bool JAC, bool BAL = false>
void launch_spmv4(const Problem &P, const VecSet &v, int kz, Ws &W, DevState *st, const SolveConsts &kc, double *hist, int ghost_store = 0)
{
    using L = Spmv4Smem<TY, S, JAC>;
    const GridDev &g = P.g;
    dim3 grid((unsigned)((g.nx + 63) / 64), (unsigned)((g.ny + TY - 1) / TY), (unsigned)((g.nzl + kz - 1) / kz));
    int table = kz;
    if (BAL)
    {
        // `kz` is reused as the number of CTAs so that tests can force ranges that start mid-column and span columns;
        // one CTA more than needed on purpose (a CTA without work still takes part in the grid reduction)
        const long long total = (long long)grid.x * grid.y * g.nzl;
        const int nctas = kz > 0 && kz < g.nzl ? kz : 3;
        const int ppc = (int)((total + nctas - 1) / nctas);
        grid = dim3((unsigned)((total + ppc - 1) / ppc) + 1);
        kz = ppc;
        table = g.nzl;
    }
    Spmv4Maps maps;
    maps.r = emu_tma_map(g, v.r, L::BW, L::BH);
    maps.p = emu_tma_map(g, v.p_in, L::BW, L::BH);
    maps.x = emu_tma_map(g, v.x, L::BX, TY);
    maps.d = emu_tma_map(g, JAC ? v.dinv : v.r, L::BW, L::BH);
    emu::launch(grid, dim3(32, TY + 1), L::total(table), [&] { k_spmv4<TY, S, MINB, JAC, BAL>(maps, g, v, kz, W.ws, W.cm, st, kc, hist, ghost_store); });
}

template <bool JAC, bool APPLY>
void spmv(const Problem &P, int tile, const VecSet &v, int kz, Ws &W, DevState *st, const SolveConsts &kc, double *hist)
{
    if constexpr (!APPLY)
    {
        const bool periodic = P.per[0] || P.per[1] || P.per[2];
        if (tile >= 50 && tile < 60 && !periodic)
        {
            const GridDev &g = P.g;
            auto run = [&](auto ty_c, auto kb_c) {
                constexpr int TY = decltype(ty_c)::value, KB = decltype(kb_c)::value;
                dim3 grid((unsigned)((g.nx + 63) / 64), (unsigned)((g.ny + TY - 1) / TY), (unsigned)((g.nzl + KB - 1) / KB));
                emu::launch(grid, dim3(32, TY), 0, [&] { k_spmv5<TY, KB, 1, JAC>(g, v, W.ws, W.cm, st, kc, hist, 0); });
            };
            if (tile == 50) return run(std::integral_constant<int, 8>{}, std::integral_constant<int, 4>{});
            if (tile == 51) return run(std::integral_constant<int, 8>{}, std::integral_constant<int, 8>{});
            if (tile == 52) return run(std::integral_constant<int, 4>{}, std::integral_constant<int, 2>{});
            return run(std::integral_constant<int, 8>{}, std::integral_constant<int, 16>{});
        }
        if (((tile >= 40 && tile < 50) || (tile >= 60 && tile < 70)) && !periodic)
        {
            if (tile == 60) return launch_spmv4<6, 4, 3, JAC>(P, v, kz, W, st, kc, hist);
            if (tile == 62) return launch_spmv4<5, 4, 3, JAC>(P, v, kz, W, st, kc, hist);
            if (tile == 40) return launch_spmv4<8, 4, 2, JAC>(P, v, kz, W, st, kc, hist);
            if (tile == 41) return launch_spmv4<4, 4, 4, JAC>(P, v, kz, W, st, kc, hist);
            if (tile == 42) return launch_spmv4<8, 3, 3, JAC>(P, v, kz, W, st, kc, hist);
            if (tile == 46) return launch_spmv4<4, 4, 4, JAC, true>(P, v, kz, W, st, kc, hist);
            if (tile == 47) return launch_spmv4<8, 3, 2, JAC, true>(P, v, kz, W, st, kc, hist);
            return launch_spmv4<16, 3, 1, JAC>(P, v, kz, W, st, kc, hist);
        }
        if (tile == 30) return launch_spmv3<12, 3, 2, JAC, false>(P, v, kz, W, st, kc, hist);
        if (tile == 31) return launch_spmv3<8, 4, 3, JAC, false>(P, v, kz, W, st, kc, hist);
        if (tile == 32) return launch_spmv3<12, 3, 2, JAC, true>(P, v, kz, W, st, kc, hist);
        if (tile == 33) return launch_spmv3<8, 4, 3, JAC, true>(P, v, kz, W, st, kc, hist);
    }
    if (tile == 18 && !APPLY) launch_spmv<12, 3, 2, JAC, APPLY>(P, v, kz, W, st, kc, hist);
    else if (tile == 13 && !APPLY) launch_spmv<6, 4, 4, JAC, APPLY>(P, v, kz, W, st, kc, hist);
    else if (tile == 15 && !APPLY) launch_spmv<8, 3, 3, JAC, APPLY>(P, v, kz, W, st, kc, hist);
    else launch_spmv<8, 4, 3, JAC, APPLY>(P, v, kz, W, st, kc, hist);
}

int g_upd_fly = 0;  // emu_set_update_variant: 1 = k_update2f (Jacobi diagonal rebuilt on the fly)

template <bool JAC, bool INIT>
void update(const Problem &P, const UpdVecs &v, int fin_kind, int blocks, Ws &W, DevState *st, const SolveConsts &kc,
            double *hist)
{
    const GridDev &g = P.g;
    if (JAC && g_upd_fly)
    {
        if (g.px != g.nx)
            emu::launch(dim3(blocks), dim3(256), 0, [&] { k_update2f<INIT, true, false, 4>(g, v, fin_kind, W.ws, W.cm, st, kc, hist); });
        else
            emu::launch(dim3(blocks), dim3(256), 0, [&] { k_update2f<INIT, false, false, 4>(g, v, fin_kind, W.ws, W.cm, st, kc, hist); });
        return;
    }
    if (g.px != g.nx)
        emu::launch(dim3(blocks), dim3(256), 0, [&] { k_update2<JAC, INIT, true, false, 4>(g, v, fin_kind, W.ws, W.cm, st, kc, hist); });
    else
        emu::launch(dim3(blocks), dim3(256), 0, [&] { k_update2<JAC, INIT, false, false, 4>(g, v, fin_kind, W.ws, W.cm, st, kc, hist); });
}

}  // namespace

#define EMU_API __attribute__((visibility("default")))

extern "C" {

EMU_API void emu_set_update_variant(int fly) { g_upd_fly = fly; }

// seed != 0: shuffled fiber order + random preemption at shared-memory accesses; 0: deterministic round-robin
EMU_API void emu_set_schedule(unsigned long long seed)
{
    emu::fuzz = seed != 0;
    emu::rng = seed * 0x9E3779B97F4A7C15ull + 1ull;
}

// y = A x through k_spmv2 in APPLY mode (the b200ls_apply path)
EMU_API int emu_stencil_apply(int dim, const int64_t *n, const int *per, const double *dx, const double *dy, const double *dz,
                      double dt, int kz, const double *x, double *y)
{
    Problem P;
    build(P, dim, n, per, dx, dy, dz, dt);
    Ws W;
    std::vector<double> r(P.vec_elems, 0.0), w(P.vec_elems, 0.0);
    DevState st{};
    SolveConsts kc{};
    emu::launch(dim3(4), dim3(256), 0, [&] { k_scatter(P.g, x, r.data(), nullptr); });
    VecSet v{r.data(), nullptr, nullptr, w.data(), nullptr, nullptr};
    if (kz <= 0) kz = P.g.nzl;
    spmv<false, true>(P, 10, v, kz, W, &st, kc, nullptr);
    emu::launch(dim3(4), dim3(256), 0, [&] { k_gather(P.g, w.data(), y); });
    return 0;
}

// KSPSolve (CG) through the emulated kernels; mirrors solve_stencil_cg
EMU_API int emu_stencil_cg(int dim, const int64_t *n, const int *per, const double *dx, const double *dy, const double *dz, double dt,
                   int jacobi, int has_const, int norm_type, double rtol, double atol, double divtol, int max_it, int tile,
                   int kz, int upd_blocks, int upd_reverse, const double *b, double *x_out, double *hist, int hist_cap,
                   int *nhist, int *its, int *reason, double *rnorm)
{
    Problem P;
    build(P, dim, n, per, dx, dy, dz, dt);
    Ws W;
    const size_t ve = P.vec_elems;
    std::vector<double> r(ve, 0.0), p0(ve, 0.0), p1(ve, 0.0), w(ve, 0.0), x(ve, 0.0), dinv;
    if (jacobi)
    {
        dinv.assign(ve, 0.0);
        emu::launch(dim3(4), dim3(256), 0, [&] { k_jacobi_setup(P.g, dinv.data()); });
    }
    SolveConsts kc{};
    kc.rtol = rtol;
    kc.atol = atol;
    kc.divtol = divtol;
    kc.nglobal = (double)P.nlocal;
    kc.max_it = max_it;
    kc.norm_type = norm_type;
    kc.has_const = has_const;
    kc.hist_cap = hist_cap;
    DevState st{};
    emu::launch(dim3(1), dim3(32), 0, [&] { k_state_reset(&st); });
    emu::launch(dim3(4), dim3(256), 0, [&] { k_scatter(P.g, b, r.data(), x.data()); });
    if (kz <= 0 && (tile < 30 || tile >= 40)) kz = P.g.nzl;  // (tile 46 / 47 reuse kz as the CTA count)
    if (upd_blocks <= 0) upd_blocks = 3;
    UpdVecs uv{r.data(), w.data(), jacobi ? dinv.data() : nullptr, upd_reverse};
    if (has_const)
    {
        if (jacobi) update<true, true>(P, uv, FIN_INIT_CENTRE, upd_blocks, W, &st, kc, hist);
        else update<false, true>(P, uv, FIN_INIT_CENTRE, upd_blocks, W, &st, kc, hist);
    }
    if (jacobi) update<true, true>(P, uv, FIN_INIT, upd_blocks, W, &st, kc, hist);
    else update<false, true>(P, uv, FIN_INIT, upd_blocks, W, &st, kc, hist);
    double *pp[2] = {p0.data(), p1.data()};
    for (int it = 0; it < max_it + 2 && !st.done; ++it)
    {
        VecSet v{r.data(), pp[it & 1], pp[(it & 1) ^ 1], w.data(), x.data(), jacobi ? dinv.data() : nullptr};
        if (jacobi) spmv<true, false>(P, tile, v, kz, W, &st, kc, hist);
        else spmv<false, false>(P, tile, v, kz, W, &st, kc, hist);
        if (jacobi) update<true, false>(P, uv, FIN_UPDATE, upd_blocks, W, &st, kc, hist);
        else update<false, false>(P, uv, FIN_UPDATE, upd_blocks, W, &st, kc, hist);
    }
    emu::launch(dim3(4), dim3(256), 0, [&] { k_xtail(P.g, x.data(), p0.data(), p1.data(), &st); });
    emu::launch(dim3(4), dim3(256), 0, [&] { k_gather(P.g, x.data(), x_out); });
    *nhist = st.nhist;
    *its = st.its;
    *reason = st.reason;
    *rnorm = st.dp;
    return st.done ? 0 : 1;
}

// ---- R emulated ranks (z-slabs), NCCL-style transport: every rank's kernel leaves its partial sums in its send
// buffer (CommDev.mode = 2), the "all-reduce" adds them in rank order, k_scalars finishes on every rank; the
// update kernel pushes its boundary planes straight into the neighbours' ghost planes (HALO_STORE) and the SpMV
// kernel keeps the ghost planes of p up to date (ghost_store).  Checks the slab / wrap / ghost logic of the
// multi-GPU path without GPUs; the mailbox all-reduce itself needs real concurrency and is covered on hardware.
EMU_API int emu_stencil_cg_ranks(int nranks, const int64_t *n, const int *per, const double *dx, const double *dy,
                                 const double *dz, double dt, int jacobi, int has_const, double rtol, double atol, int max_it,
                                 int tile, int kz, const double *b, double *x_out, double *hist, int hist_cap, int *nhist,
                                 int *its, int *reason)
{
    struct Rank
    {
        Problem P;
        Ws W;
        std::vector<double> r, p0, p1, w, x, dinv, hist, sendrecv;
        DevState st{};
        int64_t lo = 0, hi = 0;
    };
    const int64_t nx = n[0], ny = n[1], nz = n[2];
    std::vector<Rank> R((size_t)nranks);
    for (int q = 0; q < nranks; ++q)
    {
        Rank &k = R[(size_t)q];
        build(k.P, 3, n, per, dx, dy, dz, dt);
        const int64_t base = nz / nranks, rem = nz % nranks;
        k.lo = q * base + (q < rem ? q : rem);
        k.hi = k.lo + base + (q < rem ? 1 : 0);
        GridDev &g = k.P.g;
        g.nzl = (int)(k.hi - k.lo);
        g.kz0 = (int)k.lo;
        g.perz_wrap = 0;
        g.wrapz_lo = (per[2] && k.lo == 0) ? 1 : 0;
        g.wrapz_hi = (per[2] && k.hi == nz) ? 1 : 0;
        k.P.vec_elems = (size_t)g.plane * (size_t)(g.nzl + 2);
        k.P.nlocal = nx * ny * (k.hi - k.lo);
        const size_t ve = k.P.vec_elems;
        k.r.assign(ve, 0.0); k.p0.assign(ve, 0.0); k.p1.assign(ve, 0.0); k.w.assign(ve, 0.0); k.x.assign(ve, 0.0);
        k.hist.assign((size_t)hist_cap, 0.0);
        k.sendrecv.assign(16, 0.0);
        if (jacobi)
        {
            k.dinv.assign(ve, 0.0);
            emu::launch(dim3(2), dim3(256), 0, [&] { k_jacobi_setup(k.P.g, k.dinv.data()); });
        }
        k.W.cm.rank = q;
        k.W.cm.nranks = nranks;
        k.W.cm.mode = 2;
        k.W.cm.sendbuf = k.sendrecv.data();
    }
    // neighbours' ghost planes of r (periodic wrap closes the ring)
    for (int q = 0; q < nranks; ++q)
    {
        int dn = q - 1, up = q + 1;
        if (dn < 0) dn = per[2] ? nranks - 1 : -1;
        if (up >= nranks) up = per[2] ? 0 : -1;
        Rank &k = R[(size_t)q];
        k.W.cm.r_ghost_up = up >= 0 ? R[(size_t)up].r.data() : nullptr;
        k.W.cm.r_ghost_dn = dn >= 0 ? R[(size_t)dn].r.data() + (size_t)(R[(size_t)dn].P.g.nzl + 1) * (size_t)k.P.g.plane : nullptr;
    }
    SolveConsts kc{};
    kc.rtol = rtol; kc.atol = atol; kc.divtol = 1e4;
    kc.nglobal = (double)(nx * ny * nz);
    kc.max_it = max_it; kc.norm_type = 1; kc.has_const = has_const; kc.hist_cap = hist_cap;
    auto allreduce_and_finish = [&](int kind) {
        double S[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (auto &k : R)
            for (int s = 0; s < 8; ++s) S[s] += k.sendrecv[(size_t)s];
        for (auto &k : R)
        {
            for (int s = 0; s < 8; ++s) k.sendrecv[8 + (size_t)s] = S[s];
            emu::launch(dim3(1), dim3(32), 0, [&] { k_scalars(kind, k.sendrecv.data() + 8, &k.st, kc, k.hist.data()); });
        }
    };
    auto update_all = [&](bool init, int kind) {
        for (auto &k : R)
        {
            UpdVecs uv{k.r.data(), k.w.data(), jacobi ? k.dinv.data() : nullptr, 1};
            const GridDev &g = k.P.g;
#define EMU_UPD(JAC, INIT, PAD) emu::launch(dim3(getenv("EMU_UPD_BLOCKS") ? atoi(getenv("EMU_UPD_BLOCKS")) : 3), dim3(256), 0, [&] { k_update2<JAC, INIT, PAD, true, 4>(g, uv, kind, k.W.ws, k.W.cm, &k.st, kc, k.hist.data()); })
            const bool pad = g.px != g.nx;
            if (jacobi) { if (init) { if (pad) EMU_UPD(true, true, true); else EMU_UPD(true, true, false); } else { if (pad) EMU_UPD(true, false, true); else EMU_UPD(true, false, false); } }
            else { if (init) { if (pad) EMU_UPD(false, true, true); else EMU_UPD(false, true, false); } else { if (pad) EMU_UPD(false, false, true); else EMU_UPD(false, false, false); } }
#undef EMU_UPD
        }
        allreduce_and_finish(kind);
    };
    for (auto &k : R)
    {
        emu::launch(dim3(1), dim3(32), 0, [&] { k_state_reset(&k.st); });
        const double *bl = b + (size_t)(nx * ny * k.lo);
        emu::launch(dim3(2), dim3(256), 0, [&] { k_scatter(k.P.g, bl, k.r.data(), k.x.data()); });
    }
    if (has_const) update_all(true, FIN_INIT_CENTRE);
    update_all(true, FIN_INIT);
    for (int it = 0; it < max_it + 2 && !R[0].st.done; ++it)
    {
        for (auto &k : R)
        {
            double *pp[2] = {k.p0.data(), k.p1.data()};
            VecSet v{k.r.data(), pp[it & 1], pp[(it & 1) ^ 1], k.w.data(), k.x.data(), jacobi ? k.dinv.data() : nullptr};
            const GridDev &g = k.P.g;
            const int kzz = kz > 0 ? kz : g.nzl;
            using L8 = Spmv2Smem<32, 8, 4, false, false>;
            using L8J = Spmv2Smem<32, 8, 4, true, false>;
            dim3 grid((unsigned)((g.nx + 63) / 64), (unsigned)((g.ny + 5) / 6), (unsigned)((g.nzl + kzz - 1) / kzz));
            const bool periodic = per[0] || per[1] || per[2];
            (void)tile;
#define EMU_SPMV(JAC, PER, LL) emu::launch(grid, dim3(32, 8), LL::total(kzz), [&] { k_spmv2<32, 8, 4, 3, JAC, false, PER>(g, v, kzz, k.W.ws, k.W.cm, &k.st, kc, k.hist.data(), 1); })
            if (jacobi) { if (periodic) EMU_SPMV(true, true, L8J); else EMU_SPMV(true, false, L8J); }
            else { if (periodic) EMU_SPMV(false, true, L8); else EMU_SPMV(false, false, L8); }
#undef EMU_SPMV
        }
        allreduce_and_finish(FIN_SPMV);
        update_all(false, FIN_UPDATE);
    }
    for (auto &k : R)
    {
        emu::launch(dim3(2), dim3(256), 0, [&] { k_xtail(k.P.g, k.x.data(), k.p0.data(), k.p1.data(), &k.st); });
        double *xl = x_out + (size_t)(nx * ny * k.lo);
        emu::launch(dim3(2), dim3(256), 0, [&] { k_gather(k.P.g, k.x.data(), xl); });
    }
    // every rank must hold the same scalars
    for (auto &k : R)
        if (k.st.nhist != R[0].st.nhist || k.st.reason != R[0].st.reason || memcmp(k.hist.data(), R[0].hist.data(), sizeof(double) * (size_t)std::min(R[0].st.nhist, hist_cap)) != 0)
            return 2;
    *nhist = R[0].st.nhist;
    *its = R[0].st.its;
    *reason = R[0].st.reason;
    memcpy(hist, R[0].hist.data(), sizeof(double) * (size_t)std::min(R[0].st.nhist, hist_cap));
    return R[0].st.done ? 0 : 1;
}

// ---- general CSR operator: CG (no / constant / one explicit null-space vector) and BiCGStab, mirroring
// csr_solver.inc (csr_cg / csr_bcgs)
EMU_API int emu_csr_solve(int64_t n, const int64_t *rowptr, const int32_t *col, const double *val, int bcgs, int jacobi,
                          int has_const, const double *nullvec, double rtol, double atol, int max_it, const double *b,
                          double *x_out, double *hist, int hist_cap, int *nhist, int *its, int *reason)
{
    Ws W;
    CsrDev A{n, rowptr, col, val};
    std::vector<double> dinv((size_t)n, 1.0);
    for (int64_t i = 0; i < n; ++i)
    {
        double d = 0.0;
        for (int64_t q = rowptr[i]; q < rowptr[i + 1]; ++q)
            if (col[q] == i) d = val[q];
        dinv[(size_t)i] = d != 0.0 ? 1.0 / d : 1.0;
    }
    SolveConsts kc{};
    kc.rtol = rtol; kc.atol = atol; kc.divtol = 1e4; kc.nglobal = (double)n;
    kc.max_it = max_it; kc.norm_type = 1; kc.has_const = has_const; kc.hist_cap = hist_cap;
    DevState st{};
    emu::launch(dim3(1), dim3(32), 0, [&] { k_state_reset(&st); });
    std::vector<double> r((size_t)n), p0((size_t)n, 0.0), p1((size_t)n, 0.0), w((size_t)n, 0.0), x((size_t)n, 0.0);
    const int blocks = 3;
    const double *dv = jacobi ? dinv.data() : nullptr;
    if (bcgs)
    {
        std::vector<double> rp((size_t)n), vv((size_t)n), s((size_t)n), t((size_t)n);
        if (jacobi) emu::launch(dim3(blocks), dim3(256), 0, [&] { k_bcgs_init<true>(n, b, dv, r.data(), rp.data(), p0.data(), vv.data(), W.ws, &st, kc, hist); });
        else emu::launch(dim3(blocks), dim3(256), 0, [&] { k_bcgs_init<false>(n, b, dv, r.data(), rp.data(), p0.data(), vv.data(), W.ws, &st, kc, hist); });
        for (int it = 0; it < max_it + 2 && !st.done; ++it)
        {
            emu::launch(dim3(blocks), dim3(256), 0, [&] { k_bcgs_p(n, r.data(), vv.data(), p0.data(), &st); });
            if (jacobi) emu::launch(dim3(blocks), dim3(256), 0, [&] { k_bcgs_spmv1<true>(A, p0.data(), dv, rp.data(), vv.data(), W.ws, &st, kc, hist); });
            else emu::launch(dim3(blocks), dim3(256), 0, [&] { k_bcgs_spmv1<false>(A, p0.data(), dv, rp.data(), vv.data(), W.ws, &st, kc, hist); });
            if (jacobi) emu::launch(dim3(blocks), dim3(256), 0, [&] { k_bcgs_spmv2<true>(A, r.data(), vv.data(), dv, s.data(), t.data(), W.ws, &st, kc, hist); });
            else emu::launch(dim3(blocks), dim3(256), 0, [&] { k_bcgs_spmv2<false>(A, r.data(), vv.data(), dv, s.data(), t.data(), W.ws, &st, kc, hist); });
            emu::launch(dim3(blocks), dim3(256), 0, [&] { k_bcgs_upd(n, p0.data(), s.data(), t.data(), rp.data(), x.data(), r.data(), W.ws, &st, kc, hist); });
        }
        emu::launch(dim3(blocks), dim3(256), 0, [&] { k_bcgs_tail(n, p0.data(), x.data(), &st); });
    }
    else
    {
        memcpy(r.data(), b, sizeof(double) * (size_t)n);
        const int nm = nullvec ? 2 : (has_const ? 1 : 0);
        auto upd = [&](bool init, int kind) {
#define EMU_CU(JAC, NM, INIT) emu::launch(dim3(blocks), dim3(256), 0, [&] { k_csr_cg_update<JAC, NM, INIT>(n, r.data(), w.data(), dv, nullvec, kind, W.ws, &st, kc, hist); })
            if (jacobi) { if (nm == 2) { if (init) EMU_CU(true, 2, true); else EMU_CU(true, 2, false); } else if (nm == 1) { if (init) EMU_CU(true, 1, true); else EMU_CU(true, 1, false); } else { if (init) EMU_CU(true, 0, true); else EMU_CU(true, 0, false); } }
            else { if (nm == 2) { if (init) EMU_CU(false, 2, true); else EMU_CU(false, 2, false); } else if (nm == 1) { if (init) EMU_CU(false, 1, true); else EMU_CU(false, 1, false); } else { if (init) EMU_CU(false, 0, true); else EMU_CU(false, 0, false); } }
#undef EMU_CU
        };
        if (nm == 2) upd(true, FIN_CSR_INIT);
        else
        {
            if (nm == 1) upd(true, FIN_INIT_CENTRE);
            upd(true, FIN_INIT);
        }
        double *pp[2] = {p0.data(), p1.data()};
        for (int it = 0; it < max_it + 2 && !st.done; ++it)
        {
            CsrVecs v{r.data(), pp[it & 1], pp[(it & 1) ^ 1], w.data(), x.data(), dv, nullvec};
#define EMU_CS(JAC, NM) emu::launch(dim3(blocks), dim3(256), 0, [&] { k_csr_cg_spmv<JAC, NM>(A, v, W.ws, &st, kc, hist); })
            if (jacobi) { if (nm == 2) EMU_CS(true, 2); else if (nm == 1) EMU_CS(true, 1); else EMU_CS(true, 0); }
            else { if (nm == 2) EMU_CS(false, 2); else if (nm == 1) EMU_CS(false, 1); else EMU_CS(false, 0); }
#undef EMU_CS
            upd(false, nm == 2 ? (int)FIN_CSR_UPDATE : (int)FIN_UPDATE);
        }
        GridDev g1{};
        g1.nx = (int)n; g1.ny = 1; g1.nzl = 1; g1.px = (int)n; g1.plane = 0;
        emu::launch(dim3(1), dim3(256), 0, [&] { k_xtail(g1, x.data(), p0.data(), p1.data(), &st); });
    }
    memcpy(x_out, x.data(), sizeof(double) * (size_t)n);
    *nhist = st.nhist;
    *its = st.its;
    *reason = st.reason;
    return st.done ? 0 : 1;
}

// ---- line-coefficient operator (sep_kernels.cuh): same Krylov drivers as emu_csr_solve with the row product swapped.
// The structure comes from b200ls_staggered_analyze (host code of libb200ls.so), passed in as plain arrays.
}  // extern "C"

namespace {
// tiled kernels of sep_tile.cuh instead of the row-per-thread ones: 0 = off, 2 / 4 = tile of 64 / 128 cells in x
int g_sep_xr = 0, g_sep_zchunk = 0, g_sep_target = 0, g_sep_stages = 3;
}  // namespace
extern "C" EMU_API void emu_set_sep_tile(int xr, int zchunk, int target_blocks)
{
    g_sep_xr = xr;
    g_sep_zchunk = zchunk;
    g_sep_target = target_blocks;
}
extern "C" EMU_API void emu_set_sep_stages(int stages) { g_sep_stages = stages == 4 ? 4 : 3; }
// launch geometry of the tiled kernels (sep_tile_plan, host code shared with sep_solver.inc) for the tests of its rules:
// out = {stencil_blocks, surf_blocks, tail_blocks, surf_cells, then per field: tiles_x, tiles, nchunk, zchunk, block0, surf0}
extern "C" EMU_API void emu_sep_tile_plan(int nfields, const int64_t *dims, const int *periodic, int64_t nrows, int xr, int zchunk,
                                          int target_blocks, int *out)
{
    SepDev A{};
    A.nf = nfields;
    A.nrows = nrows;
    long long off = 0;
    for (int f = 0; f < nfields; ++f)
    {
        SepField &F = A.f[f];
        F.n0 = (int)dims[3 * f]; F.n1 = (int)dims[3 * f + 1]; F.n2 = (int)dims[3 * f + 2];
        F.per0 = periodic[0] && F.n0 >= 3; F.per1 = periodic[1] && F.n1 >= 3; F.per2 = periodic[2] && F.n2 >= 3;
        F.off = off;
        off += (long long)F.n0 * F.n1 * F.n2;
    }
    A.nsep = off;
    const SepTilePlan T = sep_tile_plan(A, xr, zchunk, target_blocks);
    out[0] = T.stencil_blocks; out[1] = T.surf_blocks; out[2] = T.tail_blocks; out[3] = T.surf_cells;
    for (int f = 0; f < nfields; ++f)
    {
        int *o = out + 4 + 6 * f;
        o[0] = T.f[f].tiles_x; o[1] = T.f[f].tiles; o[2] = T.f[f].nchunk; o[3] = T.f[f].zchunk; o[4] = T.f[f].block0; o[5] = T.f[f].surf0;
    }
}
namespace {
SepDev make_sep(int nfields, const int64_t *dims, const int *periodic, const double *widths, int64_t n, const double *coef,
                const double *diag, const int64_t *rem_rowptr, const int32_t *rem_col, const double *rem_val)
{
    SepDev A{};
    A.nf = nfields;
    A.nrows = n;
    A.diag = diag;
    long long off = 0, cpos = 0;
    for (int f = 0; f < nfields; ++f)
    {
        SepField &F = A.f[f];
        F.n0 = (int)dims[3 * f]; F.n1 = (int)dims[3 * f + 1]; F.n2 = (int)dims[3 * f + 2];
        F.per0 = periodic[0] && F.n0 >= 3; F.per1 = periodic[1] && F.n1 >= 3; F.per2 = periodic[2] && F.n2 >= 3;
        F.off = off;
        off += (long long)F.n0 * F.n1 * F.n2;
        long long wpos = 0;
        for (int d = 0; d < 3; ++d)
        {
            if (widths)
            {
                F.w[d] = widths + wpos;
                F.cm[d] = coef + cpos;
                F.cp[d] = coef + cpos + 1;
                wpos += dims[3 * f + d];
                cpos += dims[3 * f + d] + 1;
            }
            else
            {
                F.w[d] = nullptr;
                F.cm[d] = coef + cpos;
                F.cp[d] = coef + cpos + dims[3 * f + d];
                cpos += 2 * dims[3 * f + d];
            }
        }
    }
    A.nsep = off;
    if (rem_rowptr && rem_rowptr[n] > 0)
    {
        A.rem_rowptr = rem_rowptr;
        A.rem_col = rem_col;
        A.rem_val = rem_val;
    }
    return A;
}
}  // namespace

extern "C" {

// widths != nullptr: "hybrid" form (b200ls_set_poisson_hybrid) -- one field, coef = face arrays gx (n0+1) | gy | gz,
// widths = dx | dy | dz, coefficient = (product of the two other widths) * g
EMU_API int emu_sep_solve(int nfields, const int64_t *dims, const int *periodic, const double *widths, int64_t n, const double *coef, const double *diag,
                          const int64_t *rem_rowptr, const int32_t *rem_col, const double *rem_val, const double *dinv_in,
                          int mode /* 0 apply, 1 cg, 2 bcgs */, int jacobi, int has_const, const double *nullvec, double rtol,
                          double atol, int max_it, const double *b, double *x_out, double *hist, int hist_cap, int *nhist,
                          int *its, int *reason)
{
    Ws W;
    const SepDev A = make_sep(nfields, dims, periodic, widths, n, coef, diag, rem_rowptr, rem_col, rem_val);
    const int blocks = 3;
    // the tiled kernels (sep_solver.inc: sep_tile_xr): grid and plan as on the device
    const int xr = g_sep_xr;
    const bool hyb = widths != nullptr;
    const SepTilePlan T = sep_tile_plan(A, xr ? xr : 2, g_sep_zchunk, g_sep_target > 0 ? g_sep_target : 24);
    const dim3 tgrid((unsigned)T.blocks());
    if (mode == 0)
    {
#define EMU_TA(XR, HYB, R) emu::launch(tgrid, dim3(256), sep_tile_smem_bytes<XR, R, SepOpApply>(), [&] { k_sep_tile_apply<XR, HYB, R>(A, T, b, x_out); })
#define EMU_TAR(XR, HYB) do { if (g_sep_stages == 4) EMU_TA(XR, HYB, 4); else EMU_TA(XR, HYB, 3); } while (0)
        if (xr == 2 && hyb) EMU_TAR(2, true);
        else if (xr == 2) EMU_TAR(2, false);
        else if (xr == 4 && hyb) EMU_TAR(4, true);
        else if (xr == 4) EMU_TAR(4, false);
#undef EMU_TA
#undef EMU_TAR
        else emu::launch(dim3(blocks), dim3(256), 0, [&] { k_sep_apply(A, b, x_out); });
        return 0;
    }
    SolveConsts kc{};
    kc.rtol = rtol; kc.atol = atol; kc.divtol = 1e4; kc.nglobal = (double)n;
    kc.max_it = max_it; kc.norm_type = 1; kc.has_const = has_const; kc.hist_cap = hist_cap;
    DevState st{};
    emu::launch(dim3(1), dim3(32), 0, [&] { k_state_reset(&st); });
    std::vector<double> r((size_t)n), p0((size_t)n, 0.0), p1((size_t)n, 0.0), w((size_t)n, 0.0), x((size_t)n, 0.0);
    const double *dv = jacobi ? dinv_in : nullptr;
    if (mode == 2 && xr && !hyb)
    {
        std::vector<double> rp((size_t)n), vv((size_t)n), s((size_t)n), t((size_t)n);
        if (jacobi) emu::launch(dim3(blocks), dim3(256), 0, [&] { k_bcgs_init<true>(n, b, dv, r.data(), rp.data(), p0.data(), vv.data(), W.ws, &st, kc, hist); });
        else emu::launch(dim3(blocks), dim3(256), 0, [&] { k_bcgs_init<false>(n, b, dv, r.data(), rp.data(), p0.data(), vv.data(), W.ws, &st, kc, hist); });
        for (int it = 0; it < max_it + 2 && !st.done; ++it)
        {
            emu::launch(dim3(blocks), dim3(256), 0, [&] { k_bcgs_p(n, r.data(), vv.data(), p0.data(), &st); });
#define EMU_T1(XR, R, JAC) emu::launch(tgrid, dim3(256), sep_tile_smem_bytes<XR, R, SepOpBcgs1<JAC>>(), [&] { k_sep_tile_bcgs_spmv1<XR, R, JAC>(A, T, p0.data(), dv, rp.data(), vv.data(), W.ws, &st, kc, hist); })
#define EMU_T2(XR, R, JAC) emu::launch(tgrid, dim3(256), sep_tile_smem_bytes<XR, R, SepOpBcgs2<JAC>>(), [&] { k_sep_tile_bcgs_spmv2<XR, R, JAC>(A, T, r.data(), vv.data(), dv, s.data(), t.data(), W.ws, &st, kc, hist); })
#define EMU_T1R(XR, JAC) do { if (g_sep_stages == 4) EMU_T1(XR, 4, JAC); else EMU_T1(XR, 3, JAC); } while (0)
#define EMU_T2R(XR, JAC) do { if (g_sep_stages == 4) EMU_T2(XR, 4, JAC); else EMU_T2(XR, 3, JAC); } while (0)
            if (xr == 2) { if (jacobi) EMU_T1R(2, true); else EMU_T1R(2, false); }
            else { if (jacobi) EMU_T1R(4, true); else EMU_T1R(4, false); }
            if (xr == 2) { if (jacobi) EMU_T2R(2, true); else EMU_T2R(2, false); }
            else { if (jacobi) EMU_T2R(4, true); else EMU_T2R(4, false); }
#undef EMU_T1R
#undef EMU_T2R
#undef EMU_T1
#undef EMU_T2
            emu::launch(dim3(blocks), dim3(256), 0, [&] { k_bcgs_upd(n, p0.data(), s.data(), t.data(), rp.data(), x.data(), r.data(), W.ws, &st, kc, hist); });
        }
        emu::launch(dim3(blocks), dim3(256), 0, [&] { k_bcgs_tail(n, p0.data(), x.data(), &st); });
    }
    else if (mode == 2)
    {
        std::vector<double> rp((size_t)n), vv((size_t)n), s((size_t)n), t((size_t)n);
        if (jacobi) emu::launch(dim3(blocks), dim3(256), 0, [&] { k_bcgs_init<true>(n, b, dv, r.data(), rp.data(), p0.data(), vv.data(), W.ws, &st, kc, hist); });
        else emu::launch(dim3(blocks), dim3(256), 0, [&] { k_bcgs_init<false>(n, b, dv, r.data(), rp.data(), p0.data(), vv.data(), W.ws, &st, kc, hist); });
        for (int it = 0; it < max_it + 2 && !st.done; ++it)
        {
            emu::launch(dim3(blocks), dim3(256), 0, [&] { k_bcgs_p(n, r.data(), vv.data(), p0.data(), &st); });
            if (jacobi) emu::launch(dim3(blocks), dim3(256), 0, [&] { k_sep_bcgs_spmv1<true>(A, p0.data(), dv, rp.data(), vv.data(), W.ws, &st, kc, hist); });
            else emu::launch(dim3(blocks), dim3(256), 0, [&] { k_sep_bcgs_spmv1<false>(A, p0.data(), dv, rp.data(), vv.data(), W.ws, &st, kc, hist); });
            if (jacobi) emu::launch(dim3(blocks), dim3(256), 0, [&] { k_sep_bcgs_spmv2<true>(A, r.data(), vv.data(), dv, s.data(), t.data(), W.ws, &st, kc, hist); });
            else emu::launch(dim3(blocks), dim3(256), 0, [&] { k_sep_bcgs_spmv2<false>(A, r.data(), vv.data(), dv, s.data(), t.data(), W.ws, &st, kc, hist); });
            emu::launch(dim3(blocks), dim3(256), 0, [&] { k_bcgs_upd(n, p0.data(), s.data(), t.data(), rp.data(), x.data(), r.data(), W.ws, &st, kc, hist); });
        }
        emu::launch(dim3(blocks), dim3(256), 0, [&] { k_bcgs_tail(n, p0.data(), x.data(), &st); });
    }
    else
    {
        memcpy(r.data(), b, sizeof(double) * (size_t)n);
        const int nm = nullvec ? 2 : (has_const ? 1 : 0);
        auto upd = [&](bool init, int kind) {
#define EMU_CU(JAC, NM, INIT) emu::launch(dim3(blocks), dim3(256), 0, [&] { k_csr_cg_update<JAC, NM, INIT>(n, r.data(), w.data(), dv, nullvec, kind, W.ws, &st, kc, hist); })
            if (jacobi) { if (nm == 2) { if (init) EMU_CU(true, 2, true); else EMU_CU(true, 2, false); } else if (nm == 1) { if (init) EMU_CU(true, 1, true); else EMU_CU(true, 1, false); } else { if (init) EMU_CU(true, 0, true); else EMU_CU(true, 0, false); } }
            else { if (nm == 2) { if (init) EMU_CU(false, 2, true); else EMU_CU(false, 2, false); } else if (nm == 1) { if (init) EMU_CU(false, 1, true); else EMU_CU(false, 1, false); } else { if (init) EMU_CU(false, 0, true); else EMU_CU(false, 0, false); } }
#undef EMU_CU
        };
        if (nm == 2) upd(true, FIN_CSR_INIT);
        else
        {
            if (nm == 1) upd(true, FIN_INIT_CENTRE);
            upd(true, FIN_INIT);
        }
        double *pp[2] = {p0.data(), p1.data()};
        for (int it = 0; it < max_it + 2 && !st.done; ++it)
        {
            CsrVecs v{r.data(), pp[it & 1], pp[(it & 1) ^ 1], w.data(), x.data(), dv, nullvec};
#define EMU_CS(JAC, NM) emu::launch(dim3(blocks), dim3(256), 0, [&] { k_sep_cg_spmv<JAC, NM>(A, v, W.ws, &st, kc, hist); })
#define EMU_CTR(XR, HYB, R, JAC, NM) emu::launch(tgrid, dim3(256), sep_tile_smem_bytes<XR, R, SepOpCg<JAC, NM>>(), [&] { k_sep_tile_cg_spmv<XR, HYB, R, JAC, NM>(A, T, v, W.ws, &st, kc, hist); })
#define EMU_CT(XR, HYB, JAC, NM) do { if (g_sep_stages == 4) EMU_CTR(XR, HYB, 4, JAC, NM); else EMU_CTR(XR, HYB, 3, JAC, NM); } while (0)
#define EMU_CTN(XR, HYB, JAC) do { if (nm == 2) EMU_CT(XR, HYB, JAC, 2); else if (nm == 1) EMU_CT(XR, HYB, JAC, 1); else EMU_CT(XR, HYB, JAC, 0); } while (0)
#define EMU_CTJ(XR, HYB) do { if (jacobi) EMU_CTN(XR, HYB, true); else EMU_CTN(XR, HYB, false); } while (0)
            if (xr == 2) { if (hyb) EMU_CTJ(2, true); else EMU_CTJ(2, false); }
            else if (xr == 4) { if (hyb) EMU_CTJ(4, true); else EMU_CTJ(4, false); }
            else if (jacobi) { if (nm == 2) EMU_CS(true, 2); else if (nm == 1) EMU_CS(true, 1); else EMU_CS(true, 0); }
            else { if (nm == 2) EMU_CS(false, 2); else if (nm == 1) EMU_CS(false, 1); else EMU_CS(false, 0); }
#undef EMU_CS
#undef EMU_CT
#undef EMU_CTR
#undef EMU_CTN
#undef EMU_CTJ
            upd(false, nm == 2 ? (int)FIN_CSR_UPDATE : (int)FIN_UPDATE);
        }
        GridDev g1{};
        g1.nx = (int)n; g1.ny = 1; g1.nzl = 1; g1.px = (int)n; g1.plane = 0;
        emu::launch(dim3(1), dim3(256), 0, [&] { k_xtail(g1, x.data(), p0.data(), p1.data(), &st); });
    }
    memcpy(x_out, x.data(), sizeof(double) * (size_t)n);
    *nhist = st.nhist;
    *its = st.its;
    *reason = st.reason;
    return st.done ? 0 : 1;
}

}  // extern "C"

// ---- geometric multigrid preconditioner (mg_kernels.cuh) through the SHARED schedule of mg_schedule.h: the launcher
// below runs the kernels under the emulation instead of enqueueing them on a stream (MgCudaLauncher in mg_solver.inc).
namespace {
struct MgEmu
{
    std::vector<MgHostLevel> host;
    std::vector<MgLevel> dev;
    std::vector<std::vector<double>> axes;
    std::vector<std::vector<std::vector<double>>> bufs;
    MgParams prm;
    DevState *st = nullptr;
    int blocks = 3;
    // "mg_fuse" (see MgCudaLauncher)
    bool fuse_rupdate = false, sums_done = false;
    int fuse_sums_kind = -1;
    double *fuse_r = nullptr;
    const double *fuse_w = nullptr;
    Ws *fuse_W = nullptr;
    SolveConsts fuse_kc{};
    double *fuse_hist = nullptr;
    void first(int l, const double *b, double *dout, double inv_theta)
    {
        const MgLevel L = dev[(size_t)l];
        if (l == 0 && fuse_rupdate) emu::launch(dim3(blocks), dim3(256), 0, [&] { k_mg_first_rupd(L, fuse_r, fuse_w, dout, inv_theta, st); });
        else emu::launch(dim3(blocks), dim3(256), 0, [&] { k_mg_cheb_first(L, b, dout, inv_theta, st); });
    }
    void step(int l, bool xzero, bool, bool prolong, bool last, const double *b, const double *xin, const double *din,
              const double *ec, double *xout, double *dout, double c1, double c2)
    {
        const MgLevel L = dev[(size_t)l], Lc = dev[(size_t)std::min<int>(l + 1, (int)dev.size() - 1)];
        if (l == 0 && last && !xzero && fuse_sums_kind >= 0)
        {
            const int kind = fuse_sums_kind;
            if (prolong) emu::launch(dim3(blocks), dim3(256), 0, [&] { k_mg_last_sums<true>(L, Lc, b, xin, din, ec, xout, c1, c2, kind, fuse_W->ws, fuse_W->cm, st, fuse_kc, fuse_hist); });
            else emu::launch(dim3(blocks), dim3(256), 0, [&] { k_mg_last_sums<false>(L, Lc, b, xin, din, ec, xout, c1, c2, kind, fuse_W->ws, fuse_W->cm, st, fuse_kc, fuse_hist); });
            sums_done = true;
            return;
        }
#define EMU_MG(XZ, DZ, PR, LA) emu::launch(dim3(blocks), dim3(256), 0, [&] { k_mg_cheb_step<XZ, DZ, PR, LA>(L, Lc, b, xin, din, ec, xout, dout, c1, c2, st); })
        if (prolong) { if (last) EMU_MG(false, true, true, true); else EMU_MG(false, true, true, false); }
        else if (xzero) { if (last) EMU_MG(true, false, false, true); else EMU_MG(true, false, false, false); }
        else { if (last) EMU_MG(false, false, false, true); else EMU_MG(false, false, false, false); }
#undef EMU_MG
    }
    void restrict(int l, bool xzero, const double *b, const double *xin, const double *din, double *xsum, double *bc)
    {
        const MgLevel L = dev[(size_t)l], Lc = dev[(size_t)l + 1];
        if (xzero) emu::launch(dim3(blocks), dim3(256), 0, [&] { k_mg_restrict<true>(L, Lc, b, xin, din, xsum, bc, st); });
        else emu::launch(dim3(blocks), dim3(256), 0, [&] { k_mg_restrict<false>(L, Lc, b, xin, din, xsum, bc, st); });
    }
    std::vector<MgOp> tail_ops;
    double *tail_result = nullptr;
    double *tail()
    {
        const MgLevel *lv = dev.data();
        const MgOp *ops = tail_ops.data();
        const int nops = (int)tail_ops.size();
        emu::launch(dim3(1), dim3(512), 0, [&] { k_mg_tail(lv, ops, nops, st); });
        return tail_result;
    }
    void pointers(double *(*work)[4], double **rhs)
    {
        for (size_t l = 0; l < dev.size(); ++l)
        {
            for (int q = 0; q < 4; ++q) work[l][q] = bufs[l][(size_t)q].data();
            rhs[l] = l > 0 ? bufs[l][4].data() : nullptr;
        }
    }
    // tail_cells > 0: the levels with at most that many cells run as one launch (record_mg_tail of mg_solver.inc)
    void record_tail(int64_t tail_cells)
    {
        prm.tail_level = -1;
        const int tl = tail_cells > 0 ? mg_tail_level(host, tail_cells) : -1;
        if (tl < 1) return;
        double *work[32][4];
        double *rhs[32];
        pointers(work, rhs);
        MgProgramRecorder rec;
        tail_result = mg_cycle(tl, (int)dev.size(), rhs[tl], work, rhs, prm, rec);
        tail_ops = rec.ops;
        prm.tail_level = tl;
    }
    double *cycle(const double *r)
    {
        double *work[32][4];
        double *rhs[32];
        pointers(work, rhs);
        return mg_cycle(0, (int)dev.size(), r, work, rhs, prm, *this);
    }
};

void mg_setup(MgEmu &M, const Problem &P, int dim, const int64_t *n, double dt, int max_levels, int smooth_its, int coarse_its)
{
    const GridDev &g = P.g;
    const int64_t n3[3] = {n[0], n[1], dim == 3 ? n[2] : 1};
    std::vector<double> dx(g.dx, g.dx + n3[0]), dy(g.dy, g.dy + n3[1]), dz(g.dz, g.dz + n3[2]);
    std::vector<double> gx(g.gx, g.gx + n3[0] + 1), gy(g.gy, g.gy + n3[1] + 1), gz(g.gz, g.gz + n3[2] + 1);
    M.host = mg_build_hierarchy(n3, P.per, dx, dy, dz, gx, gy, gz, dt, max_levels);
    const size_t nl = M.host.size();
    M.dev.resize(nl);
    M.axes.resize(nl);
    M.bufs.resize(nl);
    for (size_t l = 0; l < nl; ++l)
    {
        const MgHostLevel &H = M.host[l];
        MgLevel &D = M.dev[l];
        D.nx = H.n[0]; D.ny = H.n[1]; D.nz = H.n[2];
        D.perx = H.per[0]; D.pery = H.per[1]; D.perz = H.per[2];
        D.mx = D.my = D.mz = D.sx = D.sy = D.sz = nullptr;
        if (l + 1 < nl)
        {
            D.mx = H.cmap[0].data(); D.my = H.cmap[1].data(); D.mz = H.cmap[2].data();
            D.sx = H.cstart[0].data(); D.sy = H.cstart[1].data(); D.sz = H.cstart[2].data();
        }
        size_t elems;
        if (l == 0)
        {
            D.px = g.px; D.plane = g.plane; D.base = g.plane;
            D.dx = g.dx; D.dy = g.dy; D.dz = g.dz; D.gx = g.gx; D.gy = g.gy; D.gz = g.gz;
            elems = P.vec_elems;
        }
        else
        {
            D.px = D.nx; D.plane = (long long)D.nx * D.ny; D.base = 0;
            size_t off[6];
            const std::vector<double> *parts[6] = {&H.d[0], &H.d[1], &H.d[2], &H.g[0], &H.g[1], &H.g[2]};
            for (int q = 0; q < 6; ++q)
            {
                off[q] = M.axes[l].size();
                M.axes[l].insert(M.axes[l].end(), parts[q]->begin(), parts[q]->end());
            }
            const double *base = M.axes[l].data();
            D.dx = base + off[0]; D.dy = base + off[1]; D.dz = base + off[2];
            D.gx = base + off[3]; D.gy = base + off[4]; D.gz = base + off[5];
            elems = (size_t)H.cells();
        }
        M.bufs[l].assign(5, std::vector<double>(elems, 0.0));
    }
    M.prm.smooth_its = smooth_its;
    M.prm.coarse_its = coarse_its;
}
}  // namespace

extern "C" {

// mode 0: x_out = M^-1 b (one V-cycle);  mode 1: KSPSolve_CG preconditioned with the V-cycle (solve_stencil_pcg_mg)
EMU_API int emu_mg(int dim, const int64_t *n, const int *per, const double *dx, const double *dy, const double *dz, double dt,
                   int mode, int has_const, double rtol, double atol, int max_it, int max_levels, int smooth_its, int coarse_its,
                   int tile, int tail_cells, int fuse, const double *b, double *x_out, double *hist, int hist_cap, int *nhist, int *its, int *reason,
                   int *nlevels)
{
    Problem P;
    build(P, dim, n, per, dx, dy, dz, dt);
    Ws W;
    MgEmu M;
    mg_setup(M, P, dim, n, dt, max_levels, smooth_its, coarse_its);
    M.record_tail(tail_cells);
    *nlevels = (int)M.dev.size();
    const size_t ve = P.vec_elems;
    std::vector<double> r(ve, 0.0), p0(ve, 0.0), p1(ve, 0.0), w(ve, 0.0), x(ve, 0.0);
    SolveConsts kc{};
    kc.rtol = rtol; kc.atol = atol; kc.divtol = 1e4; kc.nglobal = (double)P.nlocal;
    kc.max_it = max_it; kc.norm_type = 1; kc.has_const = has_const; kc.hist_cap = hist_cap;
    DevState st{};
    M.st = &st;
    emu::launch(dim3(1), dim3(32), 0, [&] { k_state_reset(&st); });
    emu::launch(dim3(4), dim3(256), 0, [&] { k_scatter(P.g, b, r.data(), x.data()); });
    double *z = M.cycle(r.data());
    if (mode == 0)
    {
        emu::launch(dim3(4), dim3(256), 0, [&] { k_gather(P.g, z, x_out); });
        return 0;
    }
    auto zsums = [&](int kind) {
        const MgLevel L0 = M.dev[0];
        emu::launch(dim3(3), dim3(256), 0, [&] { k_mg_zsums(L0, z, r.data(), kind, W.ws, W.cm, &st, kc, hist); });
    };
    if (has_const) zsums(FIN_INIT_CENTRE);
    zsums(FIN_INIT);
    double *pp[2] = {p0.data(), p1.data()};
    const long long nflat = (long long)P.g.plane * P.g.nzl;
    for (int it = 0; it < max_it + 2 && !st.done; ++it)
    {
        VecSet v{z, pp[it & 1], pp[(it & 1) ^ 1], w.data(), x.data(), nullptr};
        spmv<false, false>(P, tile, v, P.g.nzl, W, &st, kc, hist);
        if (fuse)
        {
            M.fuse_rupdate = true; M.fuse_sums_kind = FIN_UPDATE; M.sums_done = false;
            M.fuse_r = r.data(); M.fuse_w = w.data(); M.fuse_W = &W; M.fuse_kc = kc; M.fuse_hist = hist;
            z = M.cycle(r.data());
            M.fuse_rupdate = false; M.fuse_sums_kind = -1;
            if (!M.sums_done) zsums(FIN_UPDATE);
            continue;
        }
        emu::launch(dim3(3), dim3(256), 0, [&] { k_mg_rupdate(nflat, r.data() + P.g.plane, w.data() + P.g.plane, &st); });
        z = M.cycle(r.data());
        zsums(FIN_UPDATE);
    }
    emu::launch(dim3(4), dim3(256), 0, [&] { k_xtail(P.g, x.data(), p0.data(), p1.data(), &st); });
    emu::launch(dim3(4), dim3(256), 0, [&] { k_gather(P.g, x.data(), x_out); });
    *nhist = st.nhist;
    *its = st.its;
    *reason = st.reason;
    return st.done ? 0 : 1;
}

// KSPSolve_CG on the hybrid operator preconditioned with diag(V-cycle on the pressure block, 1/diag behind it): mirrors
// sep_pcg_mg of mg_solver.inc.  widths = dx | dy | dz, coef = gx | gy | gz of the pressure grid.
EMU_API int emu_hybrid_mg_pcg(int dim, const int64_t *n3, const int *periodic, const double *widths, double dt, int64_t n,
                              const double *coef, const double *diag, const int64_t *rem_rowptr, const int32_t *rem_col,
                              const double *rem_val, const double *dinv, int has_const, const double *nullvec, double rtol,
                              double atol, int max_it, int smooth_its, int coarse_its, const double *b, double *x_out, double *hist,
                              int hist_cap, int *nhist, int *its, int *reason)
{
    Ws W;
    const SepDev A = make_sep(1, n3, periodic, widths, n, coef, diag, rem_rowptr, rem_col, rem_val);
    // hierarchy of the pressure grid; level 0 in the compact layout of the system's first nsep rows
    MgEmu M;
    {
        std::vector<double> dx(widths, widths + n3[0]), dy(widths + n3[0], widths + n3[0] + n3[1]),
            dz(widths + n3[0] + n3[1], widths + n3[0] + n3[1] + n3[2]);
        std::vector<double> gx(coef, coef + n3[0] + 1), gy(coef + n3[0] + 1, coef + n3[0] + n3[1] + 2),
            gz(coef + n3[0] + n3[1] + 2, coef + n3[0] + n3[1] + n3[2] + 3);
        M.host = mg_build_hierarchy(n3, periodic, dx, dy, dz, gx, gy, gz, dt, 0);
        const size_t nl = M.host.size();
        M.dev.resize(nl);
        M.axes.resize(nl);
        M.bufs.resize(nl);
        for (size_t l = 0; l < nl; ++l)
        {
            const MgHostLevel &H = M.host[l];
            MgLevel &D = M.dev[l];
            D.nx = H.n[0]; D.ny = H.n[1]; D.nz = H.n[2];
            D.perx = H.per[0]; D.pery = H.per[1]; D.perz = H.per[2];
            D.px = D.nx; D.plane = (long long)D.nx * D.ny; D.base = 0;
            D.mx = D.my = D.mz = D.sx = D.sy = D.sz = nullptr;
            if (l + 1 < nl)
            {
                D.mx = H.cmap[0].data(); D.my = H.cmap[1].data(); D.mz = H.cmap[2].data();
                D.sx = H.cstart[0].data(); D.sy = H.cstart[1].data(); D.sz = H.cstart[2].data();
            }
            size_t off[6];
            const std::vector<double> *parts[6] = {&H.d[0], &H.d[1], &H.d[2], &H.g[0], &H.g[1], &H.g[2]};
            for (int q = 0; q < 6; ++q)
            {
                off[q] = M.axes[l].size();
                M.axes[l].insert(M.axes[l].end(), parts[q]->begin(), parts[q]->end());
            }
            const double *base = M.axes[l].data();
            D.dx = base + off[0]; D.dy = base + off[1]; D.dz = base + off[2];
            D.gx = base + off[3]; D.gy = base + off[4]; D.gz = base + off[5];
            M.bufs[l].assign(5, std::vector<double>(l == 0 ? (size_t)n : (size_t)H.cells(), 0.0));
        }
        M.prm.smooth_its = smooth_its;
        M.prm.coarse_its = coarse_its;
    }
    SolveConsts kc{};
    kc.rtol = rtol; kc.atol = atol; kc.divtol = 1e4; kc.nglobal = (double)n;
    kc.max_it = max_it; kc.norm_type = 1; kc.has_const = has_const; kc.hist_cap = hist_cap;
    DevState st{};
    M.st = &st;
    emu::launch(dim3(1), dim3(32), 0, [&] { k_state_reset(&st); });
    std::vector<double> r(b, b + n), p0((size_t)n, 0.0), p1((size_t)n, 0.0), w((size_t)n, 0.0), x((size_t)n, 0.0);
    const int blocks = 3;
    const int nm = nullvec ? 2 : (has_const ? 1 : 0);
    const long long nsep = A.nsep;
    double *z = nullptr;
    auto precondition = [&] {
        z = M.cycle(r.data());
        if (n > nsep) emu::launch(dim3(blocks), dim3(256), 0, [&] { k_sep_tail_pc(nsep, n, r.data(), dinv, z, &st); });
    };
    auto sums = [&](int kind) {
        if (nm == 2) emu::launch(dim3(blocks), dim3(256), 0, [&] { k_sep_zsums<2>(n, z, r.data(), nullvec, kind, W.ws, &st, kc, hist); });
        else if (nm == 1) emu::launch(dim3(blocks), dim3(256), 0, [&] { k_sep_zsums<1>(n, z, r.data(), nullvec, kind, W.ws, &st, kc, hist); });
        else emu::launch(dim3(blocks), dim3(256), 0, [&] { k_sep_zsums<0>(n, z, r.data(), nullvec, kind, W.ws, &st, kc, hist); });
    };
    precondition();
    if (nm == 2) sums(FIN_CSR_INIT);
    else
    {
        if (nm == 1) sums(FIN_INIT_CENTRE);
        sums(FIN_INIT);
    }
    double *pp[2] = {p0.data(), p1.data()};
    for (int it = 0; it < max_it + 2 && !st.done; ++it)
    {
        CsrVecs v{z, pp[it & 1], pp[(it & 1) ^ 1], w.data(), x.data(), nullptr, nullvec};
        if (g_sep_xr == 2 || g_sep_xr == 4)
        {
            // the hybrid operator always carries its face-area weights (HYB = true)
            const SepTilePlan T = sep_tile_plan(A, g_sep_xr, g_sep_zchunk, g_sep_target > 0 ? g_sep_target : 24);
            const dim3 tgrid((unsigned)T.blocks());
#define EMU_HT(XR, NM) emu::launch(tgrid, dim3(256), sep_tile_smem_bytes<XR, 3, SepOpCg<false, NM>>(), [&] { k_sep_tile_cg_spmv<XR, true, 3, false, NM>(A, T, v, W.ws, &st, kc, hist); })
            if (g_sep_xr == 2) { if (nm == 2) EMU_HT(2, 2); else if (nm == 1) EMU_HT(2, 1); else EMU_HT(2, 0); }
            else { if (nm == 2) EMU_HT(4, 2); else if (nm == 1) EMU_HT(4, 1); else EMU_HT(4, 0); }
#undef EMU_HT
        }
        else if (nm == 2) emu::launch(dim3(blocks), dim3(256), 0, [&] { k_sep_cg_spmv<false, 2>(A, v, W.ws, &st, kc, hist); });
        else if (nm == 1) emu::launch(dim3(blocks), dim3(256), 0, [&] { k_sep_cg_spmv<false, 1>(A, v, W.ws, &st, kc, hist); });
        else emu::launch(dim3(blocks), dim3(256), 0, [&] { k_sep_cg_spmv<false, 0>(A, v, W.ws, &st, kc, hist); });
        emu::launch(dim3(blocks), dim3(256), 0, [&] { k_mg_rupdate(n, r.data(), w.data(), &st); });
        precondition();
        sums(nm == 2 ? (int)FIN_CSR_UPDATE : (int)FIN_UPDATE);
    }
    GridDev g1{};
    g1.nx = (int)n; g1.ny = 1; g1.nzl = 1; g1.px = (int)n; g1.plane = 0;
    emu::launch(dim3(1), dim3(256), 0, [&] { k_xtail(g1, x.data(), p0.data(), p1.data(), &st); });
    memcpy(x_out, x.data(), sizeof(double) * (size_t)n);
    *nhist = st.nhist;
    *its = st.its;
    *reason = st.reason;
    return st.done ? 0 : 1;
}

}  // extern "C"

// ---- matrix-free staggered-grid operators (ops_kernels.cuh).  mode 0: out = D u;  1: out = G p;  2: out = (BN G) p;
// 3: projection u -= (BN G) dp (io = u, in = dp), p += dp (io2 = p)
extern "C" EMU_API int emu_stag_ops(int mode, int dim, const int64_t *n, const int *per, const double *dx, const double *dy,
                                    const double *dz, double dt, const double *in, double *io, double *io2)
{
    Problem P;
    build(P, dim, n, per, dx, dy, dz, dt);
    StagGrid s{};
    s.nx = P.g.nx; s.ny = P.g.ny; s.nz = P.g.nzl; s.dim = dim;
    s.perx = P.per[0]; s.pery = P.per[1]; s.perz = dim == 3 ? P.per[2] : 0;
    s.nu = s.nx - (s.perx ? 0 : 1); s.nv = s.ny - (s.pery ? 0 : 1); s.nw = dim == 3 ? s.nz - (s.perz ? 0 : 1) : 0;
    s.offv = (long long)s.nu * s.ny * s.nz;
    s.offw = s.offv + (long long)s.nx * s.nv * s.nz;
    s.dx = P.g.dx; s.dy = P.g.dy; s.dz = P.g.dz; s.gx = P.g.gx; s.gy = P.g.gy; s.gz = P.g.gz;
    const long long np = (long long)s.nx * s.ny * s.nz;
    if (mode == 0) emu::launch(dim3(3), dim3(256), 0, [&] { k_divergence(s, in, io); });
    else if (mode == 1) emu::launch(dim3(3), dim3(256), 0, [&] { k_gradient<0>(s, in, io); });
    else if (mode == 2) emu::launch(dim3(3), dim3(256), 0, [&] { k_gradient<1>(s, in, io); });
    else if (mode == 3)
    {
        emu::launch(dim3(3), dim3(256), 0, [&] { k_gradient<2>(s, in, io); });
        emu::launch(dim3(3), dim3(256), 0, [&] { k_axpy_one(np, io2, in); });
    }
    else
    {
        // mode 4: io = N(q), in = the three ghosted arrays back to back;  mode 5: io = ghosted arrays (back to back) from packed in
        const long long zu = dim == 3 ? s.nz + 2 : 1, zw = dim == 3 ? s.nw + 2 : 0;
        const long long su = (long long)(s.nu + 2) * (s.ny + 2) * zu, sv = (long long)(s.nx + 2) * (s.nv + 2) * zu;
        (void)zw;
        if (mode == 4)
        {
            if (dim == 3) emu::launch(dim3(3), dim3(256), 0, [&] { k_convection<3>(s, in, in + su, in + su + sv, io); });
            else emu::launch(dim3(3), dim3(256), 0, [&] { k_convection<2>(s, in, in + su, in + su + sv, io); });
        }
        else emu::launch(dim3(3), dim3(256), 0, [&] { k_ghosted_from_packed(s, in, io, io + su, io + su + sv); });
    }
    return 0;
}

// ---- direct solve of a small assembled system (dense_kernels.cuh): factorise, then nrhs solves with the same factors
extern "C" EMU_API int emu_dense_solve(int64_t n, const int64_t *rowptr, const int32_t *col, const double *val, int nrhs,
                                       const double *b, double *x, int threads)
{
    const int lda = (int)((n + 3) & ~(int64_t)3);
    std::vector<double> a((size_t)n * (size_t)lda, 0.0);
    std::vector<int> piv(n, -1), perm(n, -1);
    int info = 0;
    emu::launch(dim3(2), dim3(256), 0, [&] { k_dense_fill(n, lda, rowptr, col, val, a.data()); });
    for (int kb = 0; kb < (int)n; kb += LU_NB)   // the launch sequence of dense_solver.inc
    {
        const int nb = std::min<int>(LU_NB, (int)n - kb), rest = (int)n - kb - nb, ncols = (int)n - nb;
        emu::launch(dim3(1), dim3(threads), 0, [&] { k_lu_panel((int)n, lda, a.data(), kb, piv.data(), &info); });
        if (ncols > 0) emu::launch(dim3((ncols + 127) / 128), dim3(128), 0, [&] { k_lu_swap_trsm((int)n, lda, a.data(), kb, piv.data(), &info); });
        if (rest > 0) emu::launch(dim3((rest + 63) / 64, (rest + 63) / 64), dim3(256), 0, [&] { k_lu_gemm((int)n, lda, a.data(), kb, &info); });
    }
    if (info != 0) return info;
    emu::launch(dim3(1), dim3(32), 0, [&] { k_lu_perm((int)n, piv.data(), perm.data()); });
    const size_t smem = sizeof(double) * ((size_t)lda + (size_t)LU_NB * (LU_NB + 1));
    const int nblk = (int)((n + LU_NB - 1) / LU_NB);
    std::vector<double> work((size_t)n, 0.0);
    std::vector<unsigned long long> flags((size_t)nblk + 1, 0ull);
    unsigned long long epoch = 0;
    for (int q = 0; q < nrhs; ++q)
    {
        const double *bq = b + (size_t)q * n;
        double *xq = x + (size_t)q * n;
        if (threads == 128 && nblk > 2)
        {
            // the multi-CTA wavefront substitution (CTAs run one after the other here, in ticket order: every wait is satisfied)
            unsigned int *ticket = reinterpret_cast<unsigned int *>(&flags[(size_t)nblk]);
            *ticket = 0;
            ++epoch;
            emu::launch(dim3(nblk), dim3(128), 0, [&] { k_dense_sweep<false>((int)n, lda, a.data(), perm.data(), bq, work.data(), xq, flags.data(), ticket, epoch); });
            *ticket = 0;
            ++epoch;
            emu::launch(dim3(nblk), dim3(128), 0, [&] { k_dense_sweep<true>((int)n, lda, a.data(), perm.data(), bq, work.data(), xq, flags.data(), ticket, epoch); });
        }
        else
            emu::launch(dim3(1), dim3(threads), smem, [&] { k_dense_solve((int)n, lda, a.data(), perm.data(), bq, xq); });
    }
    return 0;
}

