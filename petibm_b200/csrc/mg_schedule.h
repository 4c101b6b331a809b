// mg_schedule.h -- host-only part of the geometric multigrid preconditioner (see mg_kernels.cuh): the level
// hierarchy of the separable operator and the ORDER in which the kernels of one V-cycle run.  No CUDA in this
// header: b200ls.cu instantiates the schedule with a launcher that enqueues the kernels on the solver's stream, the
// CPU emulation (tests/emu) with one that runs the same kernel sources under the fiber emulation -- the schedule,
// the buffer rotation and the Chebyshev coefficients are shared, not restated.
#pragma once
#include <algorithm>
#include <cstdint>
#include <vector>

namespace b200 {

struct MgHostLevel
{
    int n[3] = {1, 1, 1};
    int per[3] = {0, 0, 0};
    std::vector<double> d[3], g[3];   // widths (n) and face coefficients dt/h (n + 1; 0 at walls, wrap value if periodic)
    // transition to the next level, per axis: cmap[i] = coarse cell that holds fine cell i (one or two fine cells per
    // coarse cell), cstart[I] = first fine cell of coarse cell I (size n_coarse + 1); empty on the coarsest level
    std::vector<int> cmap[3], cstart[3];
    int64_t cells() const { return (int64_t)n[0] * n[1] * n[2]; }
};

// face coefficients of one axis: g[s] = dt * (1 / (0.5 * (d[s-1] + d[s]))) -- the grouping of creategradient.cpp:72
// and createbn.cpp:49; an axis with one cell and no wrap is inactive
inline std::vector<double> mg_faces(const std::vector<double> &d, bool periodic, double dt)
{
    const size_t m = d.size();
    std::vector<double> g(m + 1, 0.0);
    for (size_t s = 1; s < m; ++s) g[s] = dt * (1.0 / (0.5 * (d[s] + d[s - 1])));
    if (periodic && m >= 1) g[0] = g[m] = dt * (1.0 / (0.5 * (d[0] + d[m - 1])));
    return g;
}

// Width-equalising coarsening.  PetIBM's grids are stretched (cell aspect ratios of 40 and more away from the body:
// examples/ibpm/cylinder2dRe100_GPU/config.yaml), and merging every pair of cells on every axis keeps that anisotropy on
// all levels, where a point smoother cannot reduce the error along the weakly coupled direction (343 PCG iterations on
// the 450 x 450 cylinder grid).  Instead, on every level only the NARROWEST cells are merged: with smin the smallest sum
// of two neighbouring widths over all axes that still have at least four cells, a pair (i, i+1) is merged (greedy, left
// to right) iff d[i] + d[i+1] <= ratio * smin.  A uniform grid is halved on every axis as usual; on a stretched grid the
// fine band is coarsened first -- in the direction in which its cells are narrow, i.e. strongly coupled -- until the
// widths have evened out (12 iterations on the same grid at 2.5 x the fine-level work).  The hierarchy stays a tensor
// product of 1-D grids, so every level is still described by six 1-D arrays.
inline std::vector<MgHostLevel> mg_build_hierarchy(const int64_t n[3], const int per[3], const std::vector<double> &dx,
                                                   const std::vector<double> &dy, const std::vector<double> &dz,
                                                   const std::vector<double> &gx, const std::vector<double> &gy,
                                                   const std::vector<double> &gz, double dt, int max_levels,
                                                   double ratio = 2.0)
{
    std::vector<MgHostLevel> lv(1);
    const std::vector<double> *d0[3] = {&dx, &dy, &dz}, *g0[3] = {&gx, &gy, &gz};
    for (int a = 0; a < 3; ++a)
    {
        lv[0].n[a] = (int)n[a];
        lv[0].per[a] = per[a] ? 1 : 0;
        lv[0].d[a].assign(d0[a]->begin(), d0[a]->begin() + n[a]);
        lv[0].g[a].assign(g0[a]->begin(), g0[a]->begin() + n[a] + 1);
    }
    const int cap = max_levels > 0 ? max_levels : 32;
    while ((int)lv.size() < cap)
    {
        MgHostLevel f = lv.back();
        double smin = 0.0;
        bool any = false;
        for (int a = 0; a < 3; ++a)
        {
            if (f.n[a] < 4) continue;
            for (int i = 0; i + 1 < f.n[a]; ++i)
            {
                const double s = f.d[a][(size_t)i] + f.d[a][(size_t)i + 1];
                if (!any || s < smin) smin = s;
                any = true;
            }
        }
        if (!any) break;
        const double tau = ratio * smin;
        MgHostLevel c;
        for (int a = 0; a < 3; ++a)
        {
            c.per[a] = f.per[a];
            f.cmap[a].assign((size_t)f.n[a], 0);
            f.cstart[a].clear();
            int I = 0;
            for (int i = 0; i < f.n[a]; ++I)
            {
                f.cstart[a].push_back(i);
                const bool pair = f.n[a] >= 4 && i + 1 < f.n[a] && f.d[a][(size_t)i] + f.d[a][(size_t)i + 1] <= tau;
                f.cmap[a][(size_t)i] = I;
                double w = f.d[a][(size_t)i];
                if (pair)
                {
                    f.cmap[a][(size_t)i + 1] = I;
                    w = f.d[a][(size_t)i] + f.d[a][(size_t)i + 1];
                }
                c.d[a].push_back(w);
                i += pair ? 2 : 1;
            }
            f.cstart[a].push_back(f.n[a]);
            c.n[a] = I;
            // an axis that is inactive on the fine grid (2-D: one cell, wall on both sides) stays inactive
            const bool active = f.n[a] > 1 || f.per[a];
            c.g[a] = active ? mg_faces(c.d[a], c.per[a] != 0, dt) : std::vector<double>((size_t)c.n[a] + 1, 0.0);
        }
        lv.back() = f;
        lv.push_back(c);
    }
    return lv;
}

struct MgParams
{
    int smooth_its = 2;        // Chebyshev degree before and after the coarse correction
    int coarse_its = 16;       // Chebyshev degree on the coarsest level
    double lmax = 2.0;         // Gershgorin bound of D^-1 A (rows sum to zero)
    int tail_level = -1;       // >= 1: levels from here down run as ONE launch (Launcher::tail), see k_mg_tail
    double smooth_ratio = 5.0; // smoothing interval [lmax / ratio, lmax] (tuned on uniform and PetIBM-like stretched grids)
    double coarse_ratio = 40.0;
};

// Chebyshev recurrence for D^-1 A with spectrum in [lmax/ratio, lmax] (Saad, Iterative Methods, alg. 12.1):
//   d_0 = (1/theta) D^-1 r_0;   d_k = rho_k rho_{k-1} d_{k-1} + (2 rho_k / delta) D^-1 r_k,   rho_k = 1/(2 sigma - rho_{k-1})
struct MgCheb
{
    double theta, delta, sigma, rho;
    MgCheb(double lmax, double ratio)
    {
        const double lmin = lmax / ratio;
        theta = 0.5 * (lmax + lmin);
        delta = 0.5 * (lmax - lmin);
        sigma = theta / delta;
        rho = 1.0 / sigma;
    }
    double first() const { return 1.0 / theta; }
    void next(double &c1, double &c2)
    {
        const double rho1 = 1.0 / (2.0 * sigma - rho);
        c1 = rho1 * rho;
        c2 = 2.0 * rho1 / delta;
        rho = rho1;
    }
};

// One recorded step of the cycle (k_mg_tail interprets a list of these in a single CTA)
struct MgOp
{
    int kind;                  // 0 first, 1 step, 2 restrict
    int l;                     // level
    int xzero, dzero, prolong, last;
    const double *b, *xin, *din, *ec;
    double *xout, *dout;       // restrict: xout = xsum, dout = coarse right-hand side
    double c1, c2;             // first: c2 = 1/theta
};

// launcher that writes the steps down instead of running them
struct MgProgramRecorder
{
    std::vector<MgOp> ops;
    void first(int l, const double *b, double *dout, double inv_theta)
    {
        ops.push_back(MgOp{0, l, 1, 0, 0, 0, b, nullptr, nullptr, nullptr, nullptr, dout, 0.0, inv_theta});
    }
    void step(int l, bool xzero, bool dzero, bool prolong, bool last, const double *b, const double *xin, const double *din,
              const double *ec, double *xout, double *dout, double c1, double c2)
    {
        ops.push_back(MgOp{1, l, xzero, dzero, prolong, last, b, xin, din, ec, xout, dout, c1, c2});
    }
    void restrict(int l, bool xzero, const double *b, const double *xin, const double *din, double *xsum, double *bc)
    {
        ops.push_back(MgOp{2, l, xzero, 0, 0, 0, b, xin, din, nullptr, xsum, bc, 0.0, 0.0});
    }
    double *tail() { return nullptr; }
};

// the level from which the rest of the cycle is small enough for one CTA (or -1)
inline int mg_tail_level(const std::vector<MgHostLevel> &lv, int64_t max_cells)
{
    for (size_t l = 1; l < lv.size(); ++l)
        if (lv[l].cells() <= max_cells) return (int)l;
    return -1;
}

// One V-cycle on level l for the right-hand side b; returns the buffer that holds the result.
//   work[l][0..3]: four work vectors of level l;  rhs[l]: right-hand side of level l (l > 0).
// Launcher: first(l, b, dout, inv_theta); step(l, xzero, dzero, prolong, last, b, xin, din, ec, xout, dout, c1, c2);
//           restrict(l, xzero, b, xin, din, xsum, b_coarse); tail(): runs the recorded steps of the levels >= prm.tail_level
//           on rhs[tail_level] and returns that level's result buffer
template <class Launcher>
double *mg_cycle(int l, int nlevels, const double *b, double *const (*work)[4], double *const *rhs, const MgParams &prm,
                 Launcher &L)
{
    double *const *W = work[l];
    const bool coarsest = (l == nlevels - 1);
    const int deg = std::max(1, coarsest ? prm.coarse_its : prm.smooth_its);
    auto pick = [&](const double *a, const double *b2, const double *c) -> double * {
        for (int q = 0; q < 4; ++q)
            if (W[q] != a && W[q] != b2 && W[q] != c) return W[q];
        return nullptr;
    };
    // ---- from a zero guess: deg updates (pre-smoothing, or the whole coarse solve)
    MgCheb ch(prm.lmax, coarsest ? prm.coarse_ratio : prm.smooth_ratio);
    double *x = nullptr, *d = W[0];
    L.first(l, b, d, ch.first());
    for (int s = 1; s < deg; ++s)
    {
        double c1, c2;
        ch.next(c1, c2);
        const bool last = coarsest && s == deg - 1;
        double *xo = pick(x, d, nullptr), *dn = last ? nullptr : pick(x, d, xo);
        L.step(l, x == nullptr, false, false, last, b, x, d, nullptr, xo, dn, c1, c2);
        x = xo;
        d = dn;
    }
    if (coarsest) return x ? x : d;  // deg == 1: the iterate is 0 + d_0
    // ---- residual of x + d, restricted; x + d is materialised
    double *xs = pick(x, d, nullptr);
    L.restrict(l, x == nullptr, b, x, d, xs, rhs[l + 1]);
    const double *ec = (l + 1 == prm.tail_level) ? L.tail() : mg_cycle(l + 1, nlevels, rhs[l + 1], work, rhs, prm, L);
    // ---- post-smoothing: deg updates starting from xs + P e_c (prolongation folded into the first step)
    MgCheb cp(prm.lmax, prm.smooth_ratio);
    {
        const bool last = deg == 1;
        double *xo = pick(xs, nullptr, nullptr), *dn = last ? nullptr : pick(xs, xo, nullptr);
        L.step(l, false, true, true, last, b, xs, nullptr, ec, xo, dn, 0.0, cp.first());
        x = xo;
        d = dn;
    }
    for (int s = 1; s < deg; ++s)
    {
        double c1, c2;
        cp.next(c1, c2);
        const bool last = s == deg - 1;
        double *xo = pick(x, d, nullptr), *dn = last ? nullptr : pick(x, d, xo);
        L.step(l, false, false, false, last, b, x, d, nullptr, xo, dn, c1, c2);
        x = xo;
        d = dn;
    }
    return x;
}

}  // namespace b200
