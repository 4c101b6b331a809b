// ops_kernels.cuh -- the operators on either side of the pressure solve, matrix-free.
//
// Around pSolver->solve(dP, rhs2) PetIBM runs three assembled-matrix products per time step, all on the host:
//   rhs2 = D u*                         MatMult(D, UGlobal, rhs2)            navierstokes.cpp:540-551
//   u    = u - (BN G) dP                MatMult(BNG, dP, rhs1); VecAXPY(-1)  navierstokes.cpp:583-599
//   p    = p + dP                       VecAXPY(pGlobal, 1.0, dP)            navierstokes.cpp:601-615
// and G p in the velocity right-hand side (navierstokes.cpp:442).  D (createdivergence.cpp:140-223: +-face areas) and G
// (creategradient.cpp:70-128: +-1/h) of a stretched Cartesian staggered grid are fully described by the same 1-D arrays
// the Poisson kernels use, so with these kernels b and x of the solve never have to leave the device
// (b200ls_solve_device): divergence -> solve -> projection is a chain of launches on the solver's stream.
//
// Every row sum is formed in ascending column order of the packed vector [u | v | w] (what MatMult_SeqAIJ does), products
// and sums without FMA contraction, so the results are bit-identical to the assembled products (oracle:
// orc_assemble_divergence / orc_assemble_gradient / orc_bnhead_order1 + orc_matmatmult).  Boundary contributions
// (DCorrection, bc1) are the application's MatShell terms and stay with the caller.
#pragma once
#include <stdint.h>

#include "kernels.cuh"

namespace b200 {

struct StagGrid
{
    int nx, ny, nz;            // pressure cells (nz = 1 in 2-D)
    int dim;
    int perx, pery, perz;
    int nu, nv, nw;            // points of u along x, of v along y, of w along z (n - 1, or n on a periodic axis)
    long long offv, offw;      // first index of v / w in the packed velocity vector
    const double *dx, *dy, *dz;  // cell widths
    const double *gx, *gy, *gz;  // dt * (1/h) on the minus face of cell s (b200ls_set_poisson_stencil)
};

// (D u)(i,j,k): columns u(i-1), u(i), v(j-1), v(j), w(k-1), w(k) in ascending packed index; a periodic wrap puts the
// minus face behind the plus face
__global__ void __launch_bounds__(256) k_divergence(StagGrid g, const double *u, double *out)
{
    const long long n = (long long)g.nx * g.ny * g.nz;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x)
    {
        const int i = (int)(t % g.nx);
        const long long row = t / g.nx;
        const int j = (int)(row % g.ny), k = (int)(row / g.ny);
        const double ayz = __dmul_rn(g.dy[j], g.dz[k]);
        double s = 0.0;
        // one direction: minus face fm (or -1), plus face fp (or -1), base + stride * face = packed index
        auto dir = [&](double area, int c, int nc, int nf, int per, long long base, long long stride) {
            const int fm = c > 0 ? c - 1 : (per ? nc - 1 : -1);
            const int fp = c < nf ? c : (per ? c : -1);
            const bool wrap = per && c == 0;  // the minus face is the last one of the axis: larger packed index
            if (wrap && fp >= 0) s = __dadd_rn(s, __dmul_rn(area, u[base + stride * fp]));
            if (fm >= 0) s = __dadd_rn(s, __dmul_rn(-area, u[base + stride * fm]));
            if (!wrap && fp >= 0) s = __dadd_rn(s, __dmul_rn(area, u[base + stride * fp]));
        };
        dir(ayz, i, g.nx, g.nu, g.perx, (long long)g.nu * (j + (long long)g.ny * k), 1);
        const double axz = __dmul_rn(g.dx[i], g.dz[k]);
        dir(axz, j, g.ny, g.nv, g.pery, g.offv + i + (long long)g.nx * g.nv * k, g.nx);
        if (g.dim == 3)
        {
            const double axy = __dmul_rn(g.dx[i], g.dy[j]);
            dir(axy, k, g.nz, g.nw, g.perz, g.offw + i + (long long)g.nx * j, (long long)g.nx * g.ny);
        }
        out[t] = s;
    }
}

// coefficient of one face: with_bn ? dt * (1/h) (the entry of BN G, BN = dt I: createbn.cpp:49-53 + MatMatMult)
//                                  : 1/h          (the entry of G), h = 0.5 * (d[s] + d[s+1]) (wrap: d[n-1], d[0])
__device__ __forceinline__ double face_coef(const double *d, const double *gface, int s, int n, bool with_bn)
{
    if (with_bn) return gface[s + 1];
    const double dn = d[s + 1 < n ? s + 1 : 0];
    return __ddiv_rn(1.0, __dmul_rn(0.5, __dadd_rn(d[s], dn)));
}

// MODE 0: out = G p;  MODE 1: out = (BN G) p;  MODE 2: u = u + (-1.0) * ((BN G) dp)  (projection, in place)
// One thread per velocity point; row of G: columns p(s), p(s+1) -- on a periodic axis the last face sees p(n-1) and p(0),
// and p(0) is the smaller column.
template <int MODE>
__global__ void __launch_bounds__(256) k_gradient(StagGrid g, const double *p, double *out)
{
    const long long nu = (long long)g.nu * g.ny * g.nz, nv = (long long)g.nx * g.nv * g.nz;
    const long long nw = g.dim == 3 ? (long long)g.nx * g.ny * g.nw : 0;
    const long long n = nu + nv + nw;
    const bool bn = MODE != 0;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x)
    {
        long long lo, hi;  // pressure indices of the two cells of this face
        double c;
        bool wrap = false;
        if (t < nu)
        {
            const int i = (int)(t % g.nu);
            const long long row = t / g.nu;  // j + ny*k
            c = face_coef(g.dx, g.gx, i, g.nx, bn);
            wrap = (i == g.nx - 1);
            lo = i + (long long)g.nx * row;
            hi = (wrap ? 0 : i + 1) + (long long)g.nx * row;
        }
        else if (t < nu + nv)
        {
            const long long l = t - nu;
            const int i = (int)(l % g.nx);
            const long long r2 = l / g.nx;
            const int j = (int)(r2 % g.nv), k = (int)(r2 / g.nv);
            c = face_coef(g.dy, g.gy, j, g.ny, bn);
            wrap = (j == g.ny - 1);
            lo = i + (long long)g.nx * (j + (long long)g.ny * k);
            hi = i + (long long)g.nx * ((wrap ? 0 : j + 1) + (long long)g.ny * k);
        }
        else
        {
            const long long l = t - nu - nv;
            const int i = (int)(l % g.nx);
            const long long r2 = l / g.nx;
            const int j = (int)(r2 % g.ny), k = (int)(r2 / g.ny);
            c = face_coef(g.dz, g.gz, k, g.nz, bn);
            wrap = (k == g.nz - 1);
            lo = i + (long long)g.nx * (j + (long long)g.ny * k);
            hi = i + (long long)g.nx * (j + (long long)g.ny * (wrap ? 0 : k + 1));
        }
        double s;
        if (!wrap) s = __dadd_rn(__dmul_rn(-c, p[lo]), __dmul_rn(c, p[hi]));
        else s = __dadd_rn(__dmul_rn(c, p[hi]), __dmul_rn(-c, p[lo]));
        if (MODE == 2) out[t] = __dadd_rn(out[t], __dmul_rn(-1.0, s));  // VecAXPY(UGlobal, -1.0, BNG dP)
        else out[t] = s;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Convection term N(q) of the staggered grid (createconvection.cpp: ConvectionMult2D :205-262, ConvectionMult3D :266-332,
// point kernels :39-200) -- the explicit part of rhs1 every time step (navierstokes.cpp:446-470).
// Input: the GHOSTED local arrays of u, v, w (one ghost layer on every side, i fastest), i.e. what PetIBM's
// DMCompositeScatterArray + Boundary::copyValues2LocalVecs hand to the point kernels; output: the packed vector [u|v|w].
// For field f and direction d, with e_d the unit index offset:
//     d == f : ((a+ a+) - (a- a-)) / dL        a+- = (q_f(c) + q_f(c +- e_d)) / 2
//     d != f : ((t+ a+) - (t- a-)) / dL        t+ = (q_d(c) + q_d(c + e_f)) / 2,  t- = (q_d(c - e_d) + q_d(c - e_d + e_f)) / 2
// summed over d = x, y, z in that order.  dL = mesh->dL[f][d]: the cell width for d != f, the mean of the two adjacent
// widths for d == f (cartesianmesh.cpp:237-247; wrap pair on a periodic axis :251-273).  Products, differences, halvings
// and divisions are separate IEEE operations in the reference's order: bit-identical to the oracle's restatement.
// FP64 divisions make this kernel compute-heavier than its 16 B/point of traffic; it runs once per time step.
struct GhostedDims
{
    int nx[3], ny[3], nz[3];   // interior points of field f per direction
    long long off[3];          // first index of field f in the packed vector
};
__device__ __forceinline__ GhostedDims ghosted_dims(const StagGrid &g)
{
    GhostedDims d;
    d.nx[0] = g.nu; d.ny[0] = g.ny; d.nz[0] = g.nz;
    d.nx[1] = g.nx; d.ny[1] = g.nv; d.nz[1] = g.nz;
    d.nx[2] = g.nx; d.ny[2] = g.ny; d.nz[2] = g.nw;
    d.off[0] = 0; d.off[1] = g.offv; d.off[2] = g.offw;
    return d;
}

// one velocity point of field F: every index is a compile-time-unrolled register (the run-time field / direction loops of
// a first version went through local memory: 2.1 ms for 50 M points, profiles/r02_step_bench.log)
template <int DIM, int F>
__device__ __forceinline__ double convection_point(const StagGrid &g, const GhostedDims &D, const double *const (&q)[3], long long l)
{
    const int c0 = (int)(l % D.nx[F]);
    const long long row = l / D.nx[F];
    const int c1 = (int)(row % D.ny[F]), c2 = (int)(row / D.ny[F]);
    const int c[3] = {c0, c1, c2};
    auto at = [&](int fl, int a0, int a1, int a2) -> double {
        const long long px = D.nx[fl] + 2, py = D.ny[fl] + 2;
        return q[fl][(a0 + 1) + px * ((a1 + 1) + py * (DIM == 3 ? a2 + 1 : 0))];
    };
    auto half = [](double a, double b) { return __dmul_rn(__dadd_rn(a, b), 0.5); };
    const double self = at(F, c[0], c[1], c[2]);
    double sum = 0.0;
#pragma unroll
    for (int d = 0; d < DIM; ++d)
    {
        const int e0 = d == 0, e1 = d == 1, e2 = d == 2;     // unit offset along d
        const int f0 = F == 0, f1 = F == 1, f2 = F == 2;     // unit offset along the field's own direction
        const double ap = half(self, at(F, c[0] + e0, c[1] + e1, c[2] + e2));
        const double am = half(self, at(F, c[0] - e0, c[1] - e1, c[2] - e2));
        const double *w = d == 0 ? g.dx : (d == 1 ? g.dy : g.dz);
        const int nd = d == 0 ? g.nx : (d == 1 ? g.ny : g.nz);
        double dl, tp, tm;
        if (d == F)
        {
            dl = __dmul_rn(0.5, __dadd_rn(w[c[d] + 1 < nd ? c[d] + 1 : 0], w[c[d]]));
            tp = ap;
            tm = am;
        }
        else
        {
            dl = w[c[d]];
            tp = half(at(d, c[0], c[1], c[2]), at(d, c[0] + f0, c[1] + f1, c[2] + f2));
            tm = half(at(d, c[0] - e0, c[1] - e1, c[2] - e2), at(d, c[0] - e0 + f0, c[1] - e1 + f1, c[2] - e2 + f2));
        }
        const double term = __ddiv_rn(__dadd_rn(__dmul_rn(tp, ap), -__dmul_rn(tm, am)), dl);
        sum = d == 0 ? term : __dadd_rn(sum, term);
    }
    return sum;
}

template <int DIM>
__global__ void __launch_bounds__(256) k_convection(StagGrid g, const double *qu, const double *qv, const double *qw, double *out)
{
    const GhostedDims D = ghosted_dims(g);
    const double *const q[3] = {qu, qv, qw};
    const long long ntot = D.off[DIM - 1] + (long long)D.nx[DIM - 1] * D.ny[DIM - 1] * D.nz[DIM - 1];
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < ntot; t += (long long)gridDim.x * blockDim.x)
    {
        double r;
        if (DIM == 3 && t >= D.off[2]) r = convection_point<DIM, (DIM == 3 ? 2 : 0)>(g, D, q, t - D.off[2]);
        else if (t >= D.off[1]) r = convection_point<DIM, 1>(g, D, q, t - D.off[1]);
        else r = convection_point<DIM, 0>(g, D, q, t);
        out[t] = r;
    }
}

// Interior of the ghosted arrays from the packed vector, plus the wrap layers of periodic axes (what DMGlobalToLocal
// leaves in a DMDA local vector); the ghost layers of non-periodic axes belong to the boundary conditions
// (Boundary::copyValues2LocalVecs) and are not touched.
__global__ void __launch_bounds__(256) k_ghosted_from_packed(StagGrid g, const double *packed, double *qu, double *qv, double *qw)
{
    const GhostedDims D = ghosted_dims(g);
    double *const q[3] = {qu, qv, qw};
    const int per[3] = {g.perx, g.pery, g.perz};
    const bool three = g.dim == 3;
    for (int f = 0; f < g.dim; ++f)
    {
        const long long px = D.nx[f] + 2, py = D.ny[f] + 2, pz = three ? D.nz[f] + 2 : 1;
        const int n[3] = {D.nx[f], D.ny[f], D.nz[f]};
        for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < px * py * pz; t += (long long)gridDim.x * blockDim.x)
        {
            int a[3] = {(int)(t % px) - 1, (int)((t / px) % py) - 1, three ? (int)(t / (px * py)) - 1 : 0};
            bool ok = true;
            for (int d = 0; d < g.dim; ++d)
            {
                if (a[d] < 0 || a[d] >= n[d])
                {
                    if (per[d]) a[d] = a[d] < 0 ? n[d] - 1 : 0;
                    else ok = false;
                }
            }
            if (ok) q[f][t] = packed[D.off[f] + a[0] + (long long)n[0] * (a[1] + (long long)n[1] * a[2])];
        }
    }
}

// p = p + 1.0 * dp   (VecAXPY(pGlobal, 1.0, dP))
__global__ void __launch_bounds__(256) k_axpy_one(long long n, double *p, const double *dp)
{
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x)
        p[t] = __dadd_rn(p[t], __dmul_rn(1.0, dp[t]));
}

}  // namespace b200
