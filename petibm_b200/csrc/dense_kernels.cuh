// dense_kernels.cuh -- direct solve of a SMALL system (-<name>_ksp_type preonly -<name>_pc_type lu).
//
// PetIBM's decoupled IBPM solves its forces system E BN H (decoupledibpm.cpp:149-216,271-285; a few hundred to a few
// thousand unknowns, symmetric positive definite up to rounding) with a sparse direct solver: every shipped case has
// "-forces_ksp_type preonly -forces_pc_type lu -forces_pc_factor_mat_solver_type superlu_dist"
// (examples/decoupledibpm/*/config/forces_solver.info), and moving bodies re-factorise every time step
// (rigidkinematics.cpp:119-140 calls setMatrix again).  Here the matrix is expanded to dense row-major storage on the
// device (leading dimension lda = n rounded up to 4 doubles) and factorised once per setMatrix by a BLOCKED
// right-looking LU WITH PARTIAL (ROW) PIVOTING, like the sparse direct solvers it stands in for:
//   per panel of NB = 32 columns   k_lu_panel      one CTA: pivot search, row swap, scaling, rank-1 updates inside the panel
//                                  k_lu_swap_trsm  whole grid: the panel's row swaps on the other columns, U12 = L11^-1 A12
//                                  k_lu_gemm       whole grid: A22 -= L21 U12 (64 x 64 tiles, 4 x 4 per thread, FP64 FMA)
// A vanishing or NaN pivot column is reported (info = 1 + column): KSP_DIVERGED_PC_FAILED.  Every solve is a blocked
// forward and backward substitution by one CTA with the vector in shared memory (k_dense_solve).
// KSPSolve_PREONLY semantics: one application of the preconditioner, its = 1, KSP_CONVERGED_ITS.
// SURVEY.md section 8, row f4.  No bit-for-bit claim here (pivoted LU against LAPACK: 1e-11), so FMA is allowed.
#pragma once
#include <stdint.h>

#include "kernels.cuh"

namespace b200 {

constexpr int LU_NB = 32;

// dense row-major copy of a CSR matrix (a must be zeroed before)
__global__ void __launch_bounds__(256) k_dense_fill(int64_t n, int64_t lda, const int64_t *rowptr, const int32_t *col, const double *val, double *a)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        for (int64_t q = rowptr[i]; q < rowptr[i + 1]; ++q) a[i * lda + col[q]] = val[q];
}

// Panel factorisation: columns kb .. kb+nb-1, rows kb .. n-1, one CTA of a multiple of 32 threads.
// piv[kb+c] = row exchanged with row kb+c; *info = 1 + column of a zero / NaN pivot column (0 = fine, set by the host).
__global__ void __launch_bounds__(1024) k_lu_panel(int n, int lda, double *a, int kb, int *piv, int *info)
{
    __shared__ double s_val[32];
    __shared__ int s_idx[32];
    __shared__ double s_urow[LU_NB];
    __shared__ int s_stop;
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarp = nt >> 5;
    const int nb = min(LU_NB, n - kb);
    if (tid == 0) s_stop = (*info != 0);
    __syncthreads();
    if (s_stop) return;  // an earlier panel failed
    for (int c = 0; c < nb; ++c)
    {
        const int k = kb + c;
        // ---- pivot: largest |a[i][k]|, i >= k; ties to the smallest row index (deterministic)
        double best = -1.0;
        int bi = n;
        bool bad = false;
        for (int i = k + tid; i < n; i += nt)
        {
            const double v = fabs(a[(size_t)i * lda + k]);
            if (v != v) bad = true;
            if (v > best) { best = v; bi = i; }
        }
        if (bad) { best = INFINITY; bi = -1; }  // a NaN anywhere in the column wins the search and is reported
        for (int o = 16; o > 0; o >>= 1)
        {
            const double ov = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
        }
        if (lane == 0) { s_val[warp] = best; s_idx[warp] = bi; }
        __syncthreads();
        if (warp == 0)
        {
            best = lane < nwarp ? s_val[lane] : -1.0;
            bi = lane < nwarp ? s_idx[lane] : n;
            for (int o = 16; o > 0; o >>= 1)
            {
                const double ov = __shfl_xor_sync(0xffffffffu, best, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
            }
            if (lane == 0) { s_val[0] = best; s_idx[0] = bi; }
        }
        __syncthreads();
        const int p = s_idx[0];
        if (p < 0 || !(s_val[0] > 0.0))
        {
            if (tid == 0) *info = k + 1;
            return;
        }
        // ---- swap rows k and p inside the panel, publish the pivot row
        if (tid < nb)
        {
            const double vk = a[(size_t)k * lda + kb + tid], vp = a[(size_t)p * lda + kb + tid];
            if (p != k)
            {
                a[(size_t)k * lda + kb + tid] = vp;
                a[(size_t)p * lda + kb + tid] = vk;
            }
            s_urow[tid] = vp;
        }
        if (tid == 0) piv[k] = p;
        __syncthreads();
        // ---- multipliers and rank-1 update of the columns right of k inside the panel: one warp per row
        const double pivot = s_urow[c];
        const double ukj = lane < nb ? s_urow[lane] : 0.0;
        for (int i = k + 1 + warp; i < n; i += nwarp)
        {
            double *row = a + (size_t)i * lda + kb;
            const double mine = lane < nb ? row[lane] : 0.0;
            const double lik = __shfl_sync(0xffffffffu, mine, c) / pivot;
            if (lane == c) row[lane] = lik;
            else if (lane > c && lane < nb) row[lane] = fma(-lik, ukj, mine);
        }
        __syncthreads();
    }
}

// The panel's row exchanges on the columns outside the panel, then U12 = L11^-1 A12 (unit lower L11): one thread per column.
__global__ void __launch_bounds__(128) k_lu_swap_trsm(int n, int lda, double *a, int kb, const int *piv, const int *info)
{
    __shared__ double s_l[LU_NB][LU_NB + 1];
    __shared__ int s_piv[LU_NB];
    if (*info != 0) return;
    const int nb = min(LU_NB, n - kb);
    for (int e = threadIdx.x; e < nb * nb; e += blockDim.x) s_l[e / nb][e % nb] = a[(size_t)(kb + e / nb) * lda + kb + e % nb];
    if ((int)threadIdx.x < nb) s_piv[threadIdx.x] = piv[kb + threadIdx.x];
    __syncthreads();
    const int ncols = n - nb;  // all columns except the panel's own
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < ncols; q += gridDim.x * blockDim.x)
    {
        const int j = q < kb ? q : q + nb;
        for (int c = 0; c < nb; ++c)
        {
            const int p = s_piv[c];
            if (p != kb + c)
            {
                const double t = a[(size_t)(kb + c) * lda + j];
                a[(size_t)(kb + c) * lda + j] = a[(size_t)p * lda + j];
                a[(size_t)p * lda + j] = t;
            }
        }
        if (j < kb) continue;
        double u[LU_NB];
#pragma unroll
        for (int r = 0; r < LU_NB; ++r) u[r] = r < nb ? a[(size_t)(kb + r) * lda + j] : 0.0;
#pragma unroll
        for (int c = 0; c < LU_NB; ++c)
#pragma unroll
            for (int r = c + 1; r < LU_NB; ++r)
                if (r < nb) u[r] = fma(-s_l[r][c], u[c], u[r]);
#pragma unroll
        for (int r = 1; r < LU_NB; ++r)
            if (r < nb) a[(size_t)(kb + r) * lda + j] = u[r];
    }
}

// Trailing update A22 -= L21 U12 behind a full panel (nb = LU_NB): 64 x 64 tile per CTA of 256 threads, 4 x 4 per thread.
__global__ void __launch_bounds__(256) k_lu_gemm(int n, int lda, double *a, int kb, const int *info)
{
    __shared__ double s_lt[LU_NB][64 + 1];  // L21 tile, transposed: [c][row]
    __shared__ double s_u[LU_NB][64];       // U12 tile: [c][col]
    if (*info != 0) return;
    const int r0 = kb + LU_NB + blockIdx.y * 64, c0 = kb + LU_NB + blockIdx.x * 64;
    const int tid = threadIdx.x;
    for (int e = tid; e < 64 * LU_NB; e += 256)
    {
        const int r = e / LU_NB, c = e % LU_NB;  // a warp reads one 256-byte row segment of L21
        s_lt[c][r] = (r0 + r < n) ? a[(size_t)(r0 + r) * lda + kb + c] : 0.0;
        const int uc = e % 64, ur = e / 64;       // and half a 512-byte row segment of U12
        s_u[ur][uc] = (c0 + uc < n) ? a[(size_t)(kb + ur) * lda + c0 + uc] : 0.0;
    }
    __syncthreads();
    const int tx = tid & 15, ty = tid >> 4;  // columns tx + 16 q, rows ty + 16 r
    double acc[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[r][q] = 0.0;
#pragma unroll 8
    for (int c = 0; c < LU_NB; ++c)
    {
        double lv[4], uv[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) lv[r] = s_lt[c][ty + 16 * r];
#pragma unroll
        for (int q = 0; q < 4; ++q) uv[q] = s_u[c][tx + 16 * q];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[r][q] = fma(lv[r], uv[q], acc[r][q]);
    }
#pragma unroll
    for (int r = 0; r < 4; ++r)
    {
        const int i = r0 + ty + 16 * r;
        if (i >= n) continue;
#pragma unroll
        for (int q = 0; q < 4; ++q)
        {
            const int j = c0 + tx + 16 * q;
            if (j < n) a[(size_t)i * lda + j] -= acc[r][q];
        }
    }
}

// perm[i] = row of b that ends up in position i after the recorded exchanges (one thread; once per factorisation)
__global__ void k_lu_perm(int n, const int *piv, int *perm)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    for (int i = 0; i < n; ++i) perm[i] = i;
    for (int k = 0; k < n; ++k)
    {
        const int p = piv[k];
        if (p != k)
        {
            const int t = perm[k];
            perm[k] = perm[p];
            perm[p] = t;
        }
    }
}

// x = U^-1 L^-1 P b with the factors of the kernels above: one CTA, the vector lives in dynamic shared memory
// (8 n bytes, + one staged 32 x 33 diagonal block); x may alias b.
__global__ void __launch_bounds__(1024) k_dense_solve(int n, int lda, const double *a, const int *perm, const double *b, double *x)
{
    B200_DYNAMIC_SMEM(smem_raw);
    double *xs = reinterpret_cast<double *>(smem_raw);
    double(*dg)[LU_NB + 1] = reinterpret_cast<double(*)[LU_NB + 1]>(xs + ((n + 3) & ~3));
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < n; i += nt) xs[i] = b[perm[i]];
    // ---- forward: L y = P b (unit diagonal)
    for (int kb = 0; kb < n; kb += LU_NB)
    {
        const int nb = min(LU_NB, n - kb);
        for (int e = tid; e < nb * nb; e += nt) dg[e / nb][e % nb] = a[(size_t)(kb + e / nb) * lda + kb + e % nb];
        __syncthreads();
        if (warp == 0)
        {
            double xr = lane < nb ? xs[kb + lane] : 0.0;
            for (int c = 0; c < nb; ++c)
            {
                const double xc = __shfl_sync(0xffffffffu, xr, c);
                if (lane > c && lane < nb) xr = fma(-dg[lane][c], xc, xr);
            }
            if (lane < nb) xs[kb + lane] = xr;
        }
        __syncthreads();
        for (int i = kb + nb + tid; i < n; i += nt)
        {
            const double *row = a + (size_t)i * lda + kb;
            double acc = 0.0;
#pragma unroll 8
            for (int c = 0; c < LU_NB; ++c) acc = fma(row[c], xs[kb + c], acc);  // behind a partial last panel there are no rows
            xs[i] -= acc;
        }
        __syncthreads();
    }
    // ---- backward: U x = y
    for (int kb = ((n - 1) / LU_NB) * LU_NB; kb >= 0; kb -= LU_NB)
    {
        const int nb = min(LU_NB, n - kb);
        for (int e = tid; e < nb * nb; e += nt) dg[e / nb][e % nb] = a[(size_t)(kb + e / nb) * lda + kb + e % nb];
        __syncthreads();
        if (warp == 0)
        {
            double xr = lane < nb ? xs[kb + lane] : 0.0;
            for (int c = nb - 1; c >= 0; --c)
            {
                if (lane == c) xr = xr / dg[c][c];
                const double xc = __shfl_sync(0xffffffffu, xr, c);
                if (lane < c) xr = fma(-dg[lane][c], xc, xr);
            }
            if (lane < nb) xs[kb + lane] = xr;
        }
        __syncthreads();
        for (int i = tid; i < kb; i += nt)
        {
            const double *row = a + (size_t)i * lda + kb;
            double acc = 0.0;
            for (int c = 0; c < nb; ++c) acc = fma(row[c], xs[kb + c], acc);
            xs[i] -= acc;
        }
        __syncthreads();
    }
    for (int i = tid; i < n; i += nt) x[i] = xs[i];
}

// Multi-CTA substitution: one CTA of four warps per block of 32 unknowns, wavefront over the blocks.  A CTA first streams
// the off-diagonal blocks of its block row against the parts of the vector that are already final (each part is
// published with a release flag by the CTA that solved it), and solves its 32 x 32 triangle as soon as the last one has
// arrived: the dependent chain is one flag hop + one triangle per block (about 3 us) instead of one CTA streaming all
// factors.  UPPER = false: y = L^-1 P b (unit diagonal); UPPER = true: x = U^-1 y, in place in `work`, copied to `out`.
// The logical block index comes from a ticket, so a CTA only ever waits for CTAs that started before it.
template <bool UPPER>
__global__ void __launch_bounds__(128) k_dense_sweep(int n, int lda, const double *a, const int *perm, const double *b, double *work,
                                                     double *out, unsigned long long *flags, unsigned int *ticket, unsigned long long epoch)
{
    __shared__ double s_part[4][LU_NB];
    __shared__ double s_diag[LU_NB][LU_NB + 1];
    __shared__ unsigned int s_blk;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nblk = (n + LU_NB - 1) / LU_NB;
    if (tid == 0) s_blk = atomicAdd(ticket, 1u);
    __syncthreads();
    const int I = UPPER ? nblk - 1 - (int)s_blk : (int)s_blk;
    const int r0 = I * LU_NB, nr = min(LU_NB, n - r0);
    for (int e = tid; e < nr * nr; e += 128) s_diag[e / nr][e % nr] = a[(size_t)(r0 + e / nr) * lda + r0 + e % nr];
    const bool row_ok = lane < nr;
    const double *arow = a + (size_t)(r0 + (row_ok ? lane : 0)) * lda;
    double acc = 0.0;
    const int count = UPPER ? nblk - 1 - I : I;  // blocks this row depends on
    for (int q = warp; q < count; q += 4)
    {
        const int J = UPPER ? nblk - 1 - q : q;  // in the order in which they become final
        const int c0 = J * LU_NB, nc = min(LU_NB, n - c0);
        double lrow[LU_NB];
#pragma unroll
        for (int c = 0; c < LU_NB; ++c) lrow[c] = (row_ok && c < nc) ? arow[c0 + c] : 0.0;  // static data: loaded before the wait
        if (lane == 0)
        {
            unsigned int spins = 0;
            const unsigned long long t0 = global_timer_ns();
            while (ld_acquire_sys(flags + J) < epoch)
                if ((++spins & 1023u) == 0 && global_timer_ns() - t0 > 4000000000ull) break;  // never hang the GPU
        }
        __syncwarp();
        const double yv = lane < nc ? __ldcg(work + c0 + lane) : 0.0;
#pragma unroll
        for (int c = 0; c < LU_NB; ++c) acc = fma(-lrow[c], __shfl_sync(0xffffffffu, yv, c), acc);
    }
    s_part[warp][lane] = acc;
    __syncthreads();
    if (warp == 0)
    {
        double xr = 0.0;
        if (row_ok) xr = UPPER ? __ldcg(work + r0 + lane) : b[perm[r0 + lane]];
        xr += (s_part[0][lane] + s_part[1][lane]) + (s_part[2][lane] + s_part[3][lane]);
        if (!UPPER)
        {
            for (int c = 0; c < nr; ++c)
            {
                const double xc = __shfl_sync(0xffffffffu, xr, c);
                if (lane > c && lane < nr) xr = fma(-s_diag[lane][c], xc, xr);
            }
        }
        else
        {
            for (int c = nr - 1; c >= 0; --c)
            {
                if (lane == c) xr = xr / s_diag[c][c];
                const double xc = __shfl_sync(0xffffffffu, xr, c);
                if (lane < c) xr = fma(-s_diag[lane][c], xc, xr);
            }
        }
        if (row_ok)
        {
            work[r0 + lane] = xr;
            if (UPPER) out[r0 + lane] = xr;
        }
        __threadfence();
        __syncwarp();
        if (lane == 0) st_release_sys(flags + I, epoch);
    }
}

}  // namespace b200
