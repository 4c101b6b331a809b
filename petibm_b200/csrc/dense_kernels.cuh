// dense_kernels.cuh -- direct solve of a SMALL system (-<name>_ksp_type preonly -<name>_pc_type lu).
//
// PetIBM's decoupled IBPM solves its forces system E BN H (decoupledibpm.cpp:149-216,271-285; a few hundred to a few
// thousand unknowns, symmetric positive definite up to rounding) with a sparse direct solver: every shipped case has
// "-forces_ksp_type preonly -forces_pc_type lu -forces_pc_factor_mat_solver_type superlu_dist"
// (examples/decoupledibpm/*/config/forces_solver.info).  Here the matrix is expanded to dense storage on the device and
// factorised once per setMatrix by ONE thread block (right-looking LU without pivoting: the systems are SPD-like; a
// vanishing pivot is reported as KSP_DIVERGED_PC_FAILED), and every solve is a forward and a backward substitution by one
// block.  KSPSolve_PREONLY semantics: one application of the preconditioner, its = 1, KSP_CONVERGED_ITS.
// Not a hot path: correctness first, a few milliseconds per solve at n = 2000.
#pragma once
#include <stdint.h>

#include "kernels.cuh"

namespace b200 {

// dense row-major copy of a CSR matrix (a must be zeroed before)
__global__ void __launch_bounds__(256) k_dense_fill(int64_t n, const int64_t *rowptr, const int32_t *col, const double *val, double *a)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        for (int64_t q = rowptr[i]; q < rowptr[i + 1]; ++q) a[i * n + col[q]] = val[q];
}

// in-place LU (unit lower L below the diagonal, U on and above), one thread block of a multiple of 32 threads;
// *info = 0 or 1 + index of a zero pivot
__global__ void __launch_bounds__(1024) k_dense_lu(int n, double *a, int *info)
{
    __shared__ double s_pivot;
    const int tid = threadIdx.x, nt = blockDim.x;
    if (tid == 0) *info = 0;
    for (int k = 0; k < n; ++k)
    {
        if (tid == 0) s_pivot = a[(size_t)k * n + k];
        __syncthreads();
        const double pivot = s_pivot;
        if (pivot == 0.0 || pivot != pivot)
        {
            if (tid == 0) *info = k + 1;
            return;
        }
        for (int i = k + 1 + tid; i < n; i += nt) a[(size_t)i * n + k] = a[(size_t)i * n + k] / pivot;
        __syncthreads();
        // trailing update: a warp walks along a row (coalesced in j), the warps share out the rows
        const int tx = tid & 31, ty = tid >> 5, nty = nt >> 5;
        for (int i = k + 1 + ty; i < n; i += nty)
        {
            const double lik = a[(size_t)i * n + k];
            for (int j = k + 1 + tx; j < n; j += 32) a[(size_t)i * n + j] = a[(size_t)i * n + j] - lik * a[(size_t)k * n + j];
        }
        __syncthreads();
    }
}

// x = U^-1 L^-1 b with the factors of k_dense_lu, one thread block; x may alias b
__global__ void __launch_bounds__(1024) k_dense_solve(int n, const double *a, const double *b, double *x)
{
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int i = tid; i < n; i += nt) x[i] = b[i];
    __syncthreads();
    for (int k = 0; k < n; ++k)  // forward: L y = b (unit diagonal), column oriented
    {
        const double yk = x[k];
        for (int i = k + 1 + tid; i < n; i += nt) x[i] = x[i] - a[(size_t)i * n + k] * yk;
        __syncthreads();
    }
    for (int k = n - 1; k >= 0; --k)  // backward: U x = y
    {
        if (tid == 0) x[k] = x[k] / a[(size_t)k * n + k];
        __syncthreads();
        const double xk = x[k];
        for (int i = tid; i < k; i += nt) x[i] = x[i] - a[(size_t)i * n + k] * xk;
        __syncthreads();
    }
}

}  // namespace b200
