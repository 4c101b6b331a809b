// stream_bench.cu -- attainable HBM bandwidth for the access mixes of the two CG kernels
// (1R+1W copy, 2R+1W update, 3R+3W fused-SpMV-like), as plain streaming kernels.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o stream_bench stream_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

template <int U>
__global__ void __launch_bounds__(256) k_copy(const double2 *__restrict__ a, double2 *__restrict__ b, size_t n)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += stride * U)
    {
        double2 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) if (i + u * stride < n) v[u] = a[i + u * stride];
#pragma unroll
        for (int u = 0; u < U; ++u) if (i + u * stride < n) b[i + u * stride] = v[u];
    }
}
template <int U>
__global__ void __launch_bounds__(256) k_2r1w(double2 *__restrict__ r, const double2 *__restrict__ w, size_t n, double a, double *out)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    double acc = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += stride * U)
    {
        double2 rv[U], wv[U];
#pragma unroll
        for (int u = 0; u < U; ++u) if (i + u * stride < n) { rv[u] = r[i + u * stride]; wv[u] = w[i + u * stride]; }
#pragma unroll
        for (int u = 0; u < U; ++u) if (i + u * stride < n) { rv[u].x -= a * wv[u].x; rv[u].y -= a * wv[u].y; acc += rv[u].x * rv[u].x + rv[u].y * rv[u].y; r[i + u * stride] = rv[u]; }
    }
    if (acc == 123.456) *out = acc;
}
template <int U>
__global__ void __launch_bounds__(256) k_3r3w(const double2 *__restrict__ r, const double2 *__restrict__ p, double2 *__restrict__ x,
                                              double2 *__restrict__ po, double2 *__restrict__ w, size_t n, double a, double b)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += stride * U)
    {
        double2 rv[U], pv[U], xv[U];
#pragma unroll
        for (int u = 0; u < U; ++u) if (i + u * stride < n) { rv[u] = r[i + u * stride]; pv[u] = p[i + u * stride]; xv[u] = x[i + u * stride]; }
#pragma unroll
        for (int u = 0; u < U; ++u) if (i + u * stride < n)
        {
            double2 pn, xn, wn;
            pn.x = rv[u].x + b * pv[u].x; pn.y = rv[u].y + b * pv[u].y;
            xn.x = xv[u].x + a * pv[u].x; xn.y = xv[u].y + a * pv[u].y;
            wn.x = pn.x * 3.0 - rv[u].y; wn.y = pn.y * 3.0 - rv[u].x;
            x[i + u * stride] = xn; po[i + u * stride] = pn; w[i + u * stride] = wn;
        }
    }
}

int main(int argc, char **argv)
{
    const size_t N = (argc > 1) ? atol(argv[1]) : (size_t)256 * 256 * 256;
    const size_t n2 = N / 2;
    double *buf[6];
    for (auto &b : buf) { cudaMalloc(&b, N * 8); cudaMemset(b, 0, N * 8); }
    double *out; cudaMalloc(&out, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int reps = 20;
    auto time = [&](auto f, double bytes, const char *name, int blocks) {
        for (int q = 0; q < 3; ++q) f(blocks);
        cudaEventRecord(e0);
        for (int q = 0; q < reps; ++q) f(blocks);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("%-10s blocks %5d: %7.1f us  %7.0f GB/s\n", name, blocks, ms / reps * 1e3, bytes / (ms / reps * 1e-3) / 1e9);
    };
    for (int blocks : {148 * 2, 148 * 4, 148 * 8, 148 * 16, 148 * 32})
    {
        time([&](int b) { k_copy<4><<<b, 256>>>((double2 *)buf[0], (double2 *)buf[1], n2); }, 16.0 * N, "copy U4", blocks);
        time([&](int b) { k_copy<8><<<b, 256>>>((double2 *)buf[0], (double2 *)buf[1], n2); }, 16.0 * N, "copy U8", blocks);
        time([&](int b) { k_2r1w<4><<<b, 256>>>((double2 *)buf[0], (double2 *)buf[1], n2, 1e-3, out); }, 24.0 * N, "2r1w U4", blocks);
        time([&](int b) { k_2r1w<8><<<b, 256>>>((double2 *)buf[0], (double2 *)buf[1], n2, 1e-3, out); }, 24.0 * N, "2r1w U8", blocks);
        time([&](int b) { k_3r3w<2><<<b, 256>>>((double2 *)buf[0], (double2 *)buf[1], (double2 *)buf[2], (double2 *)buf[3], (double2 *)buf[4], n2, 1e-3, 0.5); }, 48.0 * N, "3r3w U2", blocks);
        time([&](int b) { k_3r3w<4><<<b, 256>>>((double2 *)buf[0], (double2 *)buf[1], (double2 *)buf[2], (double2 *)buf[3], (double2 *)buf[4], n2, 1e-3, 0.5); }, 48.0 * N, "3r3w U4", blocks);
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status %s\n", cudaGetErrorString(e));
    return 0;
}
