// hw.cuh -- every hardware-specific primitive the kernels use (inline PTX, dynamic shared memory), in one
// place.  Building with -DB200_EMULATE (tests/emu only, g++, no GPU) swaps this file for a single-threaded
// fiber emulation of the same names so that the kernel LOGIC of kernels.cuh / spmv2.cuh / csr_kernels.cuh can
// be exercised bit-for-bit against the oracle on a CPU-only machine.  The emulation is test scaffolding: it is
// never compiled into libb200ls.so.
#pragma once
#ifdef B200_EMULATE
#include "cuda_emu.h"
#else
#include <cuda.h>  // CUtensorMap (type only: the encoder is fetched through cudaGetDriverEntryPoint, no libcuda link)
#include <cuda_runtime.h>

namespace b200 {

// ---- TMA (cp.async.bulk.tensor, sm_90+): one elected thread moves a whole box global -> shared; completion is
// signalled on an mbarrier by transaction bytes.  Out-of-bounds elements of the box are filled with zeros.
struct alignas(64) TmaMap { CUtensorMap m; };
__device__ __forceinline__ void tma_prefetch_desc(const TmaMap *map)
{
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tma_load_3d(unsigned int smem_dst, const TmaMap *map, unsigned int mbar, int c0, int c1, int c2)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_dst),
        "l"(map), "r"(mbar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned int a, unsigned int bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes) : "memory");
}
// orders this thread's earlier generic-proxy accesses (an acquire load of a flag) before its later async-proxy copies
__device__ __forceinline__ void proxy_async_fence() { asm volatile("fence.proxy.async;" ::: "memory"); }
// after mbarrier.init, before the barriers are used by the async proxy (TMA)
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

__device__ __forceinline__ unsigned long long global_timer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ double ld_volatile(const double *p)
{
    double v;
    asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}

// Programmatic dependent launch (sm_90+): a kernel launched with the programmatic-serialization attribute
// may start while its predecessor is still draining; everything before pdl_sync() must not touch memory the
// predecessor writes.  The trigger is issued AFTER the wait, so at most two grids are ever co-resident.
__device__ __forceinline__ void pdl_sync()
{
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

__device__ __forceinline__ void st_relaxed_sys_u64(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys_u64(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// ---- asynchronous global->shared copies (LDGSTS) and 32-bit-addressed shared-memory accesses
__device__ __forceinline__ void cp_async16(unsigned int smem_dst, const void *gsrc)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async8(unsigned int smem_dst, const void *gsrc)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait()
{
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ double2 lds128(unsigned int a)
{
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ double lds64(unsigned int a)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts128(unsigned int a, double2 v)
{
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(a), "d"(v.x), "d"(v.y) : "memory");
}
__device__ __forceinline__ void sts64(unsigned int a, double v)
{
    asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory");
}

// ---- mbarrier (shared-memory barrier object, sm_80+): arrive and wait are separate operations, so a CTA-wide
// barrier can be ARRIVED at in step t and WAITED for in step t+1 (k_spmv3 SPLITBAR)
__device__ __forceinline__ void mbar_init(unsigned int a, unsigned int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned int a)
{
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(a) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned int a, unsigned int parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n"
        "MBAR_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra MBAR_DONE;\n\t"
        "bra MBAR_WAIT;\n"
        "MBAR_DONE:\n\t}" ::"r"(a), "r"(parity) : "memory");
}

#define B200_DYNAMIC_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
__device__ __forceinline__ unsigned int smem_u32(const void *p) { return (unsigned int)__cvta_generic_to_shared(p); }

}  // namespace b200
#endif
