// cuda_emu.h -- TEST SCAFFOLDING: a single-OS-thread fiber emulation of the handful of CUDA constructs the
// kernels of petibm_b200/csrc use, so that their LOGIC (indexing, halos, ring buffers, reductions, the KSP state
// machine) can be run bit-for-bit against the oracle on a machine without a GPU.  Selected by -DB200_EMULATE
// through csrc/hw.cuh; compiled with g++ into tests/emu/libb200emu.so by tests/test_emulated_kernels.py only.
// It is NOT a fallback: nothing in petibm_b200/ or libb200ls.so can reach it, it models no timing and no memory
// model (beyond making cp.async copies complete as late as their wait allows).
#pragma once
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>

#include <algorithm>
#include <cmath>
#include <functional>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __grid_constant__
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))

using std::isinf;
using std::isnan;
using std::max;
using std::min;

struct uint3 { unsigned int x, y, z; };
struct dim3
{
    unsigned int x, y, z;
    dim3(unsigned int a = 1, unsigned int b = 1, unsigned int c = 1) : x(a), y(b), z(c) {}
};
struct __attribute__((aligned(16))) double2 { double x, y; };
inline double2 make_double2(double a, double b) { double2 r; r.x = a; r.y = b; return r; }

inline uint3 threadIdx, blockIdx;
inline dim3 blockDim, gridDim;

namespace emu {

enum { RUN = 0, AT_BAR = 1, DONE = 2 };
struct PendingCopy { void *dst; const void *src; int bytes; long group; };
struct Fiber
{
    ucontext_t ctx;
    char *stack = nullptr;
    int state = RUN;
    uint3 tid;
    int lane = 0, warp = 0;
    long ncoll = 0;               // warp collectives executed
    long committed = 0;           // cp.async groups committed
    int sleep = 0;                // schedule fuzzing: scheduler passes to sit out
    std::vector<PendingCopy> pending;
};
struct Warp
{
    unsigned long long buf[2][32];
    long slot_gen[2] = {-1, -1};
    int arrived[2] = {0, 0};
};
inline std::vector<Fiber> fibers;
inline std::vector<Warp> warps;
inline Fiber *cur = nullptr;
inline ucontext_t sched_ctx;
inline const std::function<void()> *body = nullptr;
inline int bar_arrived = 0;
inline unsigned char *dyn_smem = nullptr;
inline unsigned long long fake_clock = 0;
inline long progress = 0;
inline void (*block_end_check)() = nullptr;  // installed by the TMA emulation: no box copy may be left undelivered
// schedule fuzzing (emu_set_schedule): shuffled fiber order and random preemption at shared-memory accesses, so
// that a missing barrier shows up as a result that depends on the schedule
inline unsigned long long rng = 0;
inline bool fuzz = false;
inline unsigned int rnd()
{
    rng ^= rng << 13;
    rng ^= rng >> 7;
    rng ^= rng << 17;
    return (unsigned int)(rng >> 11);
}

inline void yield() { swapcontext(&cur->ctx, &sched_ctx); }
inline void maybe_yield()
{
    if (fuzz && (rnd() & 3u) == 0u)
    {
        ++progress;
        if ((rnd() & 15u) == 0u) cur->sleep = (int)(rnd() % 6u);  // occasionally fall far behind
        yield();
    }
}

inline void trampoline()
{
    (*body)();
    cur->state = DONE;
    ++progress;
    swapcontext(&cur->ctx, &sched_ctx);
}

inline int warp_active(int w)
{
    int n = 0;
    for (int l = 0; l < 32; ++l)
    {
        const size_t id = (size_t)w * 32 + l;
        if (id < fibers.size() && fibers[id].state != DONE) ++n;
    }
    return n;
}

// every active lane of the warp deposits a value; returns the 32 deposited values
inline const unsigned long long *warp_exchange(unsigned long long v)
{
    Fiber *f = cur;
    Warp &w = warps[f->warp];
    const long n = f->ncoll;
    const int g = (int)(n & 1);
    if (w.slot_gen[g] != n)
    {
        w.slot_gen[g] = n;
        w.arrived[g] = 0;
    }
    w.buf[g][f->lane] = v;
    w.arrived[g]++;
    ++progress;
    while (w.arrived[g] < warp_active(f->warp)) yield();
    f->ncoll = n + 1;
    return w.buf[g];
}

template <class F>
void launch(dim3 grid, dim3 block, size_t smem_bytes, F &&kernel_body)
{
    const std::function<void()> fn = kernel_body;
    body = &fn;
    const size_t nthr = (size_t)block.x * block.y * block.z;
    if (nthr % 32 != 0) { fprintf(stderr, "emu: block size must be a multiple of 32\n"); abort(); }
    std::vector<unsigned char> smem(smem_bytes + 64);
    dyn_smem = (unsigned char *)(((uintptr_t)smem.data() + 15) & ~(uintptr_t)15);
    const size_t STACK = 128 * 1024;
    static std::vector<char> stacks;
    if (stacks.size() < nthr * STACK) stacks.resize(nthr * STACK);
    gridDim = grid;
    blockDim = block;
    for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
            for (unsigned bx = 0; bx < grid.x; ++bx)
            {
                fibers.assign(nthr, Fiber());
                warps.assign(nthr / 32, Warp());
                bar_arrived = 0;
                size_t id = 0;
                for (unsigned tz = 0; tz < block.z; ++tz)
                    for (unsigned ty = 0; ty < block.y; ++ty)
                        for (unsigned tx = 0; tx < block.x; ++tx, ++id)
                        {
                            Fiber &f = fibers[id];
                            f.tid = {tx, ty, tz};
                            f.lane = (int)(id & 31);
                            f.warp = (int)(id >> 5);
                            f.stack = stacks.data() + id * STACK;
                            getcontext(&f.ctx);
                            f.ctx.uc_stack.ss_sp = f.stack;
                            f.ctx.uc_stack.ss_size = STACK;
                            f.ctx.uc_link = nullptr;
                            makecontext(&f.ctx, (void (*)())trampoline, 0);
                        }
                size_t done = 0;
                while (done < nthr)
                {
                    const long before = progress;
                    static std::vector<size_t> order;
                    order.resize(nthr);
                    for (size_t q = 0; q < nthr; ++q) order[q] = q;
                    if (fuzz)
                        for (size_t q = nthr - 1; q > 0; --q) std::swap(order[q], order[rnd() % (q + 1)]);
                    for (size_t oq = 0; oq < nthr; ++oq)
                    {
                        Fiber &f = fibers[order[oq]];
                        if (f.state != RUN) continue;
                        if (f.sleep > 0)
                        {
                            --f.sleep;
                            ++progress;
                            continue;
                        }
                        cur = &f;
                        threadIdx = f.tid;
                        blockIdx = {bx, by, bz};
                        swapcontext(&sched_ctx, &f.ctx);
                    }
                    done = 0;
                    int waiting = 0;
                    for (auto &f : fibers)
                    {
                        if (f.state == DONE) ++done;
                        if (f.state == AT_BAR) ++waiting;
                    }
                    if (waiting > 0 && (size_t)waiting == nthr - done)
                    {
                        for (auto &f : fibers)
                            if (f.state == AT_BAR) f.state = RUN;
                        bar_arrived = 0;
                        ++progress;
                    }
                    if (progress == before && done < nthr)
                    {
                        fprintf(stderr, "emu: deadlock (divergent barrier or warp collective) in block %u,%u,%u\n", bx, by, bz);
                        abort();
                    }
                }
                for (auto &f : fibers)
                    if (!f.pending.empty()) { fprintf(stderr, "emu: cp.async copies never waited for\n"); abort(); }
                if (block_end_check) block_end_check();
            }
    body = nullptr;
    cur = nullptr;
}

}  // namespace emu

// ---- CUDA builtins -------------------------------------------------------------------------------------
inline void __syncthreads()
{
    emu::cur->state = emu::AT_BAR;
    ++emu::progress;
    emu::yield();
}
inline void __syncwarp(unsigned = 0xffffffffu) { emu::warp_exchange(0); }
inline void __threadfence() {}
inline void __threadfence_system() {}
inline double __dmul_rn(double a, double b) { return a * b; }   // built with -ffp-contract=off
inline double __dadd_rn(double a, double b) { return a + b; }
inline double __ddiv_rn(double a, double b) { return a / b; }
inline long long __double_as_longlong(double d) { long long r; memcpy(&r, &d, 8); return r; }
inline double __longlong_as_double(long long v) { double r; memcpy(&r, &v, 8); return r; }
inline double __ldcg(const double *p) { return *p; }
inline unsigned int atomicAdd(unsigned int *p, unsigned int v) { const unsigned int o = *p; *p = o + v; return o; }

template <class T>
inline T __shfl_sync(unsigned, T v, int src)
{
    unsigned long long raw = 0;
    memcpy(&raw, &v, sizeof(T));
    const unsigned long long *all = emu::warp_exchange(raw);
    T r;
    memcpy(&r, &all[src & 31], sizeof(T));
    return r;
}
template <class T>
inline T __shfl_xor_sync(unsigned m, T v, int lanemask) { return __shfl_sync(m, v, emu::cur->lane ^ lanemask); }
inline bool __all_sync(unsigned, bool p)
{
    const unsigned long long *all = emu::warp_exchange(p ? 1ull : 0ull);
    bool r = true;
    for (int l = 0; l < 32; ++l)
    {
        const size_t id = (size_t)emu::cur->warp * 32 + l;
        if (id < emu::fibers.size() && emu::fibers[id].state != emu::DONE) r = r && (all[l] != 0);
    }
    return r;
}

// ---- the primitives of csrc/hw.cuh ---------------------------------------------------------------------
namespace b200 {

inline unsigned long long global_timer_ns() { return ++emu::fake_clock; }
inline void st_release_sys(unsigned long long *p, unsigned long long v) { *p = v; }
inline unsigned long long ld_acquire_sys(const unsigned long long *p) { return *p; }
inline double ld_volatile(const double *p) { return *p; }
inline void pdl_sync() {}
inline void st_relaxed_sys_u64(unsigned long long *p, unsigned long long v) { *p = v; }
inline unsigned long long ld_relaxed_sys_u64(const unsigned long long *p) { return *p; }

#define B200_DYNAMIC_SMEM(name) unsigned char *name = emu::dyn_smem
inline unsigned int smem_u32(const void *p) { return (unsigned int)((const unsigned char *)p - emu::dyn_smem); }
inline unsigned char *smem_ptr(unsigned int a) { return emu::dyn_smem + a; }

// cp.async: the copy is performed as LATE as the matching wait allows (a read before the wait sees stale data)
inline void cp_async16(unsigned int dst, const void *src) { emu::cur->pending.push_back({smem_ptr(dst), src, 16, emu::cur->committed}); }
inline void cp_async8(unsigned int dst, const void *src) { emu::cur->pending.push_back({smem_ptr(dst), src, 8, emu::cur->committed}); }
inline void cp_async_commit() { emu::cur->committed++; }
template <int N>
inline void cp_async_wait()
{
    emu::maybe_yield();
    auto &pend = emu::cur->pending;
    const long limit = emu::cur->committed - N;  // groups < limit must be complete
    size_t keep = 0;
    for (size_t q = 0; q < pend.size(); ++q)
    {
        if (pend[q].group < limit) memcpy(pend[q].dst, pend[q].src, (size_t)pend[q].bytes);
        else pend[keep++] = pend[q];
    }
    pend.resize(keep);
}
// mbarrier: {outstanding transaction bytes, pending arrivals, expected count, parity of the current (incomplete)
// phase}; an arrival always counts towards the CURRENT phase (as on hardware), a phase completes when both the
// arrivals and the transaction bytes have reached zero; protocol errors show up as wrong results / deadlocks
struct EmuMbar { int tx : 31; unsigned int phase : 1; unsigned short pending; unsigned short count; };
static_assert(sizeof(EmuMbar) == 8, "an mbarrier is one 64-bit shared-memory word");
inline void mbar_check(EmuMbar *m)
{
    if (m->pending == 0 && m->tx == 0)
    {
        m->phase ^= 1;
        m->pending = m->count;
    }
}
inline void mbar_init(unsigned int a, unsigned int count)
{
    EmuMbar m = {0, 0u, (unsigned short)count, (unsigned short)count};
    memcpy(smem_ptr(a), &m, 8);
}
inline void mbar_init_fence() {}
inline void proxy_async_fence() {}
inline void mbar_arrive(unsigned int a)
{
    emu::maybe_yield();
    EmuMbar *m = reinterpret_cast<EmuMbar *>(smem_ptr(a));
    --m->pending;
    mbar_check(m);
    ++emu::progress;
}
inline void mbar_arrive_expect_tx(unsigned int a, unsigned int bytes)
{
    emu::maybe_yield();
    EmuMbar *m = reinterpret_cast<EmuMbar *>(smem_ptr(a));
    m->tx += (int)bytes;
    --m->pending;
    mbar_check(m);
    ++emu::progress;
}
// TMA: a 3-D box copy global -> shared with zero fill outside the tensor; like cp.async it is performed as LATE as
// possible -- when a thread waits on the barrier the copy reports to and finds the phase incomplete
struct TmaMap { const double *base; int dim[3]; long long stride[3]; int box[3]; };  // strides in elements
struct EmuPendingTma { unsigned int dst, mbar; TmaMap map; int c[3]; };
inline std::vector<EmuPendingTma> emu_tma_pending;
inline void tma_prefetch_desc(const TmaMap *) {}
inline void tma_load_3d(unsigned int dst, const TmaMap *map, unsigned int mbar, int c0, int c1, int c2)
{
    if (dst % 128 != 0) { fprintf(stderr, "emu: TMA destination not 128-byte aligned\n"); abort(); }
    if ((map->box[0] * 8) % 16 != 0) { fprintf(stderr, "emu: TMA inner box extent not a multiple of 16 bytes\n"); abort(); }
    emu_tma_pending.push_back({dst, mbar, *map, {c0, c1, c2}});
    emu::block_end_check = [] {
        if (!emu_tma_pending.empty()) { fprintf(stderr, "emu: a TMA copy was still in flight when the CTA exited\n"); abort(); }
    };
    ++emu::progress;
}
inline void emu_tma_deliver(unsigned int mbar)
{
    size_t keep = 0;
    for (size_t q = 0; q < emu_tma_pending.size(); ++q)
    {
        const EmuPendingTma &t = emu_tma_pending[q];
        if (t.mbar != mbar) { emu_tma_pending[keep++] = t; continue; }
        double *dst = reinterpret_cast<double *>(smem_ptr(t.dst));
        const TmaMap &m = t.map;
        for (int z = 0; z < m.box[2]; ++z)
            for (int y = 0; y < m.box[1]; ++y)
                for (int x = 0; x < m.box[0]; ++x)
                {
                    const int gx = t.c[0] + x, gy = t.c[1] + y, gz = t.c[2] + z;
                    const bool in = gx >= 0 && gx < m.dim[0] && gy >= 0 && gy < m.dim[1] && gz >= 0 && gz < m.dim[2];
                    *dst++ = in ? m.base[gx * m.stride[0] + gy * m.stride[1] + gz * m.stride[2]] : 0.0;
                }
        EmuMbar *b = reinterpret_cast<EmuMbar *>(smem_ptr(mbar));
        b->tx -= 8 * m.box[0] * m.box[1] * m.box[2];
        mbar_check(b);
        ++emu::progress;
    }
    emu_tma_pending.resize(keep);
}
inline void mbar_wait(unsigned int a, unsigned int parity)
{
    const EmuMbar *m = reinterpret_cast<const EmuMbar *>(smem_ptr(a));
    if (m->phase == parity) emu_tma_deliver(a);
    while (m->phase == parity)
    {
        emu::yield();
        if (m->phase == parity) emu_tma_deliver(a);
    }
    emu::maybe_yield();
}
inline double2 lds128(unsigned int a) { emu::maybe_yield(); double2 v; memcpy(&v, smem_ptr(a), 16); return v; }
inline double lds64(unsigned int a) { emu::maybe_yield(); double v; memcpy(&v, smem_ptr(a), 8); return v; }
inline void sts128(unsigned int a, double2 v) { emu::maybe_yield(); memcpy(smem_ptr(a), &v, 16); }
inline void sts64(unsigned int a, double v) { emu::maybe_yield(); memcpy(smem_ptr(a), &v, 8); }

}  // namespace b200
