// kernels.cuh -- sm_100a device code of the B200 pressure-Poisson solver.
//
// What the reference does per CG iteration inside PETSc (KSPSolve_CG, cg.c; called from
// LinSolverKSP::solve, src/linsolver/linsolverksp.cpp:85-105) as ~9 separate memory passes --
// MatMult_MPIAIJ, VecXDot, 2x VecAXPY, PCApply, MatNullSpaceRemove (VecSum + VecShift), VecNorm,
// VecXDot, VecAYPX -- is expressed here as TWO kernels and 72 B/row of HBM traffic:
//
//   class 0 (k_spmv2 in spmv2.cuh; k_spmv below is its first generation, k_spmv3 a round-2 candidate):
//                        x <- x + a' p'   (the PREVIOUS iteration's VecAXPY(X,a,P), deferred so that
//                                          p' is read once for both uses)
//                        p <- z + b p'    (z = B r + shift recomputed on the fly, halo included)
//                        w <- A p         (matrix-free 5/7-point stencil from 1-D arrays)
//                        dpi <- p.w       (warp-shuffle -> block -> last-block reduction)
//                        a <- beta/dpi    (+ KSP_DIVERGED_INDEFINITE_MAT test) by the last block
//                        reads r,p',x  writes p,w,x                                   = 48 B/row
//   class 1 (k_update2 below; k_update is its first generation):
//                        r <- r - a w
//                        {sum z0, sum d, sum d^2, sum d r, sum r, sum r^2}, d = z0 - c
//                        -> shift, dp = ||z||, beta = z.r, KSPConvergedDefault, b = beta/betaold
//                        (+ multi-GPU: boundary planes of r stored straight into the neighbour's
//                        ghost planes over NVLink, partial sums all-reduced through peer mailboxes)
//                        reads r,w  writes r                                          = 24 B/row
//
// All floating-point operations that define the *operator* and the *vector updates* are written
// with explicit __dmul_rn/__dadd_rn (no FMA contraction) in the operand order of the reference's
// assembled CSR path, so that the matrix-free SpMV is bit-identical to MatMult on D*(dt*G)
// (oracle: orc_assemble_dbng_literal + orc_spmv).  Only the reduction
// order of the dot products differs from a serial CPU sum.
//
// File map: hw.cuh (inline PTX + the emulation switch), kernels.cuh (state machine of cg.c / iterativ.c,
// grid reduction + LL mailbox all-reduce, update kernels, layout helpers), spmv2.cuh / spmv3.cuh (fused SpMV),
// csr_kernels.cuh (assembled-operator path, BiCGStab).
#pragma once
#include <stdint.h>

#include "hw.cuh"

namespace b200 {

// ------------------------------------------------------------------------------------------
// device-resident solver state (what cg.c keeps in local variables + KSP fields)
// ------------------------------------------------------------------------------------------
struct DevState
{
    double beta, betaold, dpi, dpiold, a, b;
    double shift;  // z = z0 + shift (constant null-space removal), 0 without null space
    double c;      // centre of the shifted accumulation (= -shift of the previous reduction)
    double dp, rnorm0, ttol;
    // BiCGStab scalars (bcgs.c): rho, rhoold, alpha, omega, omegaold, d1 = (s,t), d2 = (t,t)
    double rho, rhoold, alpha, omega, omegaold;
    unsigned long long seq;  // number of cross-GPU reductions performed so far
    int i;        // loop counter of KSPSolve_CG
    int its;      // ksp->its
    int reason;   // KSPConvergedReason, 0 while iterating
    int done;     // reason != 0 (or internal error)
    int nhist;    // entries written to the history
    int pending;  // x <- x + a p of the last completed iteration has not been applied yet
    int pcur;     // which of the two p buffers holds the current search direction
    int err;      // internal error: 1 = cross-GPU reduction timed out
};

struct SolveConsts
{
    double rtol, atol, divtol;
    double nglobal;  // global number of unknowns as a double (-1.0*N in MatNullSpaceRemove)
    int max_it, norm_type, has_const, hist_cap;
};

#define B200_MAX_RANKS 16
#define B200_NSUM 8  // doubles per reduction record

// cross-GPU plumbing seen by the kernels
struct CommDev
{
    int rank, nranks;
    int mode;                    // 0 = single GPU, 1 = P2P mailboxes, 2 = NCCL (sums left in sendbuf)
    // peer q's mailbox base (mapped over CUDA IPC): [parity][src rank][2*NSUM] 8-byte words, each word
    // = {flag: low 32 bits of the reduction sequence number, data: one 32-bit half of a double}
    unsigned long long *mbox_peer[B200_MAX_RANKS];
    unsigned long long *mbox_local;
    double *sendbuf;             // NCCL mode: local sums
    double *r_ghost_dn;          // neighbour below: address of ITS top ghost plane of r (or null)
    double *r_ghost_up;          // neighbour above: address of ITS bottom ghost plane of r (or null)
    // halo hand-shake of the fused transport (mode 1): the pusher CTAs of k_update2 publish "my boundary plane of reduction
    // number s has landed in your ghost plane" in the neighbours' arenas AFTER their system-scope fence, off the path of
    // the scalar all-reduce; the next fused SpMV kernel waits for the flags of its two ghost planes before reading them
    unsigned long long *halo_flag_local;    // [0]: ghost plane below is ready (written by rank-1), [1]: above (rank+1); or null
    unsigned long long *halo_flag_dn_peer;  // neighbour below: its flag [1]
    unsigned long long *halo_flag_up_peer;  // neighbour above: its flag [0]
    unsigned int *push_counter;             // pusher CTAs that have fenced (local)
    int push_items;              // items per pusher thread of k_update2 (0 = 2); experiment knob B200LS_PUSH_ITEMS
    int npush;                   // pusher CTAs the host ADDED to the update grid (0: the kernel takes them out of its grid)
    int dbg_flags;               // bit 0: skip the system-scope fence of the pushers (TIMING EXPERIMENTS ONLY: unsafe)
};


struct ReduceWs
{
    double *partials;        // [max_blocks][B200_NSUM]
    unsigned int *counter;   // ticket counter, self-resetting
    // optional timeline (b200ls_set_trace): [0] = entries written, [1] = start stamp of the running kernel,
    // [2] = capacity, entries of 5 u64 from [8]: {kernel start, local reduction done, all-reduce done,
    // scalars done, kind} in globaltimer ns
    unsigned long long *trace;
};

__device__ __forceinline__ void trace_kernel_start(const ReduceWs &ws)
{
    if (ws.trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 0 && threadIdx.y == 0)
        ws.trace[1] = global_timer_ns();
}



__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ------------------------------------------------------------------------------------------
// scalar logic of KSPSolve_CG, executed by ONE thread after a reduction is complete
// ------------------------------------------------------------------------------------------
enum { FIN_SPMV = 0, FIN_UPDATE = 1, FIN_INIT_CENTRE = 2, FIN_INIT = 3 };

// KSPConvergedDefault (iterativ.c)
__device__ inline int converged_default(DevState &s, const SolveConsts &k, int it, double rnorm)
{
    if (it == 0)
    {
        s.rnorm0 = rnorm;
        s.ttol = fmax(k.rtol * rnorm, k.atol);
    }
    if (isnan(rnorm) || isinf(rnorm)) return -9;
    if (rnorm <= s.ttol) return (rnorm < k.atol) ? 3 : 2;
    if (rnorm >= k.divtol * s.rnorm0) return -4;
    return 0;
}

__device__ inline void finalize_scalars(int kind, const double *S, DevState &s, const SolveConsts &k,
                                        double *hist)
{
    if (kind == FIN_SPMV)
    {
        // the kernel has applied the deferred x update and produced the new search direction
        s.pending = 0;
        s.pcur ^= 1;
        // cg.c: dpiold = dpi; dpi = p.w; betaold = beta; indefinite test; a = beta/dpi
        s.dpiold = s.dpi;
        s.dpi = S[0];
        s.betaold = s.beta;
        const double sg = (s.dpi >= 0.0 ? 1.0 : -1.0) * (s.dpiold >= 0.0 ? 1.0 : -1.0);
        if (isnan(s.dpi) || isinf(s.dpi))
        {
            s.reason = -9;  // KSPCheckDot
            s.done = 1;
        }
        else if (s.dpi == 0.0 || (s.i > 0 && sg < 0.0))
        {
            s.reason = -10;  // KSP_DIVERGED_INDEFINITE_MAT
            s.done = 1;
        }
        else
            s.a = s.beta / s.dpi;
        return;
    }
    // sums of the update / init reduction
    const double Sz0 = S[0], Sd = S[1], Sdd = S[2], Sdr = S[3], Sr = S[4], Srr = S[5];
    // MatNullSpaceRemove: sum = VecSum(z)/(-1.0*N); VecShift(z, sum)
    const double shift = k.has_const ? Sz0 / (-1.0 * k.nglobal) : 0.0;
    if (kind == FIN_INIT_CENTRE)
    {
        s.c = -shift;  // second init pass accumulates around the mean: no cancellation
        return;
    }
    if (kind == FIN_UPDATE) s.pending = 1;  // r was advanced with a; x += a p is still owed
    const double e = s.c + shift;  // z_i = d_i + e
    const double zz = Sdd + 2.0 * e * Sd + k.nglobal * e * e;
    const double zr = Sdr + e * Sr;
    s.shift = shift;
    s.c = -shift;
    double dp;
    if (k.norm_type == 1) dp = sqrt(fmax(zz, 0.0));
    else if (k.norm_type == 2) dp = sqrt(Srr);
    else if (k.norm_type == 3)
    {
        s.beta = zr;
        dp = sqrt(fabs(zr));
    }
    else dp = 0.0;
    if (isnan(zz) || isnan(zr) || isnan(Srr)) dp = zz + zr + Srr;  // propagate NaN like KSPCheckNorm
    s.dp = dp;
    if (s.nhist < k.hist_cap) hist[s.nhist] = dp;
    s.nhist++;
    const int it = (kind == FIN_INIT) ? 0 : s.i + 1;
    int reason = converged_default(s, k, it, dp);
    if (!reason)
    {
        if (k.norm_type != 3) s.beta = zr;
        if (kind == FIN_INIT) s.i = 0;
        else s.i = s.i + 1;
        if (kind != FIN_INIT && s.i >= k.max_it) reason = -3;  // KSP_DIVERGED_ITS
        else if (k.max_it <= 0) reason = -3;
        else
        {
            // top of the do-loop for the next iteration
            s.its = s.i + 1;
            if (s.beta == 0.0) reason = 3;                                  // "converged due to beta = 0"
            else if (s.i > 0 && s.beta * s.betaold < 0.0) reason = -8;      // indefinite PC
            else s.b = (s.i == 0) ? 0.0 : s.beta / s.betaold;
        }
    }
    else if (kind != FIN_INIT)
        s.its = s.i + 1;
    if (reason)
    {
        s.reason = reason;
        s.done = 1;
    }
}

// ------------------------------------------------------------------------------------------
// cross-GPU all-reduce of one B200_NSUM record through peer-mapped mailboxes (one warp).
// Low-latency ("LL") protocol: every 8-byte word carries 32 bits of payload and the 32-bit sequence
// number, so a word is valid the moment its flag matches -- one NVLink store hop, no fence and no
// separate flag write on the critical path.  Every rank adds the nranks records in rank order, so the
// result is bit-identical everywhere.  Slots are double-buffered by the parity of the sequence number.
// ------------------------------------------------------------------------------------------
#define B200_LLW (2 * B200_NSUM)  // words per record


// S: this rank's sums on entry (valid in every lane), the global sums on exit (valid in lane 0).
// seq: sequence number of THIS reduction (previous + 1), identical on every rank.
__device__ __forceinline__ bool mailbox_allreduce(double (&S)[B200_NSUM], const CommDev &cm,
                                                  unsigned long long seq, int lane, unsigned int *s_ll)
{
    const int par = (int)(seq & 1ull);
    const unsigned int flag = (unsigned int)seq;
    const int nr = cm.nranks;
    // ---- send: lane l < 16 owns word l of the record (double l/2, half l%2)
    {
        unsigned int half = 0u;
#pragma unroll
        for (int q = 0; q < B200_NSUM; ++q)
        {
            const unsigned long long bits = (unsigned long long)__double_as_longlong(S[q]);
            if ((lane >> 1) == q) half = (lane & 1) ? (unsigned int)(bits >> 32) : (unsigned int)bits;
        }
        const unsigned long long word = ((unsigned long long)flag << 32) | (unsigned long long)half;
        if (lane < B200_LLW)
            for (int p = 0; p < nr; ++p)
                st_relaxed_sys_u64(cm.mbox_peer[p] + ((size_t)par * nr + cm.rank) * B200_LLW + lane, word);
    }
    // ---- receive: nr * 16 words, strided over the warp
    bool ok = true;
    const unsigned long long *base = cm.mbox_local + (size_t)par * nr * B200_LLW;
    const unsigned long long t0 = global_timer_ns();
    for (int idx = lane; idx < nr * B200_LLW; idx += 32)
    {
        unsigned long long w = ld_relaxed_sys_u64(base + idx);
        unsigned int spins = 0;
        while ((unsigned int)(w >> 32) != flag)
        {
            if ((++spins & 1023u) == 0 && global_timer_ns() - t0 > 4000000000ull)
            {
                ok = false;  // a peer never arrived: give up instead of hanging the GPU
                break;
            }
            w = ld_relaxed_sys_u64(base + idx);
        }
        s_ll[idx] = (unsigned int)w;
    }
    ok = __all_sync(0xffffffffu, ok);
    __syncwarp();
    // ---- fixed-order sum over the ranks: lane q < NSUM builds sum q, lane 0 collects
    double mine = 0.0;
    if (lane < B200_NSUM)
        for (int src = 0; src < nr; ++src)
        {
            const unsigned int lo = s_ll[src * B200_LLW + 2 * lane], hi = s_ll[src * B200_LLW + 2 * lane + 1];
            mine += __longlong_as_double((long long)(((unsigned long long)hi << 32) | lo));
        }
#pragma unroll
    for (int q = 0; q < B200_NSUM; ++q) S[q] = __shfl_sync(0xffffffffu, mine, q);
    return ok;
}

// One thread: wait until the neighbours' boundary planes of reduction number `seq` have landed in this rank's ghost planes
// (see CommDev::halo_flag_*).  Bounded like the mailbox spin: a peer that never arrives also fails the next all-reduce,
// which reports the error.
__device__ __forceinline__ void halo_wait_thread(const CommDev &cm, unsigned long long seq, bool need_dn, bool need_up)
{
    if (cm.mode != 1 || cm.halo_flag_local == nullptr) return;
    const unsigned long long t0 = global_timer_ns();
    for (int side = 0; side < 2; ++side)
    {
        const bool need = side == 0 ? (need_dn && cm.halo_flag_dn_peer != nullptr) : (need_up && cm.halo_flag_up_peer != nullptr);
        if (!need) continue;
        unsigned int spins = 0;
        while (ld_acquire_sys(cm.halo_flag_local + side) < seq)
            if ((++spins & 1023u) == 0 && global_timer_ns() - t0 > 4000000000ull) break;
    }
}
// CTA-wide form: thread 0 waits, everybody meets at a barrier
__device__ __forceinline__ void halo_wait_cta(const CommDev &cm, unsigned long long seq, bool need_dn, bool need_up)
{
    if (cm.mode != 1 || cm.halo_flag_local == nullptr) return;
    if (threadIdx.x == 0 && threadIdx.y == 0 && threadIdx.z == 0) halo_wait_thread(cm, seq, need_dn, need_up);
    __syncthreads();
}

// ------------------------------------------------------------------------------------------
// block reduction + deterministic grid reduction + (multi-GPU) all-reduce + scalar logic
// ------------------------------------------------------------------------------------------
constexpr int kStateWords = (int)((sizeof(DevState) + 7) / 8);
static_assert(kStateWords <= 32, "DevState must fit one warp-wide staging pass");

template <int NS>
__device__ __forceinline__ void grid_reduce_finalize(double (&acc)[NS], int kind, const ReduceWs &ws,
                                                     const CommDev &cm, DevState *st,
                                                     const SolveConsts &k, double *hist, bool multi_fence)
{
    __shared__ double s_red[32][NS];
    __shared__ bool s_last;
    __shared__ __align__(8) unsigned long long s_state[32];  // staged copy of *st (constant during the kernel)
    __shared__ unsigned int s_ll[B200_MAX_RANKS * B200_LLW];
    const int tid = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
    const int nthr = blockDim.x * blockDim.y * blockDim.z;
    const int lane = tid & 31, wid = tid >> 5, nw = (nthr + 31) >> 5;
    const unsigned int nblocks = gridDim.x * gridDim.y * gridDim.z;
    const unsigned int bid = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
    // every block prefetches the solver state (168 B): whichever block turns out to be last finds it in
    // shared memory instead of paying serial round trips to L2 after the reduction
    if (tid < kStateWords) s_state[tid] = reinterpret_cast<const unsigned long long *>(st)[tid];
#pragma unroll
    for (int q = 0; q < NS; ++q)
    {
        const double v = warp_sum(acc[q]);
        if (lane == 0) s_red[wid][q] = v;
    }
    __syncthreads();
    if (wid == 0)
    {
#pragma unroll
        for (int q = 0; q < NS; ++q)
        {
            double v = (lane < nw) ? s_red[lane][q] : 0.0;
            v = warp_sum(v);
            if (lane == 0) ws.partials[(size_t)bid * B200_NSUM + q] = v;
        }
    }
    // make this block's partials (and, multi-GPU, its peer halo stores) visible, then take a ticket
    if (multi_fence) __threadfence_system();
    else __threadfence();
    __syncthreads();
    if (tid == 0)
    {
        const unsigned int t = atomicAdd(ws.counter, 1u);
        s_last = (t == nblocks - 1);
    }
    __syncthreads();
    if (!s_last) return;
    if (cm.mode == 1) __threadfence_system();  // order every block's peer stores before our mailbox words
    else __threadfence();
    // last block: fixed-order sum of all block partials
    double tot[NS];
#pragma unroll
    for (int q = 0; q < NS; ++q) tot[q] = 0.0;
    // four independent loads in flight per thread and sum (each costs an L2 round trip); the grouping of the additions
    // is a fixed function of (nblocks, block size), so the result stays deterministic
    {
        unsigned int b = tid;
        for (; b + 3u * nthr < nblocks; b += 4u * nthr)
        {
            double v[4][NS];
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int q = 0; q < NS; ++q) v[u][q] = __ldcg(&ws.partials[(size_t)(b + u * nthr) * B200_NSUM + q]);
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int q = 0; q < NS; ++q) tot[q] += v[u][q];
        }
        for (; b < nblocks; b += nthr)
        {
#pragma unroll
            for (int q = 0; q < NS; ++q) tot[q] += __ldcg(&ws.partials[(size_t)b * B200_NSUM + q]);
        }
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < NS; ++q)
    {
        const double v = warp_sum(tot[q]);
        if (lane == 0) s_red[wid][q] = v;
    }
    __syncthreads();
    if (wid != 0) return;
    double S[B200_NSUM];
#pragma unroll
    for (int q = 0; q < B200_NSUM; ++q) S[q] = 0.0;
#pragma unroll
    for (int q = 0; q < NS; ++q)
    {
        double v = (lane < nw) ? s_red[lane][q] : 0.0;
        S[q] = warp_sum(v);
    }
    if (lane == 0) *ws.counter = 0u;
    unsigned long long tr1 = 0, tr2 = 0;
    if (ws.trace && lane == 0) tr1 = global_timer_ns();
    if (cm.mode == 2)
    {
        // NCCL transport: leave the local sums for ncclAllReduce; k_scalars finishes the job
        if (lane < B200_NSUM) cm.sendbuf[lane] = S[lane];
        return;
    }
    DevState &sst = *reinterpret_cast<DevState *>(s_state);
    if (cm.mode == 1)
    {
        const unsigned long long seq = sst.seq + 1;
        const bool ok = mailbox_allreduce(S, cm, seq, lane, s_ll);
        if (lane == 0)
        {
            sst.seq = seq;
            if (!ok)
            {
                sst.err = 1;
                sst.done = 1;
            }
        }
        if (!ok)
        {
            __syncwarp();
            if (lane < kStateWords) reinterpret_cast<unsigned long long *>(st)[lane] = s_state[lane];
            return;
        }
    }
    if (ws.trace && lane == 0) tr2 = global_timer_ns();
    if (lane == 0) finalize_scalars(kind, S, sst, k, hist);
    __syncwarp();
    if (lane < kStateWords) reinterpret_cast<unsigned long long *>(st)[lane] = s_state[lane];
    if (ws.trace && lane == 0)
    {
        const unsigned long long idx = ws.trace[0];
        if (idx < ws.trace[2])
        {
            unsigned long long *e = ws.trace + 8 + 5 * idx;
            e[0] = ws.trace[1];
            e[1] = tr1;
            e[2] = tr2;
            e[3] = global_timer_ns();
            e[4] = (unsigned long long)kind;
        }
        ws.trace[0] = idx + 1;
    }
}

// NCCL transport: scalar logic after ncclAllReduce
__global__ void k_scalars(int kind, const double *recv, DevState *st, SolveConsts k, double *hist)
{
    if (threadIdx.x == 0 && blockIdx.x == 0)
    {
        if (st->done) return;
        double S[B200_NSUM];
        for (int q = 0; q < B200_NSUM; ++q) S[q] = recv[q];
        finalize_scalars(kind, S, *st, k, hist);
    }
}

__global__ void k_state_reset(DevState *st)
{
    if (threadIdx.x == 0 && blockIdx.x == 0)
    {
        const unsigned long long seq = st->seq;
        DevState z = {};
        z.seq = seq;
        z.betaold = 1.0;
        z.a = 1.0;
        *st = z;
    }
}

// ------------------------------------------------------------------------------------------
// operator description
// ------------------------------------------------------------------------------------------
struct GridDev
{
    int nx, ny, nzl;      // local interior extents (nzl planes owned by this rank)
    int px;               // row pitch in doubles (multiple of 16)
    long long plane;      // plane stride = px * ny
    int perx, pery;       // periodic flags handled by index wrap
    int perz_wrap;        // 1: single-GPU periodic z (ghost plane -1 aliases plane nzl-1 by index)
    int kz0;              // global index of local plane 0
    int nzg;              // global nz
    int wrapz_lo, wrapz_hi;  // this rank owns global plane 0 / nz-1 of a periodic-z grid
    const double *dx, *dy, *dz;  // cell widths (global axes)
    const double *gx, *gy, *gz;  // face coefficients dt*(1/h), shifted by one: g[s] = minus face of cell s,
                                 // g[s+1] = plus face; wall faces hold 0, periodic wrap faces the wrap value
};

struct VecSet
{
    const double *r;      // residual (with ghost planes); APPLY: the input vector
    const double *p_in;   // previous search direction (with ghost planes)
    double *p_out;        // new search direction
    double *w;            // A p
    double *x;            // solution (no ghost planes needed, same indexing)
    const double *dinv;   // Jacobi: 1/diag (with ghost planes) or null
};

// storage plane of local plane kk in [-1, nzl]
__device__ __forceinline__ int zstore(const GridDev &g, int kk)
{
    if (g.perz_wrap)
    {
        if (kk < 0) kk += g.nzl;
        else if (kk >= g.nzl) kk -= g.nzl;
    }
    return kk + 1;
}

// ------------------------------------------------------------------------------------------
// class 0: fused  x <- x + a' p' ; p <- z + b p' ; w <- A p ; dpi <- p.w
// Tile: BX = 2*TXT points in x, TY = TYT-2 rows in y, marching over a chunk of z planes.
// Thread rows 0 and TYT-1 carry the y-halo rows (they build p but no w).
// ------------------------------------------------------------------------------------------
template <int TXT, int TYT, bool JACOBI, bool APPLY>
__global__ void __launch_bounds__(TXT *TYT) k_spmv(GridDev g, VecSet v, int kz_chunk, ReduceWs ws, CommDev cm,
                                                   DevState *st, SolveConsts kc, double *hist, int ghost_store)
{
    constexpr int BX = 2 * TXT;
    constexpr int TY = TYT - 2;
    constexpr int SROW = BX + 4;  // [1] left halo, [2..BX+1] tile, [BX+2] right halo
    __shared__ __align__(16) double sp[3][TYT][SROW];

    pdl_sync();
    trace_kernel_start(ws);
    double shift = 0.0, bcoef = 0.0, aprev = 0.0;
    bool xupd = false;
    if (!APPLY)
    {
        if (st->done) return;
        shift = st->shift;
        bcoef = st->b;
        aprev = st->a;
        xupd = st->pending != 0;
    }
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int i0 = blockIdx.x * BX, j0 = blockIdx.y * TY;
    const int k0 = blockIdx.z * kz_chunk;
    const int k1 = min(k0 + kz_chunk, g.nzl);
    if (!APPLY) halo_wait_cta(cm, st->seq, k0 == 0, k1 == g.nzl);  // ghost planes of r from the neighbour GPUs
    const int i = i0 + 2 * tx;  // first of the two x points of this thread
    const int jr = j0 - 1 + ty; // row (may be -1 or >= ny)
    // row mapping (wrap if periodic, clamp otherwise: clamped rows only meet zero coefficients)
    int jm = jr;
    if (jr < 0) jm = g.pery ? g.ny - 1 : 0;
    else if (jr >= g.ny) jm = g.pery ? jr - g.ny : g.ny - 1;
    if (jm >= g.ny) jm = g.ny - 1;
    // column mapping
    const bool vec_ok = (i + 1 < g.nx);  // both points are regular columns -> one 128-bit access
    int c0 = i, c1 = i + 1;
    if (c0 >= g.nx) c0 = g.perx ? min(c0 - g.nx, g.nx - 1) : g.nx - 1;
    if (c1 >= g.nx) c1 = g.perx ? min(c1 - g.nx, g.nx - 1) : g.nx - 1;
    const bool interior_row = (ty >= 1) && (ty <= TY) && (jr < g.ny);
    const bool st0 = interior_row && (i < g.nx);
    const bool st1 = interior_row && (i + 1 < g.nx);
    // x-halo columns are carried by the first and last thread of the row
    const bool edgeL = (tx == 0), edgeR = (tx == TXT - 1);
    int ch = 0;
    if (edgeL) ch = (i0 == 0) ? (g.perx ? g.nx - 1 : 0) : i0 - 1;
    if (edgeR)
    {
        ch = i0 + BX;
        if (ch >= g.nx) ch = g.perx ? min(ch - g.nx, g.nx - 1) : g.nx - 1;
    }
    const bool edge = edgeL || edgeR;
    const long long rowoff = (long long)jm * g.px;

    // thread-constant coefficients
    double dx0 = 0, dx1 = 0, gxa = 0, gxb = 0, gxc = 0, dyj = 0, gya = 0, gyb = 0;
    if (interior_row)
    {
        dyj = g.dy[jr];
        gya = g.gy[jr];
        gyb = g.gy[jr + 1];
        if (i < g.nx)
        {
            dx0 = g.dx[i];
            gxa = g.gx[i];
            gxb = g.gx[i + 1];
        }
        if (i + 1 < g.nx)
        {
            dx1 = g.dx[i + 1];
            gxc = g.gx[i + 2];
        }
    }
    const double axy0 = __dmul_rn(dx0, dyj), axy1 = __dmul_rn(dx1, dyj);
    const bool wrapx0 = g.perx && (i == 0 || i == g.nx - 1);
    const bool wrapx1 = g.perx && (i + 1 == g.nx - 1);  // i+1 == 0 impossible
    const bool wrapy_lo = g.pery && jr == 0, wrapy_hi = g.pery && jr == g.ny - 1;

    // registers of the plane in flight
    double2 rv = make_double2(0, 0), pv = make_double2(0, 0), dv = make_double2(1, 1), xv = make_double2(0, 0);
    double rh = 0, ph = 0, dh = 1;

    auto load_plane = [&](int kk) {
        const long long base = (long long)zstore(g, kk) * g.plane + rowoff;
        if (vec_ok)
        {
            rv = *reinterpret_cast<const double2 *>(v.r + base + i);
            if (!APPLY) pv = *reinterpret_cast<const double2 *>(v.p_in + base + i);
            if (JACOBI) dv = *reinterpret_cast<const double2 *>(v.dinv + base + i);
        }
        else
        {
            rv.x = v.r[base + c0];
            rv.y = v.r[base + c1];
            if (!APPLY)
            {
                pv.x = v.p_in[base + c0];
                pv.y = v.p_in[base + c1];
            }
            if (JACOBI)
            {
                dv.x = v.dinv[base + c0];
                dv.y = v.dinv[base + c1];
            }
        }
        if (edge)
        {
            rh = v.r[base + ch];
            if (!APPLY) ph = v.p_in[base + ch];
            if (JACOBI) dh = v.dinv[base + ch];
        }
        if (!APPLY && xupd && kk >= k0 && kk < k1)
        {
            // x shares the indexing of the other vectors (its ghost planes are never touched)
            const long long xb = (long long)(kk + 1) * g.plane + (long long)jr * g.px + i;
            if (st1) xv = *reinterpret_cast<const double2 *>(v.x + xb);
            else if (st0) xv.x = v.x[xb];
        }
    };
    // does plane kk carry data that can reach a non-zero coefficient?
    auto plane_live = [&](int kk) -> bool {
        if (kk >= 0 && kk < g.nzl) return true;
        // face between the ghost and the interior: minus side -> gz[kz0], plus side -> gz[kz0 + nzl]
        const double f = (kk < 0) ? g.gz[g.kz0] : g.gz[g.kz0 + g.nzl];
        return f != 0.0;
    };

    double2 pm = make_double2(0, 0), pc = make_double2(0, 0), pn;
    double acc[1] = {0.0};
    double czm0 = 0, czm1 = 0;  // minus-face z coefficients of the plane being finished

    bool live = plane_live(k0 - 1);
    if (live) load_plane(k0 - 1);
    for (int kk = k0 - 1; kk <= k1; ++kk)
    {
        // ---- build p on plane kk from the prefetched registers
        double ph_new = 0.0;
        if (live)
        {
            double z0 = JACOBI ? __dmul_rn(rv.x, dv.x) : rv.x;
            double z1 = JACOBI ? __dmul_rn(rv.y, dv.y) : rv.y;
            if (!APPLY)
            {
                z0 = __dadd_rn(z0, shift);
                z1 = __dadd_rn(z1, shift);
            }
            pn.x = APPLY ? z0 : __dadd_rn(z0, __dmul_rn(bcoef, pv.x));
            pn.y = APPLY ? z1 : __dadd_rn(z1, __dmul_rn(bcoef, pv.y));
            if (edge)
            {
                double zh = JACOBI ? __dmul_rn(rh, dh) : rh;
                if (!APPLY) zh = __dadd_rn(zh, shift);
                ph_new = APPLY ? zh : __dadd_rn(zh, __dmul_rn(bcoef, ph));
            }
        }
        else
            pn = make_double2(0.0, 0.0);
        // ---- deferred VecAXPY(X, a', P') on the owned planes
        if (!APPLY && xupd && kk >= k0 && kk < k1)
        {
            double *dst = v.x + (long long)(kk + 1) * g.plane + (long long)jr * g.px + i;
            double2 xn;
            xn.x = __dadd_rn(xv.x, __dmul_rn(aprev, pv.x));
            xn.y = __dadd_rn(xv.y, __dmul_rn(aprev, pv.y));
            if (st1) *reinterpret_cast<double2 *>(dst) = xn;
            else if (st0) dst[0] = xn.x;
        }
        // ---- prefetch the next plane
        const bool live_next = (kk + 1 <= k1) && plane_live(kk + 1);
        if (live_next) load_plane(kk + 1);
        // ---- publish p(kk) to the tile and to global memory
        const int buf = (kk + 3) % 3;
        *reinterpret_cast<double2 *>(&sp[buf][ty][2 + 2 * tx]) = pn;
        if (edgeL) sp[buf][ty][1] = ph_new;
        if (edgeR) sp[buf][ty][BX + 2] = ph_new;
        if (!APPLY)
        {
            const bool owned = (kk >= k0 && kk < k1) || (ghost_store && live && ((kk == -1) || (kk == g.nzl)));
            if (owned)
            {
                double *dst = v.p_out + (long long)(kk + 1) * g.plane + (long long)jr * g.px + i;
                if (st1) *reinterpret_cast<double2 *>(dst) = pn;
                else if (st0) dst[0] = pn.x;
            }
        }
        __syncthreads();
        // ---- w on plane k = kk-1 (needs p on kk-2, kk-1, kk)
        const int k = kk - 1;
        if (k >= k0 && interior_row && (i < g.nx))
        {
            const int pb = (k + 3) % 3;
            const double xm0 = sp[pb][ty][1 + 2 * tx];
            const double xp1 = sp[pb][ty][4 + 2 * tx];
            const double2 ym = *reinterpret_cast<const double2 *>(&sp[pb][ty - 1][2 + 2 * tx]);
            const double2 yp = *reinterpret_cast<const double2 *>(&sp[pb][ty + 1][2 + 2 * tx]);
            const int kg = g.kz0 + k;
            const double dzk = g.dz[kg];
            const double gzb = g.gz[kg + 1];
            const double ayz = __dmul_rn(dyj, dzk);
            const bool wrapz_lo = g.wrapz_lo && k == 0, wrapz_hi = g.wrapz_hi && k == g.nzl - 1;
            const bool slow = wrapy_lo || wrapy_hi || wrapz_lo || wrapz_hi;
            double wout[2];
#pragma unroll
            for (int q = 0; q < 2; ++q)
            {
                const double dxi = q ? dx1 : dx0;
                const double cxm = __dmul_rn(ayz, q ? gxb : gxa);
                const double cxp = __dmul_rn(ayz, q ? gxc : gxb);
                const double axz = __dmul_rn(dxi, dzk);
                const double cym = __dmul_rn(axz, gya), cyp = __dmul_rn(axz, gyb);
                const double czm = q ? czm1 : czm0;
                const double czp = __dmul_rn(q ? axy1 : axy0, gzb);
                const double x0 = q ? pc.y : pc.x;
                const double xm = q ? pc.x : xm0;
                const double xp = q ? xp1 : pc.y;
                const double vym = q ? ym.y : ym.x, vyp = q ? yp.y : yp.x;
                const double vzm = q ? pm.y : pm.x, vzp = q ? pn.y : pn.x;
                const bool wx = q ? wrapx1 : wrapx0;
                double t;
                if (!(slow || wx))
                {
                    // diagonal: MatMatMult accumulation over the D row u(i-1),u(i),v(j-1),v(j),w(k-1),w(k)
                    double dg = __dadd_rn(cxm, cxp);
                    dg = __dadd_rn(dg, cym);
                    dg = __dadd_rn(dg, cyp);
                    dg = __dadd_rn(dg, czm);
                    dg = __dadd_rn(dg, czp);
                    dg = -dg;
                    // MatMult_SeqAIJ in ascending column order: k-1, j-1, i-1, diag, i+1, j+1, k+1
                    t = __dmul_rn(czm, vzm);
                    t = __dadd_rn(t, __dmul_rn(cym, vym));
                    t = __dadd_rn(t, __dmul_rn(cxm, xm));
                    t = __dadd_rn(t, __dmul_rn(dg, x0));
                    t = __dadd_rn(t, __dmul_rn(cxp, xp));
                    t = __dadd_rn(t, __dmul_rn(cyp, vyp));
                    t = __dadd_rn(t, __dmul_rn(czp, vzp));
                }
                else
                {
                    // periodic wrap rows: the wrapped neighbour changes its place in the sorted row
                    const int iq = i + q;
                    const bool xlo = g.perx && iq == 0, xhi = g.perx && iq == g.nx - 1;
                    double dg = __dadd_rn(cxm, cxp);  // first pair: commutative
                    dg = __dadd_rn(dg, wrapy_lo ? cyp : cym);
                    dg = __dadd_rn(dg, wrapy_lo ? cym : cyp);
                    dg = __dadd_rn(dg, wrapz_lo ? czp : czm);
                    dg = __dadd_rn(dg, wrapz_lo ? czm : czp);
                    dg = -dg;
                    const double tzm = __dmul_rn(czm, vzm), tzp = __dmul_rn(czp, vzp);
                    const double tym = __dmul_rn(cym, vym), typ = __dmul_rn(cyp, vyp);
                    const double txm = __dmul_rn(cxm, xm), txp = __dmul_rn(cxp, xp);
                    const double td = __dmul_rn(dg, x0);
                    // sorted columns: zp(wrapped) < zm < yp(w) < ym < xp(w) < xm < d < xp < xm(w) < yp < ym(w) < zp < zm(w)
                    t = 0.0;
                    if (wrapz_hi) t = __dadd_rn(t, tzp);
                    if (!wrapz_lo) t = __dadd_rn(t, tzm);
                    if (wrapy_hi) t = __dadd_rn(t, typ);
                    if (!wrapy_lo) t = __dadd_rn(t, tym);
                    if (xhi) t = __dadd_rn(t, txp);
                    if (!xlo) t = __dadd_rn(t, txm);
                    t = __dadd_rn(t, td);
                    if (!xhi) t = __dadd_rn(t, txp);
                    if (xlo) t = __dadd_rn(t, txm);
                    if (!wrapy_hi) t = __dadd_rn(t, typ);
                    if (wrapy_lo) t = __dadd_rn(t, tym);
                    if (!wrapz_hi) t = __dadd_rn(t, tzp);
                    if (wrapz_lo) t = __dadd_rn(t, tzm);
                }
                wout[q] = t;
                if (q ? st1 : st0) acc[0] = fma(x0, t, acc[0]);
                if (q) czm1 = czp;
                else czm0 = czp;
            }
            double *dst = v.w + (long long)(k + 1) * g.plane + (long long)jr * g.px + i;
            if (st1) *reinterpret_cast<double2 *>(dst) = make_double2(wout[0], wout[1]);
            else dst[0] = wout[0];
        }
        else if (k == k0 - 1 && interior_row && (i < g.nx))
        {
            // entering the chunk: minus-face z coefficient of plane k0
            const double gza = g.gz[g.kz0 + k0];
            czm0 = __dmul_rn(axy0, gza);
            czm1 = __dmul_rn(axy1, gza);
        }
        pm = pc;
        pc = pn;
        live = live_next;
    }
    if (!APPLY) grid_reduce_finalize<1>(acc, FIN_SPMV, ws, cm, st, kc, hist, false);
}

// ------------------------------------------------------------------------------------------
// class 1: fused  r -= a w ; reductions ; convergence test ; (multi-GPU halo push)
// Work item = one double2 of one row; items are grid-strided, UNROLL loads in flight per thread.
// INIT: r already holds b; only the sums (and the halo push) are produced.
// ------------------------------------------------------------------------------------------
struct UpdVecs
{
    double *r;
    const double *w;
    const double *dinv;
    int reverse;  // walk the owned range from the top: the planes k_spmv2 wrote last are still in L2
};

template <bool JACOBI, bool INIT, int UNROLL>
__global__ void __launch_bounds__(256) k_update(GridDev g, UpdVecs v, int fin_kind, ReduceWs ws, CommDev cm,
                                                DevState *st, SolveConsts kc, double *hist)
{
    pdl_sync();
    if (st->done) return;
    const double ma = INIT ? 0.0 : -st->a;
    const double c = st->c;
    const unsigned int ncol2 = (unsigned int)(g.nx + 1) >> 1;
    const unsigned int nrows = (unsigned int)g.ny * (unsigned int)g.nzl;
    const unsigned int nitems = ncol2 * nrows;
    const unsigned int nthreads = gridDim.x * blockDim.x;
    double acc[6] = {0, 0, 0, 0, 0, 0};
    const bool push = (cm.r_ghost_dn != nullptr || cm.r_ghost_up != nullptr);

    for (unsigned int t0 = blockIdx.x * blockDim.x + threadIdx.x; t0 < nitems; t0 += nthreads * UNROLL)
    {
        double2 rr[UNROLL], wr[UNROLL], dr[UNROLL];
        long long off[UNROLL];
        unsigned int rowv[UNROLL];
        bool two[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)
        {
            const unsigned int t = t0 + u * nthreads;
            if (t < nitems)
            {
                const unsigned int row = t / ncol2, col2 = t - row * ncol2;
                const unsigned int kl = row / (unsigned int)g.ny, j = row - kl * (unsigned int)g.ny;
                const int i = 2 * (int)col2;
                rowv[u] = row;
                two[u] = (i + 1 < g.nx);
                off[u] = (long long)(kl + 1) * g.plane + (long long)j * g.px + i;
                if (two[u])
                {
                    rr[u] = *reinterpret_cast<const double2 *>(v.r + off[u]);
                    if (!INIT) wr[u] = *reinterpret_cast<const double2 *>(v.w + off[u]);
                    if (JACOBI) dr[u] = *reinterpret_cast<const double2 *>(v.dinv + off[u]);
                }
                else
                {
                    rr[u] = make_double2(v.r[off[u]], 0.0);
                    if (!INIT) wr[u] = make_double2(v.w[off[u]], 0.0);
                    if (JACOBI) dr[u] = make_double2(v.dinv[off[u]], 0.0);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)
        {
            const unsigned int t = t0 + u * nthreads;
            if (t < nitems)
            {
                double2 rn = rr[u];
                if (!INIT)
                {
                    rn.x = __dadd_rn(rr[u].x, __dmul_rn(ma, wr[u].x));  // VecAXPY(R, -a, W)
                    rn.y = __dadd_rn(rr[u].y, __dmul_rn(ma, wr[u].y));
                    if (two[u]) *reinterpret_cast<double2 *>(v.r + off[u]) = rn;
                    else v.r[off[u]] = rn.x;
                }
                if (push)
                {
                    const unsigned int kl = rowv[u] / (unsigned int)g.ny;
                    const long long o2 = off[u] - (long long)(kl + 1) * g.plane;
                    if (kl == 0 && cm.r_ghost_dn)
                    {
                        if (two[u]) *reinterpret_cast<double2 *>(cm.r_ghost_dn + o2) = rn;
                        else cm.r_ghost_dn[o2] = rn.x;
                    }
                    if (kl == (unsigned int)g.nzl - 1 && cm.r_ghost_up)
                    {
                        if (two[u]) *reinterpret_cast<double2 *>(cm.r_ghost_up + o2) = rn;
                        else cm.r_ghost_up[o2] = rn.x;
                    }
                }
                const double z0 = JACOBI ? __dmul_rn(rn.x, dr[u].x) : rn.x;
                const double d0 = z0 - c;
                acc[0] += z0;
                acc[1] += d0;
                acc[2] = fma(d0, d0, acc[2]);
                acc[3] = fma(d0, rn.x, acc[3]);
                acc[4] += rn.x;
                acc[5] = fma(rn.x, rn.x, acc[5]);
                if (two[u])
                {
                    const double z1 = JACOBI ? __dmul_rn(rn.y, dr[u].y) : rn.y;
                    const double d1 = z1 - c;
                    acc[0] += z1;
                    acc[1] += d1;
                    acc[2] = fma(d1, d1, acc[2]);
                    acc[3] = fma(d1, rn.y, acc[3]);
                    acc[4] += rn.y;
                    acc[5] = fma(rn.y, rn.y, acc[5]);
                }
            }
        }
    }
    grid_reduce_finalize<6>(acc, fin_kind, ws, cm, st, kc, hist, push);
}

// Second-generation class-1 kernel: the owned planes of a solver-layout vector are ONE contiguous range
// (plane = px*ny, no gaps between planes), so the update is a flat 128-bit stream with no index
// arithmetic; only grids whose row pitch is padded (nx not a multiple of 16) mask the pad columns out
// of the shifted sums.  Guard-free main loop (U loads in flight per thread) + scalar tail.
template <bool JACOBI, bool INIT, bool PADDED, bool PUSH>
struct UpdItem
{
    __device__ static __forceinline__ void run(unsigned int i, double2 rr, double2 wr, double2 dr, double ma, double c,
                                               double2 *r2, double2 *gdn, double2 *gup, unsigned int plane2,
                                               unsigned int last0, unsigned int pxh, unsigned int nx, double (&acc)[6])
    {
        double2 rn = rr;
        if (!INIT)
        {
            rn.x = __dadd_rn(rr.x, __dmul_rn(ma, wr.x));  // VecAXPY(R, -a, W)
            rn.y = __dadd_rn(rr.y, __dmul_rn(ma, wr.y));
            r2[i] = rn;
        }
        if (PUSH)
        {
            if (gdn && i < plane2) gdn[i] = rn;
            if (gup && i >= last0) gup[i - last0] = rn;
        }
        bool v0 = true, v1 = true;
        if (PADDED)
        {
            const unsigned int col = 2u * (i % pxh);
            v0 = col < nx;
            v1 = col + 1u < nx;
        }
        const double z0 = JACOBI ? __dmul_rn(rn.x, dr.x) : rn.x;
        const double z1 = JACOBI ? __dmul_rn(rn.y, dr.y) : rn.y;
        const double d0 = v0 ? z0 - c : 0.0;
        const double d1 = v1 ? z1 - c : 0.0;
        acc[0] += z0;
        acc[1] += d0;
        acc[2] = fma(d0, d0, acc[2]);
        acc[3] = fma(d0, rn.x, acc[3]);
        acc[4] += rn.x;
        acc[5] = fma(rn.x, rn.x, acc[5]);
        acc[0] += z1;
        acc[1] += d1;
        acc[2] = fma(d1, d1, acc[2]);
        acc[3] = fma(d1, rn.y, acc[3]);
        acc[4] += rn.y;
        acc[5] = fma(rn.y, rn.y, acc[5]);
    }
};

// With neighbour GPUs (PUSH) the CTAs split into two roles.  The first `npush` CTAs own the two boundary planes: their new
// r values also go into the neighbours' ghost planes (peer stores over NVLink), and only these CTAs pay the system-scope
// fence before their ticket.  All other CTAs stream the interior planes exactly like the single-GPU kernel.  (Letting
// every CTA touch the boundary planes -- first or last -- put slow remote stores and a system fence on the path of all
// of them: 24.6 us against 9.1 us for the same slab on one GPU, profiles/r02_trace_2gpu_slab.log.)
__device__ __forceinline__ unsigned int update_pushers(unsigned int nb2, unsigned int nblocks, int items)
{
    const unsigned int per = 256u * (unsigned int)(items > 0 ? items : 2);  // 2 double2 items per pusher thread (8: +2.6 us, 32: +5 us; profiles/r02_trace_2gpu_slab.log)
    unsigned int np = (nb2 + per - 1u) / per;
    const unsigned int cap = nblocks > 4u ? nblocks / 2u : 1u;
    if (np > cap) np = cap;
    return np < 1u ? 1u : np;
}

template <bool JACOBI, bool INIT, bool PADDED, bool PUSH, int U>
__global__ void __launch_bounds__(256) k_update2(GridDev g, UpdVecs v, int fin_kind, ReduceWs ws, CommDev cm,
                                                 DevState *st, SolveConsts kc, double *hist)
{
    pdl_sync();
    if (st->done) return;
    trace_kernel_start(ws);
    const double ma = INIT ? 0.0 : -st->a;
    const double c = st->c;
    const unsigned long long seq0 = PUSH ? st->seq : 0ull;      // this kernel's reduction will be number seq0 + 1
    const unsigned int plane2 = (unsigned int)(g.plane >> 1);
    const unsigned int n2 = plane2 * (unsigned int)g.nzl;       // host guarantees < 2^32
    const unsigned int last0 = plane2 * (unsigned int)(g.nzl - 1);
    double2 *const __restrict__ r2 = reinterpret_cast<double2 *>(v.r + g.plane);
    const double2 *const __restrict__ w2 = reinterpret_cast<const double2 *>(v.w + g.plane);
    const double2 *const __restrict__ d2 = reinterpret_cast<const double2 *>(v.dinv + (JACOBI ? g.plane : 0));
    double2 *const gdn = reinterpret_cast<double2 *>(cm.r_ghost_dn);
    double2 *const gup = reinterpret_cast<double2 *>(cm.r_ghost_up);
    const unsigned int pxh = (unsigned int)g.px >> 1;
    const unsigned int nx = (unsigned int)g.nx;
    double acc[6] = {0, 0, 0, 0, 0, 0};
    const bool rev = v.reverse != 0;
    // the range of items a CTA role covers: [lo, lo + span), walked by `nb` CTAs of that role.  Single GPU: one role, all
    // items.  PUSH: role 0 = the boundary planes (pusher CTAs), role 1 = the interior planes (all other CTAs; the pushers
    // themselves when the grid has no other CTA).
    bool pusher = false;
    unsigned int npush = 0;
    if (PUSH)
    {
        const unsigned int nb2 = g.nzl >= 2 ? 2u * plane2 : plane2;  // items of the boundary planes
        npush = cm.npush > 0 ? (unsigned int)cm.npush : update_pushers(nb2, gridDim.x, cm.push_items);
        pusher = blockIdx.x < npush;
    }
    const bool no_interior_ctas = PUSH && npush >= gridDim.x;
    for (int role = (PUSH && !pusher) ? 1 : 0; role < ((PUSH && (no_interior_ctas || !pusher)) ? 2 : 1); ++role)
    {
        unsigned int lo = 0, span = n2, nb = gridDim.x, bid = blockIdx.x;
        const bool boundary = PUSH && role == 0;
        if (PUSH)
        {
            if (boundary)
            {
                // boundary planes as one virtual range [0, nb2): the bottom plane, then the top plane (mapped below)
                span = g.nzl >= 2 ? 2u * plane2 : plane2;
                nb = npush;
            }
            else
            {
                lo = plane2;
                span = g.nzl > 2 ? last0 - plane2 : 0u;
                nb = no_interior_ctas ? gridDim.x : gridDim.x - npush;
                bid = no_interior_ctas ? blockIdx.x : blockIdx.x - npush;
            }
        }
        const unsigned int stride = nb * blockDim.x;
        unsigned int i0 = bid * blockDim.x + threadIdx.x;
        const unsigned int n2r = span;  // items of this role
        // main loop: all U items in range (n2r - i0 > (U-1)*stride, written without overflow)
        while (i0 < n2r && n2r - i0 > (unsigned int)(U - 1) * stride)
        {
            double2 rr[U], wr[U], dr[U];
            unsigned int idx[U];
#pragma unroll
            for (int u = 0; u < U; ++u)
            {
                const unsigned int j = i0 + u * stride;
                idx[u] = lo + (rev ? span - 1u - j : j);
                if (boundary && idx[u] >= plane2) idx[u] = last0 + (idx[u] - plane2);  // second half: the top plane
                rr[u] = r2[idx[u]];
                if (!INIT) wr[u] = w2[idx[u]];
                if (JACOBI) dr[u] = d2[idx[u]];
            }
#pragma unroll
            for (int u = 0; u < U; ++u)
                UpdItem<JACOBI, INIT, PADDED, PUSH>::run(idx[u], rr[u], wr[u], dr[u], ma, c, r2, gdn, gup, plane2, last0, pxh, nx,
                                                         acc);
            if (n2r - i0 <= (unsigned int)U * stride)
            {
                i0 = n2r;
                break;
            }
            i0 += U * stride;
        }
        for (; i0 < n2r; i0 = (n2r - i0 > stride) ? i0 + stride : n2r)
        {
            unsigned int i = lo + (rev ? span - 1u - i0 : i0);
            if (boundary && i >= plane2) i = last0 + (i - plane2);  // second half of the boundary range: the top plane
            double2 rr = r2[i], wr = make_double2(0, 0), dr = make_double2(0, 0);
            if (!INIT) wr = w2[i];
            if (JACOBI) dr = d2[i];
            UpdItem<JACOBI, INIT, PADDED, PUSH>::run(i, rr, wr, dr, ma, c, r2, gdn, gup, plane2, last0, pxh, nx, acc);
        }
    }
    // with the halo hand-shake the pushers' system fence comes AFTER their ticket: it no longer delays the reduction
    const bool flagged = PUSH && cm.mode == 1 && cm.halo_flag_local != nullptr;
    grid_reduce_finalize<6>(acc, fin_kind, ws, cm, st, kc, hist, PUSH && pusher && !flagged && !(cm.dbg_flags & 1));
    if (flagged && pusher)
    {
        __threadfence_system();  // this thread's peer stores are performed
        __syncthreads();
        if (threadIdx.x == 0)
        {
            const unsigned int t = atomicAdd(cm.push_counter, 1u);
            if (t == npush - 1u)
            {
                *cm.push_counter = 0u;  // the next launch cannot start before this grid has completed
                __threadfence_system();
                if (cm.halo_flag_dn_peer) st_release_sys(cm.halo_flag_dn_peer, seq0 + 1ull);
                if (cm.halo_flag_up_peer) st_release_sys(cm.halo_flag_up_peer, seq0 + 1ull);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// tail: apply the x update still owed when the loop ended (cg.c updates x before every test)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_xtail(GridDev g, double *x, const double *p0, const double *p1, DevState *st)
{
    if (!st->pending) return;
    const double a = st->a;
    const double *p = st->pcur ? p1 : p0;
    const long long nrows = (long long)g.ny * g.nzl;
    for (long long row = blockIdx.x; row < nrows; row += gridDim.x)
    {
        const long long kl = row / g.ny, j = row - kl * g.ny;
        const long long base = (kl + 1) * g.plane + j * g.px;
        for (int i = threadIdx.x; i < g.nx; i += blockDim.x)
            x[base + i] = __dadd_rn(x[base + i], __dmul_rn(a, p[base + i]));
    }
}

// ------------------------------------------------------------------------------------------
// layout helpers: compact (i fastest, no padding, no ghosts) <-> solver layout
// ------------------------------------------------------------------------------------------
// dst(solver layout, owned planes) <- src(compact); optionally zero a second solver-layout vector
__global__ void __launch_bounds__(256) k_scatter(GridDev g, const double *src, double *dst, double *zero_me)
{
    const long long nrows = (long long)g.ny * g.nzl;
    for (long long row = blockIdx.x; row < nrows; row += gridDim.x)
    {
        const long long kl = row / g.ny, j = row - kl * g.ny;
        const long long base = (kl + 1) * g.plane + j * g.px;
        for (int i = threadIdx.x; i < g.px; i += blockDim.x)
        {
            if (src) dst[base + i] = (i < g.nx) ? src[row * g.nx + i] : 0.0;
            if (zero_me) zero_me[base + i] = 0.0;
        }
    }
}
__global__ void __launch_bounds__(256) k_gather(GridDev g, const double *src, double *dst)
{
    const long long nrows = (long long)g.ny * g.nzl;
    for (long long row = blockIdx.x; row < nrows; row += gridDim.x)
    {
        const long long kl = row / g.ny, j = row - kl * g.ny;
        const long long base = (kl + 1) * g.plane + j * g.px;
        for (int i = threadIdx.x; i < g.nx; i += blockDim.x) dst[row * g.nx + i] = src[base + i];
    }
}

// Jacobi set-up: dinv = 1/diag on every storage plane (ghost planes included), diag accumulated
// exactly as in k_spmv (PCSetUp_Jacobi: MatGetDiagonal + VecReciprocal; zero diagonal -> 1).
__global__ void k_jacobi_setup(GridDev g, double *dinv)
{
    const long long total = (long long)(g.nzl + 2) * g.ny * g.nx;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x)
    {
        const int i = (int)(t % g.nx);
        const long long q = t / g.nx;
        const int j = (int)(q % g.ny);
        const int ks = (int)(q / g.ny);  // storage plane
        int kg = g.kz0 + ks - 1;         // global plane
        double val = 1.0;
        if (kg < 0) kg += g.nzg;
        if (kg >= g.nzg) kg -= g.nzg;
        {
            const double dyj = g.dy[j], dzk = g.dz[kg], dxi = g.dx[i];
            const double ayz = __dmul_rn(dyj, dzk), axz = __dmul_rn(dxi, dzk), axy = __dmul_rn(dxi, dyj);
            const double cxm = __dmul_rn(ayz, g.gx[i]), cxp = __dmul_rn(ayz, g.gx[i + 1]);
            const double cym = __dmul_rn(axz, g.gy[j]), cyp = __dmul_rn(axz, g.gy[j + 1]);
            const double czm = __dmul_rn(axy, g.gz[kg]), czp = __dmul_rn(axy, g.gz[kg + 1]);
            const bool wy = g.pery && j == 0;
            const bool wz = (g.gz[0] != 0.0) && kg == 0;  // periodic z: wrap face coefficient is non-zero
            double dg = __dadd_rn(cxm, cxp);
            dg = __dadd_rn(dg, wy ? cyp : cym);
            dg = __dadd_rn(dg, wy ? cym : cyp);
            dg = __dadd_rn(dg, wz ? czp : czm);
            dg = __dadd_rn(dg, wz ? czm : czp);
            dg = -dg;
            val = (dg != 0.0) ? __ddiv_rn(1.0, dg) : 1.0;
        }
        dinv[(long long)ks * g.plane + (long long)j * g.px + i] = val;
    }
}

// push the boundary planes of a vector into the neighbours' ghost planes (b200ls_apply, multi-GPU)
__global__ void k_push_halo(GridDev g, const double *vsrc, double *ghost_dn, double *ghost_up)
{
    const long long n = (long long)g.ny * g.px;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x)
    {
        if (ghost_dn) ghost_dn[t] = vsrc[1 * g.plane + t];
        if (ghost_up) ghost_up[t] = vsrc[(long long)g.nzl * g.plane + t];
    }
}

// cross-GPU barrier through the mailboxes (a reduction whose result is dropped)
__global__ void k_barrier(CommDev cm, DevState *st)
{
    __shared__ unsigned int s_ll[B200_MAX_RANKS * B200_LLW];
    if (blockIdx.x == 0 && threadIdx.x < 32)
    {
        double S[B200_NSUM];
#pragma unroll
        for (int q = 0; q < B200_NSUM; ++q) S[q] = 0.0;
        __threadfence_system();
        const unsigned long long seq = st->seq + 1;
        const bool ok = mailbox_allreduce(S, cm, seq, (int)threadIdx.x, s_ll);
        if (threadIdx.x == 0)
        {
            st->seq = seq;
            if (!ok)
            {
                st->err = 1;
                st->done = 1;
            }
        }
    }
}

// L2 flush helper for kernel timing
__global__ void k_fill(double *buf, long long n, double val)
{
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x)
        buf[t] = val;
}

}  // namespace b200
