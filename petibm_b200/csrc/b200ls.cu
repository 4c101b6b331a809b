// b200ls.cu -- host side of libb200ls.so: the C ABI declared in include/b200ls.h.
//
// Mirrors what petibm::linsolver::LinSolverKSP does with a KSP object
// (src/linsolver/linsolverksp.cpp:16-132 of barbagroup/PetIBM): create -> options -> operator ->
// solve (zero initial guess) -> iterations / residual / reason; all arithmetic runs in the sm_100a
// kernels of kernels.cuh, all scalars of the Krylov recurrence stay on the device, the host only
// enqueues work and polls a "done" flag.  There is no CPU solve path in this file.
#include "../../include/b200ls.h"

#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <map>
#include <string>
#include <vector>

#include "kernels.cuh"
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

// ------------------------------------------------------------------------------------------
// NCCL through dlopen: the library is already in the process when the host is PyTorch or an MPI
// application linked with NCCL; we never link against a second copy.
// ------------------------------------------------------------------------------------------
namespace {
typedef struct ncclComm *ncclComm_t;
struct NcclUniqueId { char internal[128]; };
struct NcclApi
{
    void *lib = nullptr;
    int (*GetUniqueId)(NcclUniqueId *) = nullptr;
    int (*CommInitRank)(ncclComm_t *, int, NcclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    bool ok = false;
};
NcclApi &nccl_api()
{
    static NcclApi api;
    static bool tried = false;
    if (tried) return api;
    tried = true;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    // prefer a copy that is already loaded (torch's bundled one)
    for (const char *n : names)
    {
        api.lib = dlopen(n, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
        if (api.lib) break;
    }
    if (!api.lib)
        for (const char *n : names)
        {
            api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (api.lib) break;
        }
    if (!api.lib) return api;
    api.GetUniqueId = (int (*)(NcclUniqueId *))dlsym(api.lib, "ncclGetUniqueId");
    api.CommInitRank = (int (*)(ncclComm_t *, int, NcclUniqueId, int))dlsym(api.lib, "ncclCommInitRank");
    api.CommDestroy = (int (*)(ncclComm_t))dlsym(api.lib, "ncclCommDestroy");
    api.AllReduce = (int (*)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t))dlsym(api.lib, "ncclAllReduce");
    api.GetErrorString = (const char *(*)(int))dlsym(api.lib, "ncclGetErrorString");
    api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllReduce;
    return api;
}
constexpr int kNcclDouble = 8;  // ncclFloat64
constexpr int kNcclSum = 0;     // ncclSum
}  // namespace

// ------------------------------------------------------------------------------------------
// solver object
// ------------------------------------------------------------------------------------------
enum OperatorKind { OP_NONE = 0, OP_STENCIL = 1, OP_CSR = 2, OP_SEP = 3 };

struct b200ls_solver
{
    int device = 0;
    cudaStream_t stream = nullptr;
    b200ls_options opt;
    std::string err;

    // ---- operator: separable stencil
    int op = OP_NONE;
    int dim = 3;
    int64_t n[3] = {0, 0, 0};
    int per[3] = {0, 0, 0};
    int64_t slab_lo = 0, slab_hi = 0;
    double dt = 0.0;
    std::vector<double> hdx, hdy, hdz, hgx, hgy, hgz;  // host copies (verification)
    double *d_axes = nullptr;                          // one allocation for the six 1-D arrays
    GridDev g{};
    int64_t nlocal = 0;      // unknowns owned by this rank
    int64_t nglobal = 0;
    size_t vec_elems = 0;    // doubles per solver-layout vector (ghost planes included)

    // ---- operator: general CSR (single GPU)
    CsrDev csr{};
    int64_t *d_rowptr = nullptr;
    int32_t *d_col = nullptr;
    double *d_val = nullptr;

    // ---- operator: line-coefficient form of an assembled staggered-grid matrix (single GPU, sep_kernels.cuh)
    SepDev sep{};
    bool sep_hybrid = false;   // the stencil block is the pressure operator of the mesh (b200ls_set_poisson_hybrid)
    double *d_sep_coef = nullptr, *d_sep_diag = nullptr;
    int64_t *d_rem_rowptr = nullptr;
    int32_t *d_rem_col = nullptr;
    double *d_rem_val = nullptr;

    // ---- geometric multigrid preconditioner of the separable operator (mg_kernels.cuh, mg_schedule.h)
    std::vector<MgHostLevel> mg_host;
    std::vector<MgLevel> mg_dev;
    std::vector<std::vector<double *>> mg_bufs;  // per level: 4 work vectors (+ the right-hand side of a coarse level)
    std::vector<double *> mg_axes;               // per coarse level: its six 1-D arrays
    std::vector<int *> mg_maps;                  // per level but the coarsest: fine -> coarse index maps
    MgParams mg_prm;
    bool mg_ready = false;
    int mg_built_levels = 0;
    int64_t graph_launches = 0;   // launches inside the graph replayed by graph_batch (csr_solver.inc)
    int sep_tile = -1;            // tuning "sep_tile": tiled plane-marching kernels of sep_tile.cuh for the line-coefficient operator (2: on, tiles of 64 cells in x; 0: row-per-thread kernels; -1: the rule of sep_tile_xr)
    int sep_stages = 3;           // tuning "sep_stages": planes in the per-thread cp.async queue of those kernels (3 or 4)
    int sep_zchunk = 0;           // tuning "sep_zchunk": planes per z chunk of those kernels (0: enough chunks for 8 CTAs per SM)
    int csr_graph = 0;            // tuning "csr_graph": CG / BiCGStab batches of the assembled-operator paths as CUDA graphs
    int mg_fuse = 1;         // tuning "mg_fuse": r -= a w and the six sums ride on the first / last fine-level step of the cycle (on: 15.1 -> 14.2 ms at 256^3, profiles/r02_tts_multigrid.log)
    int mg_tail = 1;         // tuning "mg_tail": the coarse levels of the cycle as one launch (k_mg_tail; on: 597 -> 267 launches, 15.7 -> 15.1 ms at 256^3)
    MgOp *mg_tail_ops = nullptr;
    MgLevel *mg_tail_levels = nullptr;
    int mg_tail_nops = 0;
    int mg_tail_degrees[2] = {0, 0};
    double *mg_tail_result = nullptr;
    int mg_graph = 1;        // tuning "mg_graph": replay pairs of preconditioned iterations as one CUDA graph (on: 14.2 -> 13.8 ms at 256^3; 2-D 448^2: 4.05 -> 2.26 ms)

    // ---- direct solve of a small assembled system (preonly + lu; dense_kernels.cuh)
    double *d_dense = nullptr;   // n x n factors
    int *d_dense_info = nullptr;
    double *d_dense_work = nullptr;          // substitution: work vector | block flags | ticket
    unsigned long long dense_epoch = 0;
    bool dense_ready = false;

    // ---- vectors (solver layout)
    double *arena = nullptr;  // [mailboxes | flags | r]; exported over CUDA IPC
    size_t arena_bytes = 0;
    double *r = nullptr, *p[2] = {nullptr, nullptr}, *w = nullptr, *x = nullptr, *dinv = nullptr;
    double *bcg[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // BiCGStab work vectors
    double *stage = nullptr;  // compact staging buffer for the host-pointer API
    int has_const = 0;
    int nnullvecs = 0;
    double *d_nullvecs = nullptr;

    // ---- device state
    DevState *d_state = nullptr;
    DevState *h_state = nullptr;  // pinned, 4 slots
    cudaEvent_t ev_slot[4] = {nullptr, nullptr, nullptr, nullptr};
    ReduceWs ws{};
    int max_blocks = 0;
    double *d_hist = nullptr;
    int hist_cap = 0;
    double *d_sendrecv = nullptr;  // NCCL mode: [0..7] send, [8..15] recv

    // ---- comm
    int rank = 0, nranks = 1;
    int reduce_mode = B200LS_REDUCE_P2P, halo_mode = B200LS_HALO_STORE;
    bool connected = false;
    std::vector<void *> peer_base;  // arena base of every rank as mapped here
    CommDev cm{};
    ncclComm_t nccl = nullptr;
    double *ghost_dn = nullptr, *ghost_up = nullptr;  // neighbours' ghost planes of r (mapped)

    // ---- results
    int its = 0, reason = 0;
    double rnorm = 0.0;
    std::vector<double> history;

    // ---- launch configuration / tuning
    int num_sms = 148;
    int kz_chunk = 0;        // 0 = auto
    int upd_blocks = 0;      // 0 = auto
    std::map<std::pair<const void *, int>, TmaMap> tma_maps;  // (vector, box kind) -> tensor map of k_spmv4
    int tile = -1;           // K1 tile variant (-1: cost model picks 10 or 18; 10+: k_spmv2; <10: k_spmv)
    int upd_variant = 0;     // 0: flat k_update2, 1: first-generation k_update, 2: k_update2f (Jacobi diagonal on the fly)
    int upd_reverse = 1;     // k_update2 walks the owned range top-down (L2 reuse of what k_spmv2 wrote last)
    int use_graph = 1;
    int use_pdl = 1;
    bool in_loop = false;    // launches issued from the CG loop may overlap their predecessor (PDL)
    cudaGraphExec_t graph_exec = nullptr;
    int graph_iters = 0;
    unsigned long long graph_key = 0;

    // ---- measurement
    cudaEvent_t ev_a = nullptr, ev_b = nullptr, ev_c = nullptr, ev_d = nullptr;
    double solve_ms = 0.0, loop_ms = 0.0, e2e_ms = 0.0;
    cudaEvent_t ev_e0 = nullptr, ev_e1 = nullptr;
    int64_t launches = 0;
    int profile = 0;
    std::vector<cudaEvent_t> prof_ev;  // pairs
    std::vector<int> prof_cls;
    double prof_ms[4] = {0, 0, 0, 0};
    int64_t prof_cnt[4] = {0, 0, 0, 0};
    double *flush_buf = nullptr;
    size_t flush_elems = 0;
};

namespace {

int fail(b200ls_solver *h, int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (h) h->err = buf;
    return code;
}

#define CU(h, call)                                                                                  \
    do                                                                                               \
    {                                                                                                \
        cudaError_t e__ = (call);                                                                    \
        if (e__ != cudaSuccess)                                                                      \
            return fail(h, B200LS_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call,             \
                        cudaGetErrorString(e__));                                                    \
    } while (0)

#define TRY(expr)                 \
    do                            \
    {                             \
        int rc__ = (expr);        \
        if (rc__ != B200LS_OK) return rc__; \
    } while (0)

inline int64_t round_up(int64_t a, int64_t b) { return (a + b - 1) / b * b; }

void free_mg(b200ls_solver *h);

void free_vectors(b200ls_solver *h)
{
    free_mg(h);
    auto fr = [](auto *&p) {
        if (p) cudaFree(p);
        p = nullptr;
    };
    if (h->connected)
    {
        for (int q = 0; q < (int)h->peer_base.size(); ++q)
            if (q != h->rank && h->peer_base[q]) cudaIpcCloseMemHandle(h->peer_base[q]);
        h->peer_base.clear();
        h->connected = false;
    }
    fr(h->arena);
    h->r = nullptr;
    fr(h->p[0]);
    fr(h->p[1]);
    fr(h->w);
    fr(h->x);
    fr(h->dinv);
    for (auto &b : h->bcg) fr(b);
    fr(h->stage);
    fr(h->d_axes);
    fr(h->d_rowptr);
    fr(h->d_col);
    fr(h->d_val);
    fr(h->d_nullvecs);
    h->sep_hybrid = false;
    h->tma_maps.clear();
    fr(h->d_dense);
    fr(h->d_dense_info);
    fr(h->d_dense_work);
    h->dense_ready = false;
    fr(h->d_sep_coef);
    fr(h->d_sep_diag);
    fr(h->d_rem_rowptr);
    fr(h->d_rem_col);
    fr(h->d_rem_val);
    if (h->graph_exec)
    {
        cudaGraphExecDestroy(h->graph_exec);
        h->graph_exec = nullptr;
    }
    h->op = OP_NONE;
}

SolveConsts make_consts(const b200ls_solver *h)
{
    SolveConsts k;
    k.rtol = h->opt.rtol;
    k.atol = h->opt.atol;
    k.divtol = h->opt.divtol;
    k.nglobal = (double)h->nglobal;
    k.max_it = h->opt.max_it;
    k.norm_type = h->opt.norm_type;
    k.has_const = h->has_const;
    k.hist_cap = h->hist_cap;
    return k;
}

void invalidate_graph(b200ls_solver *h);

int ensure_hist(b200ls_solver *h)
{
    const int want = std::min(std::max(h->opt.max_it, 0) + 2, 1 << 22);
    if (want > h->hist_cap)
    {
        if (h->d_hist) cudaFree(h->d_hist);
        h->d_hist = nullptr;
        CU(h, cudaMalloc(&h->d_hist, sizeof(double) * (size_t)want));
        h->hist_cap = want;
        invalidate_graph(h);
    }
    return B200LS_OK;
}

// ---- profiling helpers: one event pair per launch, resolved after the solve
struct ProfScope
{
    b200ls_solver *h;
    int cls;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    ProfScope(b200ls_solver *h_, int cls_) : h(h_), cls(cls_)
    {
        if (h->profile)
        {
            cudaEventCreate(&e0);
            cudaEventCreate(&e1);
            cudaEventRecord(e0, h->stream);
        }
    }
    ~ProfScope()
    {
        if (h->profile)
        {
            cudaEventRecord(e1, h->stream);
            h->prof_ev.push_back(e0);
            h->prof_ev.push_back(e1);
            h->prof_cls.push_back(cls);
        }
    }
};

void resolve_profile(b200ls_solver *h)
{
    for (size_t q = 0; q < h->prof_cls.size(); ++q)
    {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, h->prof_ev[2 * q], h->prof_ev[2 * q + 1]) == cudaSuccess)
        {
            h->prof_ms[h->prof_cls[q]] += ms;
            h->prof_cnt[h->prof_cls[q]] += 1;
        }
        cudaEventDestroy(h->prof_ev[2 * q]);
        cudaEventDestroy(h->prof_ev[2 * q + 1]);
    }
    h->prof_ev.clear();
    h->prof_cls.clear();
}

// ------------------------------------------------------------------------------------------
// launch wrappers
// ------------------------------------------------------------------------------------------
// Kernel launch with (optionally) the programmatic-dependent-launch attribute.
template <typename... KArgs, typename... Args>
cudaError_t launch_k(b200ls_solver *h, bool pdl, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, Args... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = h->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

inline bool pdl_on(const b200ls_solver *h)
{
    return h->use_pdl && !h->profile && !(h->nranks > 1 && h->reduce_mode == B200LS_REDUCE_NCCL);
}

struct TileCfg { int txt, tyt; };
// tile < 10: first-generation register-prefetch kernel (k_spmv); tile >= 10: cp.async ring kernel (k_spmv2)
inline TileCfg tile_dims(int tile)
{
    switch (tile)
    {
        case 0: return {32, 8};    // first-generation kernel (kept for the profile comparison)
        case 13: return {32, 6};   // 64 x 4 tile, S = 4, 4 CTAs/SM
        case 15: return {32, 8};   // 64 x 6 tile, S = 3, 3 CTAs/SM
        case 18: return {32, 12};  // 64 x 10 tile, S = 3, 2 CTAs/SM
        case 30: case 32: return {32, 12};  // k_spmv3 (balanced split; 32: + split barrier), 64 x 10 tile -- round-2 candidates
        case 31: case 33: return {32, 8};   // k_spmv3 (balanced split; 33: + split barrier), 64 x 6 tile  -- round-2 candidates
        case 40: return {32, 9};   // k_spmv4 (TMA): 64 x 8 tile + producer warp, S = 4, 2 CTAs/SM
        case 41: case 44: case 45: case 46: case 47: return {32, 5};   // k_spmv4: 64 x 4 tile, S = 4 (44: S = 6, 45: S = 3), 4 CTAs/SM; 46 / 47: balanced split (S = 4 / 6)
        case 42: return {32, 9};   // k_spmv4: 64 x 8 tile, S = 3, 3 CTAs/SM
        case 43: return {32, 17};  // k_spmv4: 64 x 16 tile, S = 3, 1 CTA/SM
        case 48: case 49: return {32, 3};   // k_spmv4: 64 x 2 tile, S = 4 / 6, 6 CTAs/SM
        case 60: case 61: return {32, 7};   // k_spmv4: 64 x 6 tile, S = 4 / 3, 3 CTAs/SM
        case 62: return {32, 6};            // k_spmv4: 64 x 5 tile, S = 4, 3 CTAs/SM
        case 50: case 51: case 53: case 54: case 55: case 56: return {32, 8};  // k_spmv5 (no staging, cache-resident slabs): 64 x 8 rows, 4 / 8 / 16 planes per thread
        case 52: return {32, 4};                    // k_spmv5: 64 x 4 rows, 2 planes per thread
        default: return {32, 8};   // 10: 64 x 6 tile, S = 4, 3 CTAs/SM
    }
}
inline bool tile_is_tma(int tile) { return (tile >= 40 && tile < 50) || (tile >= 60 && tile < 70); }
inline bool tile_is_direct(int tile) { return tile >= 50 && tile < 60; }
inline int tile_direct_planes(int tile) { return (tile == 50 || tile == 54 || tile == 56) ? 4 : (tile == 51 || tile == 55) ? 8 : tile == 52 ? 2 : 16; }
// rows of a tile that produce results
inline int tile_rows(int tile)
{
    const TileCfg t = tile_dims(tile);
    return tile_is_direct(tile) ? t.tyt : tile_is_tma(tile) ? t.tyt - 1 : t.tyt - 2;
}
inline int tile_ctas_per_sm(int tile)
{
    switch (tile)
    {
        case 13: case 41: case 44: case 45: case 46: case 47: return 4;
        case 18: case 30: case 32: case 40: return 2;
        case 0: return 2;
        case 43: return 1;
        case 48: case 49: return 6;
        case 60: case 61: case 62: return 3;
        default: return 3;
    }
}

// Launch shape of the fused SpMV kernel: tile variant and planes per CTA.  Deterministic cost model fitted to
// the sweeps in profiles/ (scripts/sweep_k1.py): a CTA marches kz planes plus a ~5-plane pipeline fill,
// CTAs fill SMs in waves of (SMs x resident CTAs), and a tile with TY of TYT thread rows doing stencil work
// has efficiency TY/TYT.
struct K1Cfg { int tile, kz; };
inline K1Cfg k1_config(const b200ls_solver *h)
{
    const int nzl = h->g.nzl;
    int tiles[2] = {10, 18};
    int ntiles = 2;
    // the TMA kernel covers non-periodic grids (periodic wrap rows re-sort their columns: k_spmv2<PER>)
    const bool tma_refused = (tile_is_tma(h->tile) || tile_is_direct(h->tile)) && (h->per[0] || h->per[1] || h->per[2]);
    if (tile_is_direct(h->tile) && !tma_refused) return {h->tile, std::min(tile_direct_planes(h->tile), std::max(1, nzl))};
    if (h->tile >= 0 && !tma_refused)
    {
        tiles[0] = h->tile;
        ntiles = 1;
    }
    else if (h->tile < 0 && !(h->per[0] || h->per[1] || h->per[2]) && h->opt.pc_type != B200LS_PC_JACOBI)
    {
        // (Jacobi keeps the cp.async kernel: with the extra 1/diag halo box and z rebuilt at six neighbours the TMA kernel
        // is 17 % slower there, 3 155 against 3 785 iterations/s at 256^3, gpurun call r02o.)
        // Non-periodic grids: the TMA kernel (k_spmv4, 64 x 4 tiles, 4 CTAs per SM).  When the five solver vectors fit the
        // 126 MB L2 (8-GPU slabs of 256^3, 128^3 on one GPU) it wins clearly (18.4 us against 24.1 us per launch on a
        // 256 x 256 x 32 slab, profiles/r02_trace_2gpu_slab.log) and the wave model below picks its z chunks.  Out of HBM it
        // is level with the cp.async kernel kernel-by-kernel but faster inside the solve with chunks of about 43 planes
        // (4 674 against 4 530 iterations/s at 256^3 on the same box, gpurun call r02n; 26 - 64 planes are within 2 %,
        // 128 planes lose 10 %): nch = round(nzl / 43).
        tiles[0] = 41;
        ntiles = 1;
        if (40.0 * (double)(nzl + 2) * (double)h->g.plane > 110.0e6)
        {
            // Out of HBM the kernel runs at (resident CTAs) x (one plane per ~1.2 us each): it needs z chunks of 40 .. 72
            // planes (shorter: redundant halo planes and prologues) AND at least 1.5 waves of (SMs x 4) CTAs -- 256^3 with
            // 43 or 64 planes per CTA: 4 660 - 4 680 iterations/s; 37 / 52 / 86 planes (last wave 3 % / 16 % / 30 % full):
            // 4 470 / 4 390 / 4 290; 128 planes = ONE wave of 512 CTAs: -10 %; a 256 x 256 x 64 slab (4 GPUs) as one wave of
            // 256 CTAs: 11.2 k instead of 13.9 k iterations/s (gpurun calls r02p, r02s).  Grids too small for that keep the
            // cp.async kernel and its tuned launch shape.
            const int64_t xy = (int64_t)((h->g.nx + 63) / 64) * ((h->g.ny + 3) / 4);
            const int64_t slots = (int64_t)h->num_sms * 4;
            int best_kz = 0;
            double best_score = -1.0;
            for (int nch = 1; nch <= std::max(1, nzl / 40) && h->kz_chunk <= 0; ++nch)
            {
                const int kz = (nzl + nch - 1) / nch;
                if (kz > 72) continue;
                const int64_t blocks = xy * ((nzl + kz - 1) / kz);
                if (2 * blocks < 3 * slots) continue;
                const int64_t waves = (blocks + slots - 1) / slots;
                const double score = (double)blocks / (double)(waves * slots) * ((double)kz / (double)(kz + 3));
                if (score >= best_score * (1.0 - 1e-12))   // ties: fewer, longer chunks
                {
                    if (score > best_score * (1.0 + 1e-12) || kz > best_kz) best_kz = kz;
                    best_score = std::max(score, best_score);
                }
            }
            if (best_kz > 0) return {41, best_kz};
            if (h->kz_chunk <= 0)
            {
                tiles[0] = 10;   // back to the cp.async candidates
                tiles[1] = 18;
                ntiles = 2;
            }
        }
    }
    K1Cfg best{tiles[0], std::max(1, std::min(nzl, 512))};
    double best_score = -1.0;
    for (int q = 0; q < ntiles; ++q)
    {
        const TileCfg t = tile_dims(tiles[q]);
        const int ty = tile_rows(tiles[q]);
        const int64_t xy = (int64_t)((h->g.nx + 2 * t.txt - 1) / (2 * t.txt)) * ((h->g.ny + ty - 1) / ty);
        const int64_t slots = (int64_t)h->num_sms * tile_ctas_per_sm(tiles[q]);
        const int max_nch = (h->kz_chunk > 0) ? 1 : std::max(1, std::min(64, nzl / 4));
        for (int nch = 1; nch <= max_nch; ++nch)
        {
            int kz = (h->kz_chunk > 0) ? std::min(h->kz_chunk, nzl) : (nzl + nch - 1) / nch;
            if (kz > 512) kz = 512;
            const int64_t chunks = (nzl + kz - 1) / kz;
            const int64_t blocks = xy * chunks;
            const int64_t waves = (blocks + slots - 1) / slots;
            const double fill = (double)blocks / (double)(waves * slots);
            const double score = fill * ((double)kz / (double)(kz + 5)) * ((double)ty / (double)t.tyt);
            if (score > best_score * (1.0 + 1e-12))
            {
                best_score = score;
                best = {tiles[q], kz};
            }
        }
    }
    return best;
}
inline TileCfg tile_cfg(const b200ls_solver *h) { return tile_dims(k1_config(h).tile); }
inline int auto_kz_chunk(const b200ls_solver *h) { return k1_config(h).kz; }

inline bool grid_periodic(const b200ls_solver *h) { return h->per[0] || h->per[1] || h->per[2]; }

template <int TYT, int S, int MINB, bool JAC, bool APPLY>
int launch_spmv2_cfg(b200ls_solver *h, const VecSet &v, int ghost_store, dim3 grid, int kz)
{
    using L = Spmv2Smem<32, TYT, S, JAC, APPLY>;
    const SolveConsts kc = make_consts(h);
    dim3 block(32, TYT);
    if (grid_periodic(h))
    {
        auto kern = k_spmv2<32, TYT, S, MINB, JAC, APPLY, true>;
        static bool attr_done = false;
        if (!attr_done)
        {
            CU(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::total(512)));
            attr_done = true;
        }
        CU(h, launch_k(h, h->in_loop && pdl_on(h), kern, grid, block, L::total(kz), h->g, v, kz, h->ws, h->cm, h->d_state, kc,
                       h->d_hist, ghost_store));
    }
    else
    {
        auto kern = k_spmv2<32, TYT, S, MINB, JAC, APPLY, false>;
        static bool attr_done = false;
        if (!attr_done)
        {
            CU(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::total(512)));
            attr_done = true;
        }
        CU(h, launch_k(h, h->in_loop && pdl_on(h), kern, grid, block, L::total(kz), h->g, v, kz, h->ws, h->cm, h->d_state, kc,
                       h->d_hist, ghost_store));
    }
    return B200LS_OK;
}

// k_spmv3: one resident wave of CTAs, equal ranges of the linearised (tile, plane) space
struct K3Cfg { int nctas, planes_per_cta, max_seg; };
inline K3Cfg k3_config(const b200ls_solver *h, int tyt, int ctas_per_sm)
{
    const int ty = tyt - 2;
    const int64_t ntiles = (int64_t)((h->g.nx + 63) / 64) * ((h->g.ny + ty - 1) / ty);
    const int64_t total = ntiles * h->g.nzl;
    int64_t nctas = std::min<int64_t>((int64_t)h->num_sms * ctas_per_sm, std::max<int64_t>(1, total / 4));
    int64_t ppc = (total + nctas - 1) / nctas;
    ppc = std::min<int64_t>(ppc, 512);  // the coefficient table of a segment holds 512 planes
    nctas = (total + ppc - 1) / ppc;
    return {(int)nctas, (int)ppc, (int)std::min<int64_t>(ppc, h->g.nzl)};
}

template <int TYT, int S, int MINB, bool JAC, bool SPLIT>
int launch_spmv3_cfg(b200ls_solver *h, const VecSet &v, int ghost_store)
{
    using L = Spmv3Smem<32, TYT, S, JAC, false, SPLIT ? 4 : 3>;
    const K3Cfg c = k3_config(h, TYT, MINB);
    const SolveConsts kc = make_consts(h);
    dim3 grid((unsigned)c.nctas), block(32, TYT);
    const size_t smem = L::total(c.max_seg);
    if (grid_periodic(h))
    {
        auto kern = k_spmv3<32, TYT, S, MINB, JAC, false, true, SPLIT>;
        static bool attr_done = false;
        if (!attr_done)
        {
            CU(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::total(512)));
            attr_done = true;
        }
        CU(h, launch_k(h, h->in_loop && pdl_on(h), kern, grid, block, smem, h->g, v, c.planes_per_cta, c.max_seg, h->ws, h->cm,
                       h->d_state, kc, h->d_hist, ghost_store));
    }
    else
    {
        auto kern = k_spmv3<32, TYT, S, MINB, JAC, false, false, SPLIT>;
        static bool attr_done = false;
        if (!attr_done)
        {
            CU(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::total(512)));
            attr_done = true;
        }
        CU(h, launch_k(h, h->in_loop && pdl_on(h), kern, grid, block, smem, h->g, v, c.planes_per_cta, c.max_seg, h->ws, h->cm,
                       h->d_state, kc, h->d_hist, ghost_store));
    }
    return B200LS_OK;
}

// k_spmv4<BAL>: CTAs and planes per CTA of the balanced split
struct K4bCfg { int nctas, planes_per_cta; };
inline K4bCfg k4b_config(const b200ls_solver *h, int ty, int ctas_per_sm)
{
    const int64_t ntiles = (int64_t)((h->g.nx + 63) / 64) * ((h->g.ny + ty - 1) / ty);
    const int64_t total = ntiles * h->g.nzl;
    int64_t nctas = std::min<int64_t>((int64_t)h->num_sms * ctas_per_sm, std::max<int64_t>(1, total / 8));
    const int64_t ppc = (total + nctas - 1) / nctas;
    nctas = (total + ppc - 1) / ppc;
    return {(int)nctas, (int)ppc};
}

// ---- k_spmv4: tensor maps (one per vector and box shape, cached in the handle) and launch
using TmaEncodeFn = CUresult (*)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                 const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline TmaEncodeFn tma_encoder()
{
    static TmaEncodeFn fn = nullptr;
    static bool tried = false;
    if (!tried)
    {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<TmaEncodeFn>(p);
    }
    return fn;
}

// box (bw x bh x 1) over a solver-layout vector [nzl+2][ny][px]; the extents are (nx, ny, nzl+2), so the pad columns of
// the row pitch and everything outside the grid read as zeros
int tma_map_for(b200ls_solver *h, const double *vec, int bw, int bh, const TmaMap **out)
{
    const auto key = std::make_pair((const void *)vec, bw * 1024 + bh);
    auto it = h->tma_maps.find(key);
    if (it == h->tma_maps.end())
    {
        TmaEncodeFn enc = tma_encoder();
        if (!enc) return fail(h, B200LS_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
        TmaMap m;
        const cuuint64_t dims[3] = {(cuuint64_t)h->g.nx, (cuuint64_t)h->g.ny, (cuuint64_t)h->g.nzl + 2};
        const cuuint64_t strides[2] = {(cuuint64_t)h->g.px * 8, (cuuint64_t)h->g.plane * 8};
        const cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, 1};
        const cuuint32_t estr[3] = {1, 1, 1};
        // L2 promotion of the box rows (experiment knob B200LS_TMA_L2 = 0 none / 1 64 B / 2 128 B / 3 256 B; default 256 B)
        static const int l2p = getenv("B200LS_TMA_L2") ? atoi(getenv("B200LS_TMA_L2")) : 3;
        const CUtensorMapL2promotion prom = l2p == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : l2p == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
                                            : l2p == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
        const CUresult rc = enc(&m.m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, (void *)vec, dims, strides, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, prom, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (rc != CUDA_SUCCESS) return fail(h, B200LS_ERR_CUDA, "cuTensorMapEncodeTiled failed with code %d (box %d x %d)", (int)rc, bw, bh);
        it = h->tma_maps.emplace(key, m).first;
    }
    *out = &it->second;
    return B200LS_OK;
}

template <int TY, int S, int MINB, bool JAC, bool BAL = false>
int launch_spmv4_cfg(b200ls_solver *h, const VecSet &v, int ghost_store, dim3 grid, int kz)
{
    using L = Spmv4Smem<TY, S, JAC>;
    const SolveConsts kc = make_consts(h);
    dim3 block(32, TY + 1);
    Spmv4Maps maps;
    const TmaMap *m = nullptr;
    TRY(tma_map_for(h, v.r, L::BW, L::BH, &m));
    maps.r = *m;
    TRY(tma_map_for(h, v.p_in, L::BW, L::BH, &m));
    maps.p = *m;
    TRY(tma_map_for(h, v.x, L::BX, TY, &m));
    maps.x = *m;
    if (JAC) TRY(tma_map_for(h, v.dinv, L::BW, L::BH, &m));
    maps.d = *m;  // without Jacobi: any valid map (never read)
    auto kern = k_spmv4<TY, S, MINB, JAC, BAL>;
    static bool attr_done = false;
    if (!attr_done)
    {
        CU(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::total(2048)));
        attr_done = true;
    }
    int table = kz;
    if (BAL)
    {
        // balanced split: one resident wave of CTAs, equal ranges of the linearised (tile, plane) space
        if (h->g.nzl > 2048) return fail(h, B200LS_ERR_UNSUPPORTED, "balanced split: more than 2048 local planes");
        const K4bCfg c = k4b_config(h, TY, MINB);
        grid = dim3((unsigned)c.nctas);
        kz = c.planes_per_cta;
        table = h->g.nzl;
    }
    CU(h, launch_k(h, h->in_loop && pdl_on(h), kern, grid, block, L::total(table), maps, h->g, v, kz, h->ws, h->cm, h->d_state, kc,
                   h->d_hist, ghost_store));
    return B200LS_OK;
}

template <int TY, int KB, int MINB, bool JAC>
int launch_spmv5_cfg(b200ls_solver *h, const VecSet &v, int ghost_store, dim3 grid)
{
    const SolveConsts kc = make_consts(h);
    CU(h, launch_k(h, h->in_loop && pdl_on(h), k_spmv5<TY, KB, MINB, JAC>, grid, dim3(32, TY), 0, h->g, v, h->ws, h->cm, h->d_state, kc,
                   h->d_hist, ghost_store));
    return B200LS_OK;
}

template <bool JAC, bool APPLY>
int launch_spmv_t(b200ls_solver *h, const VecSet &v, int ghost_store)
{
    const TileCfg t = tile_cfg(h);
    const int kz = auto_kz_chunk(h);
    const int tile = k1_config(h).tile;
    const int rows = tile_rows(tile);
    dim3 grid((unsigned)((h->g.nx + 2 * t.txt - 1) / (2 * t.txt)), (unsigned)((h->g.ny + rows - 1) / rows),
              (unsigned)((h->g.nzl + kz - 1) / kz));
    dim3 block(t.txt, t.tyt);
    const SolveConsts kc = make_consts(h);
    h->launches++;
#define B200_SPMV_CASE(TXT, TYT)                                                                               \
    k_spmv<TXT, TYT, JAC, APPLY><<<grid, block, 0, h->stream>>>(h->g, v, kz, h->ws, h->cm, h->d_state, kc, \
                                                                 h->d_hist, ghost_store)
    if (APPLY && tile >= 10) return launch_spmv2_cfg<8, 4, 3, JAC, APPLY>(h, v, ghost_store, dim3(grid.x, (unsigned)((h->g.ny + 5) / 6), grid.z), kz);
    switch (tile)
    {
        case 0: B200_SPMV_CASE(32, 8); break;
        case 13: return launch_spmv2_cfg<6, 4, 4, JAC, false>(h, v, ghost_store, grid, kz);
        case 15: return launch_spmv2_cfg<8, 3, 3, JAC, false>(h, v, ghost_store, grid, kz);
        case 18: return launch_spmv2_cfg<12, 3, 2, JAC, false>(h, v, ghost_store, grid, kz);
        case 30: return launch_spmv3_cfg<12, 3, 2, JAC, false>(h, v, ghost_store);
        case 31: return launch_spmv3_cfg<8, 4, 3, JAC, false>(h, v, ghost_store);
        case 32: return launch_spmv3_cfg<12, 3, 2, JAC, true>(h, v, ghost_store);
        case 33: return launch_spmv3_cfg<8, 4, 3, JAC, true>(h, v, ghost_store);
        case 40: return launch_spmv4_cfg<8, 4, 2, JAC>(h, v, ghost_store, grid, kz);
        case 41: return launch_spmv4_cfg<4, 4, 4, JAC>(h, v, ghost_store, grid, kz);
        case 44: return launch_spmv4_cfg<4, 6, 4, JAC>(h, v, ghost_store, grid, kz);
        case 45: return launch_spmv4_cfg<4, 3, 4, JAC>(h, v, ghost_store, grid, kz);
        case 46: return launch_spmv4_cfg<4, 4, 4, JAC, true>(h, v, ghost_store, grid, kz);
        case 48: return launch_spmv4_cfg<2, 4, 6, JAC>(h, v, ghost_store, grid, kz);
        case 60: return launch_spmv4_cfg<6, 4, 3, JAC>(h, v, ghost_store, grid, kz);
        case 61: return launch_spmv4_cfg<6, 3, 3, JAC>(h, v, ghost_store, grid, kz);
        case 62: return launch_spmv4_cfg<5, 4, 3, JAC>(h, v, ghost_store, grid, kz);
        case 49: return launch_spmv4_cfg<2, 6, 6, JAC>(h, v, ghost_store, grid, kz);
        case 47: return launch_spmv4_cfg<4, 6, 4, JAC, true>(h, v, ghost_store, grid, kz);
        case 42: return launch_spmv4_cfg<8, 3, 3, JAC>(h, v, ghost_store, grid, kz);
        case 43: return launch_spmv4_cfg<16, 3, 1, JAC>(h, v, ghost_store, grid, kz);
        case 50: return launch_spmv5_cfg<8, 4, 2, JAC>(h, v, ghost_store, grid);
        case 51: return launch_spmv5_cfg<8, 8, 2, JAC>(h, v, ghost_store, grid);
        case 52: return launch_spmv5_cfg<4, 2, 4, JAC>(h, v, ghost_store, grid);
        case 53: return launch_spmv5_cfg<8, 16, 2, JAC>(h, v, ghost_store, grid);
        case 54: return launch_spmv5_cfg<8, 4, 3, JAC>(h, v, ghost_store, grid);
        case 55: return launch_spmv5_cfg<8, 8, 3, JAC>(h, v, ghost_store, grid);
        case 56: return launch_spmv5_cfg<8, 4, 4, JAC>(h, v, ghost_store, grid);
        default: return launch_spmv2_cfg<8, 4, 3, JAC, false>(h, v, ghost_store, grid, kz);
    }
#undef B200_SPMV_CASE
    return B200LS_OK;
}

int spmv_grid_blocks(const b200ls_solver *h)
{
    const TileCfg t = tile_cfg(h);
    const int kz = auto_kz_chunk(h);
    const int rows = tile_rows(k1_config(h).tile);
    return (int)(((h->g.nx + 2 * t.txt - 1) / (2 * t.txt)) * ((h->g.ny + rows - 1) / rows) * ((h->g.nzl + kz - 1) / kz));
}

int upd_grid_blocks(const b200ls_solver *h)
{
    if (h->upd_blocks > 0) return h->upd_blocks;
    // one resident wave (4 CTAs/SM); the block count is chosen so that every thread runs the same number
    // of guard-free 4-item trips (no serial scalar tail on small slabs)
    const int64_t items = (int64_t)(h->g.plane / 2) * h->g.nzl;
    const int64_t units = (items + 1023) / 1024;  // 256 threads x 4 items
    const int64_t trips = (units + (int64_t)h->num_sms * 4 - 1) / ((int64_t)h->num_sms * 4);
    return (int)std::max<int64_t>(1, (units + trips - 1) / trips);
}

template <bool JAC, bool INIT>
void launch_update_t(b200ls_solver *h, int fin_kind, bool push)
{
    UpdVecs v{h->r, h->w, h->dinv, h->upd_reverse};
    CommDev cm = h->cm;
    if (!push)
    {
        cm.r_ghost_dn = nullptr;
        cm.r_ghost_up = nullptr;
    }
    const SolveConsts kc = make_consts(h);
    int blocks = upd_grid_blocks(h);
    cm.npush = 0;
    static const bool push_extra = !(getenv("B200LS_PUSH_EXTRA") && atoi(getenv("B200LS_PUSH_EXTRA")) == 0);
    if (push && push_extra && h->upd_variant == 0 && h->upd_blocks <= 0 && h->g.nzl > 2)
    {
        // the boundary planes get their own CTAs ON TOP of the interior's grid, so the interior streams with as many
        // CTAs as on one GPU (pushers taken out of the grid: 46.5 us for a 256 x 256 x 128 slab, 33 us of HBM time)
        const int64_t nb2 = (int64_t)h->g.plane;  // two planes of plane/2 double2 items
        const int items = cm.push_items > 0 ? cm.push_items : 2;
        cm.npush = (int)std::max<int64_t>(1, std::min<int64_t>((nb2 + 256 * items - 1) / (256 * items), blocks / 2));
        blocks += cm.npush;
    }
    if (h->upd_variant == 1)
        k_update<JAC, INIT, 4><<<blocks, 256, 0, h->stream>>>(h->g, v, fin_kind, h->ws, cm, h->d_state, kc, h->d_hist);
    else if (JAC && h->upd_variant == 2)
    {
        // round-2 candidate: 1/diag rebuilt from the 1-D arrays instead of streamed (update_fly.cuh)
        const bool padded = h->g.px != h->g.nx;
        const bool psh = cm.r_ghost_dn || cm.r_ghost_up;
#define B200_UPDF(PAD, PSH)                                                                                        \
    launch_k(h, h->in_loop && pdl_on(h), k_update2f<INIT, PAD, PSH, 4>, dim3(blocks), dim3(256), 0, h->g, v, fin_kind, \
             h->ws, cm, h->d_state, kc, h->d_hist)
        if (padded && psh) B200_UPDF(true, true);
        else if (padded) B200_UPDF(true, false);
        else if (psh) B200_UPDF(false, true);
        else B200_UPDF(false, false);
#undef B200_UPDF
    }
    else
    {
        const bool padded = h->g.px != h->g.nx;
        const bool psh = cm.r_ghost_dn || cm.r_ghost_up;
#define B200_UPD(PAD, PSH)                                                                                            \
    launch_k(h, h->in_loop && pdl_on(h), k_update2<JAC, INIT, PAD, PSH, 4>, dim3(blocks), dim3(256), 0, h->g, v, fin_kind, \
             h->ws, cm, h->d_state, kc, h->d_hist)
        if (padded && psh) B200_UPD(true, true);
        else if (padded) B200_UPD(true, false);
        else if (psh) B200_UPD(false, true);
        else B200_UPD(false, false);
#undef B200_UPD
    }
    h->launches++;
}

// reduction epilogue on the host side: NCCL transport
int post_reduce(b200ls_solver *h, int kind)
{
    if (h->nranks > 1 && h->reduce_mode == B200LS_REDUCE_NCCL)
    {
        NcclApi &api = nccl_api();
        if (!h->nccl) return fail(h, B200LS_ERR_NCCL, "NCCL reduce mode selected but b200ls_nccl_init was not called");
        const int rc = api.AllReduce(h->d_sendrecv, h->d_sendrecv + B200_NSUM, B200_NSUM, kNcclDouble, kNcclSum,
                                     h->nccl, h->stream);
        if (rc != 0) return fail(h, B200LS_ERR_NCCL, "ncclAllReduce failed: %s", api.GetErrorString ? api.GetErrorString(rc) : "?");
        k_scalars<<<1, 32, 0, h->stream>>>(kind, h->d_sendrecv + B200_NSUM, h->d_state, make_consts(h), h->d_hist);
        h->launches += 2;
    }
    return B200LS_OK;
}

// halo of r by explicit peer copies (B200LS_HALO_MEMCPY)
int memcpy_halo(b200ls_solver *h, const double *vec)
{
    if (h->nranks <= 1 || h->halo_mode != B200LS_HALO_MEMCPY) return B200LS_OK;
    const size_t bytes = sizeof(double) * (size_t)h->g.plane;
    if (h->ghost_dn) CU(h, cudaMemcpyAsync(h->ghost_dn, vec + h->g.plane, bytes, cudaMemcpyDefault, h->stream));
    if (h->ghost_up)
        CU(h, cudaMemcpyAsync(h->ghost_up, vec + (size_t)h->g.nzl * h->g.plane, bytes, cudaMemcpyDefault, h->stream));
    return B200LS_OK;
}

inline bool kernel_push(const b200ls_solver *h) { return h->nranks > 1 && h->halo_mode == B200LS_HALO_STORE; }

int enqueue_update(b200ls_solver *h, bool init, int kind)
{
    const bool jac = h->opt.pc_type == B200LS_PC_JACOBI;
    {
        ProfScope ps(h, 1);
        const bool push = kernel_push(h);
        if (init)
        {
            if (jac) launch_update_t<true, true>(h, kind, push);
            else launch_update_t<false, true>(h, kind, push);
        }
        else
        {
            if (jac) launch_update_t<true, false>(h, kind, push);
            else launch_update_t<false, false>(h, kind, push);
        }
    }
    TRY(memcpy_halo(h, h->r));
    return post_reduce(h, kind);
}

int enqueue_spmv(b200ls_solver *h, int parity)
{
    const bool jac = h->opt.pc_type == B200LS_PC_JACOBI;
    VecSet v{h->r, h->p[parity], h->p[parity ^ 1], h->w, h->x, h->dinv};
    {
        ProfScope ps(h, 0);
        const int ghost_store = h->nranks > 1 ? 1 : 0;
        if (jac) TRY((launch_spmv_t<true, false>(h, v, ghost_store)));
        else TRY((launch_spmv_t<false, false>(h, v, ghost_store)));
    }
    return post_reduce(h, FIN_SPMV);
}

int enqueue_cg_iterations(b200ls_solver *h, int first_iter, int count)
{
    h->in_loop = true;
    int rc = B200LS_OK;
    for (int q = 0; q < count && rc == B200LS_OK; ++q)
    {
        rc = enqueue_spmv(h, (first_iter + q) & 1);
        if (rc == B200LS_OK) rc = enqueue_update(h, false, FIN_UPDATE);
    }
    h->in_loop = false;
    return rc;
}

// a graph of `count` CG iterations starting at an even iteration (buffer parity repeats every 2)
int launch_cg_batch(b200ls_solver *h, int first_iter, int count)
{
    // NCCL collectives are enqueued directly: capturing them hung in testing (NCCL 2.28 on this image), and
    // that transport is the literal north-star baseline, not the fast path.
    const bool nccl_path = h->nranks > 1 && h->reduce_mode == B200LS_REDUCE_NCCL;
    const bool graphable = h->use_graph && !h->profile && !nccl_path && (count % 2 == 0) && (first_iter % 2 == 0);
    if (!graphable) return enqueue_cg_iterations(h, first_iter, count);
    const unsigned long long key = (unsigned long long)count;
    if (!h->graph_exec || h->graph_key != key || h->graph_iters != count)
    {
        if (h->graph_exec)
        {
            cudaGraphExecDestroy(h->graph_exec);
            h->graph_exec = nullptr;
        }
        cudaGraph_t graph = nullptr;
        const int64_t launches_before = h->launches;
        CU(h, cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
        const int rc = enqueue_cg_iterations(h, 0, count);
        cudaError_t ce = cudaStreamEndCapture(h->stream, &graph);
        h->launches = launches_before;
        if (rc != B200LS_OK)
        {
            if (graph) cudaGraphDestroy(graph);
            return rc;
        }
        if (ce != cudaSuccess) return fail(h, B200LS_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(ce));
        ce = cudaGraphInstantiate(&h->graph_exec, graph, 0);
        cudaGraphDestroy(graph);
        if (ce != cudaSuccess) return fail(h, B200LS_ERR_CUDA, "graph instantiate failed: %s", cudaGetErrorString(ce));
        h->graph_key = key;
        h->graph_iters = count;
    }
    CU(h, cudaGraphLaunch(h->graph_exec, h->stream));
    int per_iter = 2;
    if (h->nranks > 1 && h->reduce_mode == B200LS_REDUCE_NCCL) per_iter += 4;
    h->launches += (int64_t)per_iter * count;
    return B200LS_OK;
}

void invalidate_graph(b200ls_solver *h)
{
    if (h->graph_exec)
    {
        cudaGraphExecDestroy(h->graph_exec);
        h->graph_exec = nullptr;
    }
}

int alloc_state(b200ls_solver *h)
{
    if (h->d_state) return B200LS_OK;
    CU(h, cudaMalloc(&h->d_state, sizeof(DevState)));
    CU(h, cudaMemset(h->d_state, 0, sizeof(DevState)));
    CU(h, cudaMallocHost(&h->h_state, 4 * sizeof(DevState)));
    for (auto &e : h->ev_slot) CU(h, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    h->max_blocks = 1 << 16;
    CU(h, cudaMalloc(&h->ws.partials, sizeof(double) * B200_NSUM * (size_t)h->max_blocks));
    CU(h, cudaMalloc(&h->ws.counter, sizeof(unsigned int) * 4));
    CU(h, cudaMemset(h->ws.counter, 0, sizeof(unsigned int) * 4));
    h->ws.trace = nullptr;
    CU(h, cudaMalloc(&h->d_sendrecv, sizeof(double) * 2 * B200_NSUM));
    CU(h, cudaMemset(h->d_sendrecv, 0, sizeof(double) * 2 * B200_NSUM));
    CU(h, cudaEventCreate(&h->ev_a));
    CU(h, cudaEventCreate(&h->ev_b));
    CU(h, cudaEventCreate(&h->ev_c));
    CU(h, cudaEventCreate(&h->ev_d));
    CU(h, cudaEventCreate(&h->ev_e0));
    CU(h, cudaEventCreate(&h->ev_e1));
    return B200LS_OK;
}

void build_commdev(b200ls_solver *h)
{
    CommDev &cm = h->cm;
    memset(&cm, 0, sizeof cm);
    cm.rank = h->rank;
    cm.nranks = h->nranks;
    cm.mode = 0;
    cm.push_items = getenv("B200LS_PUSH_ITEMS") ? atoi(getenv("B200LS_PUSH_ITEMS")) : 0;
    cm.dbg_flags = getenv("B200LS_DBG_FLAGS") ? atoi(getenv("B200LS_DBG_FLAGS")) : 0;
    cm.sendbuf = h->d_sendrecv;
    if (h->nranks > 1) cm.mode = (h->reduce_mode == B200LS_REDUCE_NCCL) ? 2 : 1;
}

// arena layout: [mailboxes 2 parities * nranks records * 16 LL words | pad to 256 B | halo flags (2 words) | pad to 256 B | r]
inline size_t arena_flags_offset(int nranks)
{
    const size_t b = sizeof(unsigned long long) * 2 * (size_t)nranks * B200_LLW;
    return (size_t)round_up((int64_t)b, 256);
}
inline size_t arena_head_bytes(int nranks) { return arena_flags_offset(nranks) + 256; }

// The halo hand-shake (CommDev::halo_flag_*) needs both sides: k_update2 publishes, every fused SpMV kernel waits.  Used with
// the fused transport and the flat update kernel; otherwise the pushers fence before their ticket as before.
void apply_halo_flag_policy(b200ls_solver *h)
{
    CommDev &cm = h->cm;
    cm.halo_flag_local = cm.halo_flag_dn_peer = cm.halo_flag_up_peer = nullptr;
    cm.push_counter = nullptr;
    if (!h->connected || h->nranks < 2 || h->reduce_mode != B200LS_REDUCE_P2P || h->halo_mode != B200LS_HALO_STORE) return;
    if (h->upd_variant != 0 || h->op != OP_STENCIL) return;  // the same on every rank: tuning keys are set collectively
    // Opt-in (B200LS_HALO_FLAGS=1): verified on 2 GPUs (tests/mgpu_check.py, all cases) but it buys nothing measurable --
    // k_update2 15.5 us with the hand-shake, 15.4 us with the pushers' fence before their ticket, 13.2 us with no system
    // fence at all (unsafe floor), same box (profiles/r02_trace_2gpu_slab.log): the fence is not what the reduction waits for.
    if (!getenv("B200LS_HALO_FLAGS")) return;
    const size_t fo = arena_flags_offset(h->nranks);
    auto flags = [&](int q) { return reinterpret_cast<unsigned long long *>(reinterpret_cast<char *>(h->peer_base[q]) + fo); };
    const bool perz = h->dim == 3 && h->per[2];
    int dn = h->rank - 1, up = h->rank + 1;
    if (dn < 0) dn = perz ? h->nranks - 1 : -1;
    if (up >= h->nranks) up = perz ? 0 : -1;
    cm.halo_flag_local = flags(h->rank);
    if (dn >= 0) cm.halo_flag_dn_peer = flags(dn) + 1;  // I am above it: its "ghost plane above is ready"
    if (up >= 0) cm.halo_flag_up_peer = flags(up) + 0;  // I am below it
    cm.push_counter = h->ws.counter + 1;
}

int setup_stencil_vectors(b200ls_solver *h)
{
    const size_t ve = h->vec_elems;
    const size_t head = arena_head_bytes(h->nranks);
    h->arena_bytes = head + sizeof(double) * ve;
    CU(h, cudaMalloc(&h->arena, h->arena_bytes));
    CU(h, cudaMemset(h->arena, 0, h->arena_bytes));
    h->r = reinterpret_cast<double *>(reinterpret_cast<char *>(h->arena) + head);
    for (int q = 0; q < 2; ++q)
    {
        CU(h, cudaMalloc(&h->p[q], sizeof(double) * ve));
        CU(h, cudaMemset(h->p[q], 0, sizeof(double) * ve));
    }
    CU(h, cudaMalloc(&h->w, sizeof(double) * ve));
    CU(h, cudaMemset(h->w, 0, sizeof(double) * ve));
    CU(h, cudaMalloc(&h->x, sizeof(double) * ve));
    CU(h, cudaMemset(h->x, 0, sizeof(double) * ve));
    CU(h, cudaMalloc(&h->stage, sizeof(double) * (size_t)std::max<int64_t>(h->nlocal, 1)));
    return B200LS_OK;
}

int ensure_jacobi(b200ls_solver *h)
{
    if (h->opt.pc_type != B200LS_PC_JACOBI || h->dinv || h->op != OP_STENCIL) return B200LS_OK;
    CU(h, cudaMalloc(&h->dinv, sizeof(double) * h->vec_elems));
    CU(h, cudaMemsetAsync(h->dinv, 0, sizeof(double) * h->vec_elems, h->stream));
    k_jacobi_setup<<<h->num_sms * 8, 256, 0, h->stream>>>(h->g, h->dinv);
    CU(h, cudaGetLastError());
    return B200LS_OK;
}

int check_state_err(b200ls_solver *h, const DevState &s)
{
    if (s.err) return fail(h, B200LS_ERR_CUDA, "cross-GPU reduction timed out (a peer rank never arrived)");
    return B200LS_OK;
}

// ------------------------------------------------------------------------------------------
// CG on the separable stencil
// ------------------------------------------------------------------------------------------
int solve_stencil_cg(b200ls_solver *h, const double *b_dev, double *x_dev)
{
    TRY(ensure_hist(h));
    TRY(ensure_jacobi(h));
    const bool multi = h->nranks > 1;
    if (multi && !h->connected) return fail(h, B200LS_ERR_ARG, "multi-GPU solver used before b200ls_comm_connect");
    h->launches = 0;
    CU(h, cudaEventRecord(h->ev_a, h->stream));
    k_state_reset<<<1, 32, 0, h->stream>>>(h->d_state);
    CU(h, cudaMemsetAsync(h->p[0], 0, sizeof(double) * h->vec_elems, h->stream));
    k_scatter<<<h->num_sms * 8, 256, 0, h->stream>>>(h->g, b_dev, h->r, h->x);
    h->launches += 2;
    if (h->has_const) TRY(enqueue_update(h, true, FIN_INIT_CENTRE));
    TRY(enqueue_update(h, true, FIN_INIT));
    CU(h, cudaEventRecord(h->ev_b, h->stream));

    // iterate in batches; the host only looks at the pinned snapshot of batch n-1 while batch n is
    // already queued (kernels of a finished solve return immediately)
    int check = std::max(2, h->opt.check_every);
    check += check & 1;
    int issued = 0, next_slot = 0;
    std::deque<int> outstanding;
    auto snapshot = [&]() -> int {
        const int s = next_slot;
        next_slot = (next_slot + 1) % 3;
        CU(h, cudaMemcpyAsync(&h->h_state[s], h->d_state, sizeof(DevState), cudaMemcpyDeviceToHost, h->stream));
        CU(h, cudaEventRecord(h->ev_slot[s], h->stream));
        outstanding.push_back(s);
        return B200LS_OK;
    };
    TRY(snapshot());
    for (;;)
    {
        while (outstanding.size() < 2 && issued < h->opt.max_it)
        {
            const int count = std::min(check, h->opt.max_it - issued);  // never queue past max_it
            TRY(launch_cg_batch(h, issued, count));
            issued += count;
            TRY(snapshot());
        }
        const int s = outstanding.front();
        outstanding.pop_front();
        CU(h, cudaEventSynchronize(h->ev_slot[s]));
        if (h->h_state[s].done) break;
        if (outstanding.empty() && issued >= h->opt.max_it)
            return fail(h, B200LS_ERR_CUDA, "internal: iteration budget exhausted without a reason");
    }
    CU(h, cudaEventRecord(h->ev_c, h->stream));
    k_xtail<<<h->num_sms * 8, 256, 0, h->stream>>>(h->g, h->x, h->p[0], h->p[1], h->d_state);
    k_gather<<<h->num_sms * 8, 256, 0, h->stream>>>(h->g, h->x, x_dev);
    h->launches += 2;
    CU(h, cudaMemcpyAsync(&h->h_state[3], h->d_state, sizeof(DevState), cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaEventRecord(h->ev_d, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    CU(h, cudaGetLastError());
    const DevState s = h->h_state[3];
    TRY(check_state_err(h, s));
    h->its = s.its;
    h->reason = s.reason;
    h->rnorm = s.dp;
    const int nh = std::min(s.nhist, h->hist_cap);
    h->history.resize((size_t)std::max(nh, 0));
    if (nh > 0) CU(h, cudaMemcpy(h->history.data(), h->d_hist, sizeof(double) * (size_t)nh, cudaMemcpyDeviceToHost));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, h->ev_a, h->ev_d);
    h->solve_ms = ms;
    cudaEventElapsedTime(&ms, h->ev_b, h->ev_c);
    h->loop_ms = ms;
    if (h->profile) resolve_profile(h);
    return B200LS_OK;
}

}  // namespace

#include "csr_solver.inc"
#include "sep_solver.inc"
#include "mg_solver.inc"
#include "dense_solver.inc"

// ------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------
extern "C" {

int b200ls_version(void) { return B200LS_VERSION; }

const char *b200ls_error_string(int code)
{
    switch (code)
    {
        case B200LS_OK: return "ok";
        case B200LS_ERR_ARG: return "bad argument or call order";
        case B200LS_ERR_CUDA: return "CUDA error or no device";
        case B200LS_ERR_UNSUPPORTED: return "unsupported option or operator";
        case B200LS_ERR_NCCL: return "NCCL error";
        case B200LS_ERR_DIVERGED: return "solver diverged (KSP reason < 0)";
        case B200LS_ERR_MISMATCH: return "matrix-free operator does not match the assembled matrix";
        case B200LS_ERR_PARSE: return "options text could not be parsed";
        default: return "unknown error";
    }
}

int b200ls_device_count(int *n)
{
    if (!n) return B200LS_ERR_ARG;
    int c = 0;
    if (cudaGetDeviceCount(&c) != cudaSuccess)
    {
        cudaGetLastError();
        *n = 0;
        return B200LS_ERR_CUDA;
    }
    *n = c;
    return B200LS_OK;
}

// parser::parseSubDomains / parseOneSubDomain (src/parser/parser.cpp:297-356) and
// misc::stretchGrid (include/petibm/misc.h:148-163)
int b200ls_axis_from_subdomains(double start, int nsub, const double *ends, const int *cells, const double *ratios,
                                double *dL_out, int cap, int *n_out)
{
    if (nsub < 0 || (nsub > 0 && (!ends || !cells || !ratios)) || !dL_out || !n_out) return B200LS_ERR_ARG;
    int ntot = 0;
    double bg = start;
    for (int s = 0; s < nsub; ++s)
    {
        const int n = cells[s];
        if (n <= 0 || ntot + n > cap) return B200LS_ERR_ARG;
        const double ed = ends[s], r = ratios[s];
        double *dL = dL_out + ntot;
        if (std::abs(r - 1.0) <= 1e-12)
        {
            const double hcell = (ed - bg) / n;
            for (int i = 0; i < n; ++i) dL[i] = hcell;
        }
        else
        {
            dL[0] = (ed - bg) * (r - 1.0) / (std::pow(r, n) - 1.0);
            for (int i = 1; i < n; ++i) dL[i] = dL[i - 1] * r;
        }
        ntot += n;
        bg = ed;
    }
    *n_out = ntot;
    return B200LS_OK;
}

void b200ls_default_options(b200ls_options *o)
{
    if (!o) return;
    o->ksp_type = B200LS_KSP_CG;  // linsolverksp.cpp:64
    o->pc_type = B200LS_PC_NONE;
    o->norm_type = B200LS_NORM_PRECONDITIONED;
    o->max_it = 10000;
    o->rtol = 1e-5;
    o->atol = 1e-50;
    o->divtol = 1e4;
    o->check_every = 32;
    o->variant = 0;
    o->mg_levels = 0;
    o->mg_smooth_its = 2;
    o->mg_coarse_its = 16;
}

static void set_err(char *errbuf, size_t errlen, const char *fmt, ...)
{
    if (!errbuf || !errlen) return;
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(errbuf, errlen, fmt, ap);
    va_end(ap);
}

// PETSc options-file syntax: one or more "-name [value]" per line, '#' starts a comment
// (what PetscOptionsInsertFile accepts, linsolverksp.cpp:58).
int b200ls_parse_options(const char *text, const char *prefix, b200ls_options *opts, char *errbuf, size_t errlen)
{
    if (!text || !prefix || !opts) return B200LS_ERR_ARG;
    std::vector<std::string> tok;
    {
        std::string cur;
        bool comment = false;
        for (const char *c = text;; ++c)
        {
            const char ch = *c;
            if (ch == '\0' || ch == '\n')
            {
                if (!cur.empty()) tok.push_back(cur);
                cur.clear();
                comment = false;
                if (ch == '\0') break;
                continue;
            }
            if (comment) continue;
            if (ch == '#' || ch == '!' || ch == '%')
            {
                if (cur.empty() || ch == '#')
                {
                    if (!cur.empty()) tok.push_back(cur);
                    cur.clear();
                    comment = true;
                    continue;
                }
            }
            if (ch == ' ' || ch == '\t' || ch == '\r')
            {
                if (!cur.empty()) tok.push_back(cur);
                cur.clear();
                continue;
            }
            cur.push_back(ch);
        }
    }
    auto is_name = [](const std::string &s) {
        return s.size() >= 2 && s[0] == '-' && !(s[1] >= '0' && s[1] <= '9') && s[1] != '.';
    };
    const std::string pre = std::string("-") + prefix;
    for (size_t q = 0; q < tok.size(); ++q)
    {
        if (!is_name(tok[q])) { set_err(errbuf, errlen, "stray value '%s'", tok[q].c_str()); return B200LS_ERR_PARSE; }
        const std::string name = tok[q];
        std::string val;
        bool has_val = false;
        if (q + 1 < tok.size() && !is_name(tok[q + 1]))
        {
            val = tok[++q];
            has_val = true;
        }
        if (name.compare(0, pre.size(), pre) != 0) continue;  // another solver's / global option
        const std::string key = name.substr(pre.size());
        auto need = [&]() -> bool {
            if (!has_val) set_err(errbuf, errlen, "option %s needs a value", name.c_str());
            return has_val;
        };
        auto num = [&](double &out) -> bool {
            if (!need()) return false;
            char *end = nullptr;
            out = strtod(val.c_str(), &end);
            if (end == val.c_str() || *end != '\0')
            {
                set_err(errbuf, errlen, "option %s: '%s' is not a number", name.c_str(), val.c_str());
                return false;
            }
            return true;
        };
        double d = 0.0;
        if (key == "ksp_type")
        {
            if (!need()) return B200LS_ERR_PARSE;
            if (val == "cg") opts->ksp_type = B200LS_KSP_CG;
            else if (val == "bcgs") opts->ksp_type = B200LS_KSP_BCGS;
            else if (val == "preonly") opts->ksp_type = B200LS_KSP_PREONLY;
            else { set_err(errbuf, errlen, "-%sksp_type %s is not implemented by the B200 backend (cg, bcgs, preonly)", prefix, val.c_str()); return B200LS_ERR_UNSUPPORTED; }
        }
        else if (key == "pc_type")
        {
            if (!need()) return B200LS_ERR_PARSE;
            if (val == "none") opts->pc_type = B200LS_PC_NONE;
            else if (val == "jacobi") opts->pc_type = B200LS_PC_JACOBI;
            else if (val == "mg") opts->pc_type = B200LS_PC_MG;
            else if (val == "lu" || val == "cholesky") opts->pc_type = B200LS_PC_LU;
            else { set_err(errbuf, errlen, "-%spc_type %s is not implemented by the B200 backend (none, jacobi, mg, lu)", prefix, val.c_str()); return B200LS_ERR_UNSUPPORTED; }
        }
        // geometric multigrid (extension; PETSc's PCMG option names, only the combination that is implemented)
        else if (key == "pc_mg_levels") { if (!num(d)) return B200LS_ERR_PARSE; opts->mg_levels = (int)d; }
        else if (key == "mg_levels_ksp_max_it") { if (!num(d)) return B200LS_ERR_PARSE; opts->mg_smooth_its = (int)d; }
        else if (key == "mg_coarse_ksp_max_it") { if (!num(d)) return B200LS_ERR_PARSE; opts->mg_coarse_its = (int)d; }
        else if (key == "mg_levels_ksp_type" || key == "mg_coarse_ksp_type")
        {
            if (!need()) return B200LS_ERR_PARSE;
            if (val != "chebyshev") { set_err(errbuf, errlen, "-%s%s %s unsupported (chebyshev)", prefix, key.c_str(), val.c_str()); return B200LS_ERR_UNSUPPORTED; }
        }
        else if (key == "mg_levels_pc_type" || key == "mg_coarse_pc_type")
        {
            if (!need()) return B200LS_ERR_PARSE;
            if (val != "jacobi") { set_err(errbuf, errlen, "-%s%s %s unsupported (jacobi)", prefix, key.c_str(), val.c_str()); return B200LS_ERR_UNSUPPORTED; }
        }
        else if (key == "pc_mg_cycle_type")
        {
            if (!need()) return B200LS_ERR_PARSE;
            if (val != "v") { set_err(errbuf, errlen, "-%spc_mg_cycle_type %s unsupported (v)", prefix, val.c_str()); return B200LS_ERR_UNSUPPORTED; }
        }
        else if (key == "pc_factor_mat_solver_type")
        {
            // which package factorises (superlu_dist, mumps, petsc) is PETSc's business: here the direct solve is the
            // dense factorisation of dense_kernels.cuh
            if (!need()) return B200LS_ERR_PARSE;
        }
        else if (key == "pc_jacobi_type")
        {
            if (!need()) return B200LS_ERR_PARSE;
            if (val != "diagonal") { set_err(errbuf, errlen, "-%spc_jacobi_type %s unsupported (diagonal)", prefix, val.c_str()); return B200LS_ERR_UNSUPPORTED; }
        }
        else if (key == "ksp_norm_type")
        {
            if (!need()) return B200LS_ERR_PARSE;
            if (val == "preconditioned") opts->norm_type = B200LS_NORM_PRECONDITIONED;
            else if (val == "unpreconditioned") opts->norm_type = B200LS_NORM_UNPRECONDITIONED;
            else if (val == "natural") opts->norm_type = B200LS_NORM_NATURAL;
            else { set_err(errbuf, errlen, "-%sksp_norm_type %s unsupported", prefix, val.c_str()); return B200LS_ERR_UNSUPPORTED; }
        }
        else if (key == "ksp_rtol") { if (!num(d)) return B200LS_ERR_PARSE; opts->rtol = d; }
        else if (key == "ksp_atol") { if (!num(d)) return B200LS_ERR_PARSE; opts->atol = d; }
        else if (key == "ksp_divtol") { if (!num(d)) return B200LS_ERR_PARSE; opts->divtol = d; }
        else if (key == "ksp_max_it") { if (!num(d)) return B200LS_ERR_PARSE; opts->max_it = (int)d; }
        else if (key == "ksp_initial_guess_nonzero")
        {
            if (has_val && val != "0" && val != "false" && val != "no")
            {
                set_err(errbuf, errlen, "-%sksp_initial_guess_nonzero is unsupported: PetIBM always starts from zero", prefix);
                return B200LS_ERR_UNSUPPORTED;
            }
            if (!has_val) { set_err(errbuf, errlen, "-%sksp_initial_guess_nonzero is unsupported", prefix); return B200LS_ERR_UNSUPPORTED; }
        }
        else if (key == "ksp_monitor" || key == "ksp_view" || key == "ksp_converged_reason" || key == "ksp_reuse_preconditioner")
        {
            // diagnostics / no-ops for this backend
        }
        else if (key == "b200_check_every") { if (!num(d)) return B200LS_ERR_PARSE; opts->check_every = (int)d; }
        else if (key == "b200_variant") { if (!num(d)) return B200LS_ERR_PARSE; opts->variant = (int)d; }
        else
        {
            set_err(errbuf, errlen, "option %s is not implemented by the B200 backend; no silent fallback", name.c_str());
            return B200LS_ERR_UNSUPPORTED;
        }
    }
    if ((opts->ksp_type == B200LS_KSP_PREONLY) != (opts->pc_type == B200LS_PC_LU))
    {
        set_err(errbuf, errlen, "-%sksp_type preonly and -%spc_type lu are only implemented together (direct solve)", prefix, prefix);
        return B200LS_ERR_UNSUPPORTED;
    }
    return B200LS_OK;
}

int b200ls_create(b200ls_solver **out, int device)
{
    if (!out) return B200LS_ERR_ARG;
    *out = nullptr;
    int cnt = 0;
    if (cudaGetDeviceCount(&cnt) != cudaSuccess || cnt <= 0)
    {
        cudaGetLastError();
        return B200LS_ERR_CUDA;  // no CPU fallback
    }
    if (device < 0 || device >= cnt) return B200LS_ERR_ARG;
    b200ls_solver *h = new (std::nothrow) b200ls_solver();
    if (!h) return B200LS_ERR_ARG;
    b200ls_default_options(&h->opt);
    h->device = device;
    if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess)
    {
        cudaGetLastError();
        delete h;
        return B200LS_ERR_CUDA;
    }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) h->num_sms = prop.multiProcessorCount;
    const int rc = alloc_state(h);
    if (rc != B200LS_OK)
    {
        delete h;
        return rc;
    }
    if (const char *e = getenv("B200LS_TILE")) h->tile = atoi(e);
    if (const char *e = getenv("B200LS_KZ_CHUNK")) h->kz_chunk = atoi(e);
    if (const char *e = getenv("B200LS_UPD_VARIANT")) h->upd_variant = atoi(e);
    // experiment switches of the multigrid path (same as b200ls_set_tuning): lets the whole device test-suite run with them
    if (const char *e = getenv("B200LS_MG_GRAPH")) h->mg_graph = atoi(e);
    if (const char *e = getenv("B200LS_CSR_GRAPH")) h->csr_graph = atoi(e);
    if (const char *e = getenv("B200LS_SEP_TILE")) h->sep_tile = atoi(e);
    if (const char *e = getenv("B200LS_SEP_ZCHUNK")) h->sep_zchunk = atoi(e);
    if (const char *e = getenv("B200LS_SEP_STAGES")) h->sep_stages = atoi(e);
    if (const char *e = getenv("B200LS_MG_TAIL")) h->mg_tail = atoi(e);
    if (const char *e = getenv("B200LS_MG_FUSE")) h->mg_fuse = atoi(e);
    build_commdev(h);
    *out = h;
    return B200LS_OK;
}

int b200ls_destroy(b200ls_solver *h)
{
    if (!h) return B200LS_OK;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->nccl && nccl_api().ok) nccl_api().CommDestroy(h->nccl);
    free_vectors(h);
    if (h->d_state) cudaFree(h->d_state);
    if (h->h_state) cudaFreeHost(h->h_state);
    for (auto &e : h->ev_slot)
        if (e) cudaEventDestroy(e);
    if (h->ws.partials) cudaFree(h->ws.partials);
    if (h->ws.counter) cudaFree(h->ws.counter);
    if (h->ws.trace) cudaFree(h->ws.trace);
    if (h->d_hist) cudaFree(h->d_hist);
    if (h->d_sendrecv) cudaFree(h->d_sendrecv);
    if (h->flush_buf) cudaFree(h->flush_buf);
    for (cudaEvent_t e : {h->ev_a, h->ev_b, h->ev_c, h->ev_d, h->ev_e0, h->ev_e1})
        if (e) cudaEventDestroy(e);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return B200LS_OK;
}

const char *b200ls_last_error(const b200ls_solver *h) { return h ? h->err.c_str() : "null handle"; }

int b200ls_set_options(b200ls_solver *h, const b200ls_options *o)
{
    if (!h || !o) return B200LS_ERR_ARG;
    if (o->ksp_type != B200LS_KSP_CG && o->ksp_type != B200LS_KSP_BCGS && o->ksp_type != B200LS_KSP_PREONLY)
        return fail(h, B200LS_ERR_UNSUPPORTED, "ksp_type %d", o->ksp_type);
    if (o->pc_type != B200LS_PC_NONE && o->pc_type != B200LS_PC_JACOBI && o->pc_type != B200LS_PC_MG && o->pc_type != B200LS_PC_LU)
        return fail(h, B200LS_ERR_UNSUPPORTED, "pc_type %d", o->pc_type);
    if ((o->ksp_type == B200LS_KSP_PREONLY) != (o->pc_type == B200LS_PC_LU))
        return fail(h, B200LS_ERR_UNSUPPORTED, "ksp_type preonly and pc_type lu are only implemented together (direct solve)");
    if (o->pc_type == B200LS_PC_MG && o->ksp_type != B200LS_KSP_CG) return fail(h, B200LS_ERR_UNSUPPORTED, "pc_type mg is used with ksp_type cg");
    if (o->mg_levels < 0 || o->mg_levels > 32) return fail(h, B200LS_ERR_ARG, "pc_mg_levels %d", o->mg_levels);
    if (o->norm_type < B200LS_NORM_PRECONDITIONED || o->norm_type > B200LS_NORM_NATURAL)
        return fail(h, B200LS_ERR_UNSUPPORTED, "norm_type %d (KSP_NORM_NONE is not supported)", o->norm_type);
    if (o->max_it < 0) return fail(h, B200LS_ERR_ARG, "max_it < 0");
    h->opt = *o;
    if (h->opt.check_every <= 0) h->opt.check_every = 32;
    invalidate_graph(h);
    return B200LS_OK;
}

int b200ls_get_options(const b200ls_solver *h, b200ls_options *o)
{
    if (!h || !o) return B200LS_ERR_ARG;
    *o = h->opt;
    return B200LS_OK;
}

int b200ls_set_tuning(b200ls_solver *h, const char *key, int value)
{
    if (!h || !key) return B200LS_ERR_ARG;
    const std::string k(key);
    if (k == "kz_chunk") h->kz_chunk = value;
    else if (k == "upd_blocks") h->upd_blocks = value;
    else if (k == "tile") h->tile = value;
    else if (k == "upd_variant") h->upd_variant = value;
    else if (k == "upd_reverse") h->upd_reverse = value;
    else if (k == "use_graph") h->use_graph = value;
    else if (k == "use_pdl") h->use_pdl = value;
    else if (k == "mg_graph") h->mg_graph = value;
    else if (k == "csr_graph") h->csr_graph = value;
    else if (k == "sep_tile") h->sep_tile = value;
    else if (k == "sep_zchunk") h->sep_zchunk = value;
    else if (k == "sep_stages") h->sep_stages = value;
    else if (k == "mg_tail") h->mg_tail = value;
    else if (k == "mg_fuse") h->mg_fuse = value;
    else return fail(h, B200LS_ERR_ARG, "unknown tuning key %s", key);
    apply_halo_flag_policy(h);
    invalidate_graph(h);
    return B200LS_OK;
}

// ---- communicator -----------------------------------------------------------------------
int b200ls_comm_init(b200ls_solver *h, int rank, int nranks, int reduce_mode, int halo_mode)
{
    if (!h) return B200LS_ERR_ARG;
    if (nranks < 1 || nranks > B200_MAX_RANKS || rank < 0 || rank >= nranks) return fail(h, B200LS_ERR_ARG, "bad rank/nranks %d/%d", rank, nranks);
    if (h->op != OP_NONE) return fail(h, B200LS_ERR_ARG, "b200ls_comm_init must precede the operator set-up");
    if (reduce_mode != B200LS_REDUCE_P2P && reduce_mode != B200LS_REDUCE_NCCL) return fail(h, B200LS_ERR_ARG, "bad reduce_mode");
    if (halo_mode != B200LS_HALO_STORE && halo_mode != B200LS_HALO_MEMCPY) return fail(h, B200LS_ERR_ARG, "bad halo_mode");
    if (reduce_mode == B200LS_REDUCE_P2P && halo_mode == B200LS_HALO_MEMCPY && nranks > 1)
        return fail(h, B200LS_ERR_UNSUPPORTED, "memcpy halos need the NCCL all-reduce as their cross-GPU ordering point");
    h->rank = rank;
    h->nranks = nranks;
    h->reduce_mode = reduce_mode;
    h->halo_mode = halo_mode;
    build_commdev(h);
    return B200LS_OK;
}

int b200ls_comm_export(b200ls_solver *h, void *handle64)
{
    if (!h || !handle64) return B200LS_ERR_ARG;
    if (!h->arena) return fail(h, B200LS_ERR_ARG, "no arena: set the operator first");
    cudaSetDevice(h->device);
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t hd;
    CU(h, cudaIpcGetMemHandle(&hd, h->arena));
    memcpy(handle64, &hd, 64);
    return B200LS_OK;
}

int b200ls_comm_connect(b200ls_solver *h, const void *handles, int nranks)
{
    if (!h || !handles) return B200LS_ERR_ARG;
    if (nranks != h->nranks) return fail(h, B200LS_ERR_ARG, "nranks mismatch");
    if (!h->arena) return fail(h, B200LS_ERR_ARG, "no arena: set the operator first");
    cudaSetDevice(h->device);
    h->peer_base.assign((size_t)nranks, nullptr);
    for (int q = 0; q < nranks; ++q)
    {
        if (q == h->rank)
        {
            h->peer_base[q] = h->arena;
            continue;
        }
        cudaIpcMemHandle_t hd;
        memcpy(&hd, (const char *)handles + 64 * (size_t)q, 64);
        void *ptr = nullptr;
        CU(h, cudaIpcOpenMemHandle(&ptr, hd, cudaIpcMemLazyEnablePeerAccess));
        h->peer_base[q] = ptr;
    }
    h->connected = true;
    const size_t head = arena_head_bytes(nranks);
    CommDev &cm = h->cm;
    for (int q = 0; q < nranks; ++q)
    {
        cm.mbox_peer[q] = reinterpret_cast<unsigned long long *>(h->peer_base[q]);
    }
    cm.mbox_local = cm.mbox_peer[h->rank];
    // neighbours along the slab axis; periodic wrap closes the ring
    const bool perz = h->dim == 3 && h->per[2];
    int dn = h->rank - 1, up = h->rank + 1;
    if (dn < 0) dn = perz ? nranks - 1 : -1;
    if (up >= nranks) up = perz ? 0 : -1;
    auto rbase = [&](int q) { return reinterpret_cast<double *>(reinterpret_cast<char *>(h->peer_base[q]) + head); };
    h->ghost_dn = h->ghost_up = nullptr;
    // our top plane goes to the BOTTOM ghost plane (storage plane 0) of the neighbour above
    if (up >= 0) h->ghost_up = rbase(up);
    // our plane 0 goes to the TOP ghost plane (storage plane nzl_peer + 1) of the neighbour below; slab sizes
    // follow the PETSc ownership rule (checked in b200ls_set_poisson_stencil), so nzl_peer is computable here
    if (dn >= 0)
    {
        const int64_t M = h->n[2], m = nranks;
        const int64_t nzl_peer = M / m + ((dn < (M % m)) ? 1 : 0);
        h->ghost_dn = rbase(dn) + (size_t)(nzl_peer + 1) * (size_t)h->g.plane;
    }
    if (kernel_push(h))
    {
        cm.r_ghost_dn = h->ghost_dn;
        cm.r_ghost_up = h->ghost_up;
    }
    apply_halo_flag_policy(h);
    invalidate_graph(h);
    return B200LS_OK;
}

// Drop the mappings of the peers' arenas (collective: every rank calls it, then the host transport
// barriers, and only then may any rank free or replace its own arena, e.g. in a second setMatrix).
int b200ls_comm_disconnect(b200ls_solver *h)
{
    if (!h) return B200LS_ERR_ARG;
    if (!h->connected) return B200LS_OK;
    cudaSetDevice(h->device);
    CU(h, cudaStreamSynchronize(h->stream));
    for (int q = 0; q < (int)h->peer_base.size(); ++q)
        if (q != h->rank && h->peer_base[q]) cudaIpcCloseMemHandle(h->peer_base[q]);
    h->peer_base.clear();
    h->connected = false;
    h->ghost_dn = h->ghost_up = nullptr;
    build_commdev(h);
    invalidate_graph(h);
    return B200LS_OK;
}

int b200ls_nccl_unique_id(void *id128)
{
    if (!id128) return B200LS_ERR_ARG;
    NcclApi &api = nccl_api();
    if (!api.ok) return B200LS_ERR_NCCL;
    NcclUniqueId id;
    if (api.GetUniqueId(&id) != 0) return B200LS_ERR_NCCL;
    memcpy(id128, &id, 128);
    return B200LS_OK;
}

int b200ls_nccl_init(b200ls_solver *h, const void *id128)
{
    if (!h || !id128) return B200LS_ERR_ARG;
    NcclApi &api = nccl_api();
    if (!api.ok) return fail(h, B200LS_ERR_NCCL, "libnccl.so.2 could not be loaded");
    cudaSetDevice(h->device);
    NcclUniqueId id;
    memcpy(&id, id128, 128);
    const int rc = api.CommInitRank(&h->nccl, h->nranks, id, h->rank);
    if (rc != 0) return fail(h, B200LS_ERR_NCCL, "ncclCommInitRank failed: %s", api.GetErrorString ? api.GetErrorString(rc) : "?");
    return B200LS_OK;
}

// ---- operator ---------------------------------------------------------------------------
int b200ls_set_poisson_stencil(b200ls_solver *h, int dim, const int64_t n[3], const int periodic[3], const double *dx,
                               const double *dy, const double *dz, double dt, int64_t slab_lo, int64_t slab_hi)
{
    if (!h || !n || !periodic || !dx || !dy) return B200LS_ERR_ARG;
    if (dim != 2 && dim != 3) return fail(h, B200LS_ERR_ARG, "dim must be 2 or 3");
    if (dim == 3 && !dz) return fail(h, B200LS_ERR_ARG, "dz missing");
    cudaSetDevice(h->device);
    const int64_t nx = n[0], ny = n[1], nz = (dim == 3) ? n[2] : 1;
    if (nx < 1 || ny < 1 || nz < 1) return fail(h, B200LS_ERR_ARG, "empty grid");
    if (nx > (1 << 24) || ny > (1 << 24) || nz > (1 << 24) || nx * ny * nz > (int64_t)3 << 31)
        return fail(h, B200LS_ERR_UNSUPPORTED, "grid too large for 32-bit item indexing");
    for (int d = 0; d < dim; ++d)
        if (periodic[d] && n[d] < 3) return fail(h, B200LS_ERR_UNSUPPORTED, "periodic axis %d needs at least 3 cells", d);
    if (dim == 2 && h->nranks > 1)
    {
        // several GPUs: a 2-D grid is the 3-D code on (nx, 1, ny) with y-slabs.  The arithmetic is bit-identical to
        // the 2-D form: dy' = {1.0} makes every product an exact multiplication by one, the y' faces are walls
        // (coefficient 0, exact additions of zero), the D-row / CSR-row orders keep x before the slab axis.
        const int64_t n3[3] = {n[0], 1, n[1]};
        const int per3[3] = {periodic[0], 0, periodic[1]};
        const double one = 1.0;
        return b200ls_set_poisson_stencil(h, 3, n3, per3, dx, &one, dy, dt, slab_lo, slab_hi);
    }
    if (dim == 2)
    {
        slab_lo = 0;
        slab_hi = 1;
    }
    if (slab_lo < 0 || slab_hi > nz || slab_lo >= slab_hi) return fail(h, B200LS_ERR_ARG, "bad slab [%lld,%lld)", (long long)slab_lo, (long long)slab_hi);
    if (h->nranks == 1 && (slab_lo != 0 || slab_hi != nz)) return fail(h, B200LS_ERR_ARG, "single-GPU solver must own the whole grid");
    if (h->nranks > 1)
    {
        // PETSc DMDA ownership rule along the slab axis: first M mod m ranks get one more
        const int64_t M = nz, m = h->nranks;
        if (M < m) return fail(h, B200LS_ERR_ARG, "fewer planes than ranks");
        const int64_t base = M / m, rem = M % m;
        const int64_t lo = h->rank * base + std::min<int64_t>(h->rank, rem);
        const int64_t hi = lo + base + (h->rank < rem ? 1 : 0);
        if (lo != slab_lo || hi != slab_hi)
            return fail(h, B200LS_ERR_ARG, "slab [%lld,%lld) is not the DMDA range [%lld,%lld) of rank %d", (long long)slab_lo,
                        (long long)slab_hi, (long long)lo, (long long)hi, h->rank);
    }
    free_vectors(h);
    h->dim = dim;
    h->n[0] = nx;
    h->n[1] = ny;
    h->n[2] = nz;
    for (int d = 0; d < 3; ++d) h->per[d] = (d < dim) ? (periodic[d] != 0) : 0;
    h->slab_lo = slab_lo;
    h->slab_hi = slab_hi;
    h->dt = dt;
    h->hdx.assign(dx, dx + nx);
    h->hdy.assign(dy, dy + ny);
    if (dim == 3) h->hdz.assign(dz, dz + nz);
    else h->hdz.assign(1, 1.0);  // cartesianmesh.cpp:91-99,166-171
    // face coefficients dt*(1/h), h = 0.5*(dL[s-1]+dL[s]) (cartesianmesh.cpp:237-247, creategradient.cpp:72,
    // createbn.cpp:49); g[s] = minus face of cell s
    auto faces = [&](const std::vector<double> &d, bool per, bool active) {
        const size_t m = d.size();
        std::vector<double> g(m + 1, 0.0);
        if (!active) return g;
        for (size_t s = 1; s < m; ++s)
        {
            const double hh = 0.5 * (d[s] + d[s - 1]);
            const double inv = 1.0 / hh;
            g[s] = dt * inv;
        }
        if (per)
        {
            const double hh = 0.5 * (d[0] + d[m - 1]);
            const double inv = 1.0 / hh;
            g[0] = g[m] = dt * inv;
        }
        return g;
    };
    h->hgx = faces(h->hdx, h->per[0], true);
    h->hgy = faces(h->hdy, h->per[1], true);
    h->hgz = faces(h->hdz, h->per[2], dim == 3);
    // device copies
    const size_t tot = (size_t)(2 * (nx + ny + nz) + 3);
    std::vector<double> pack;
    pack.reserve(tot);
    size_t o_dx = 0, o_dy, o_dz, o_gx, o_gy, o_gz;
    pack.insert(pack.end(), h->hdx.begin(), h->hdx.end());
    o_dy = pack.size();
    pack.insert(pack.end(), h->hdy.begin(), h->hdy.end());
    o_dz = pack.size();
    pack.insert(pack.end(), h->hdz.begin(), h->hdz.end());
    o_gx = pack.size();
    pack.insert(pack.end(), h->hgx.begin(), h->hgx.end());
    o_gy = pack.size();
    pack.insert(pack.end(), h->hgy.begin(), h->hgy.end());
    o_gz = pack.size();
    pack.insert(pack.end(), h->hgz.begin(), h->hgz.end());
    CU(h, cudaMalloc(&h->d_axes, sizeof(double) * pack.size()));
    CU(h, cudaMemcpy(h->d_axes, pack.data(), sizeof(double) * pack.size(), cudaMemcpyHostToDevice));
    GridDev &g = h->g;
    g.nx = (int)nx;
    g.ny = (int)ny;
    g.nzl = (int)(slab_hi - slab_lo);
    g.px = (int)round_up(nx, 16);
    g.plane = (long long)g.px * g.ny;
    g.perx = h->per[0];
    g.pery = h->per[1];
    g.perz_wrap = (h->per[2] && h->nranks == 1) ? 1 : 0;
    g.kz0 = (int)slab_lo;
    g.nzg = (int)nz;
    g.wrapz_lo = (h->per[2] && slab_lo == 0) ? 1 : 0;
    g.wrapz_hi = (h->per[2] && slab_hi == nz) ? 1 : 0;
    g.dx = h->d_axes + o_dx;
    g.dy = h->d_axes + o_dy;
    g.dz = h->d_axes + o_dz;
    g.gx = h->d_axes + o_gx;
    g.gy = h->d_axes + o_gy;
    g.gz = h->d_axes + o_gz;
    h->nlocal = nx * ny * (slab_hi - slab_lo);
    h->nglobal = nx * ny * nz;
    h->vec_elems = (size_t)g.plane * (size_t)(g.nzl + 2);
    h->op = OP_STENCIL;
    TRY(setup_stencil_vectors(h));
    build_commdev(h);
    invalidate_graph(h);
    if (spmv_grid_blocks(h) > h->max_blocks) return fail(h, B200LS_ERR_UNSUPPORTED, "grid needs more reduction slots than allocated");
    return B200LS_OK;
}

namespace {
// |a - b| in units of the last place of b (b != 0, finite)
inline double ulps_apart(double a, double b)
{
    int e;
    std::frexp(b, &e);
    return std::fabs(a - b) / std::ldexp(1.0, e - 53);
}

// natural_rows == nullptr: rows are this rank's slab in natural order
int verify_rows(b200ls_solver *h, int64_t nrows, const int64_t *natural_rows, const int64_t *rowptr, const int32_t *col,
                const double *val, int diag_ulps, double *max_abs_diff)
{
    if (!h || !rowptr || !col || !val) return B200LS_ERR_ARG;
    if (h->op != OP_STENCIL) return fail(h, B200LS_ERR_ARG, "no stencil operator to verify");
    if (!natural_rows && nrows != h->nlocal)
        return fail(h, B200LS_ERR_MISMATCH, "row count %lld != %lld", (long long)nrows, (long long)h->nlocal);
    const int64_t nx = h->n[0], ny = h->n[1], nz = h->n[2];
    const int dim = h->dim;
    double worst = 0.0;
    int64_t bad_row = -1;
    const volatile double *dx = h->hdx.data(), *dy = h->hdy.data(), *dz = h->hdz.data();
    const volatile double *gx = h->hgx.data(), *gy = h->hgy.data(), *gz = h->hgz.data();
    for (int64_t row = 0; row < nrows; ++row)
    {
        int64_t i, j, k;
        if (natural_rows)
        {
            const int64_t nat = natural_rows[row];
            if (nat < 0 || nat >= nx * ny * nz) return fail(h, B200LS_ERR_ARG, "natural row index out of range");
            i = nat % nx;
            j = (nat / nx) % ny;
            k = nat / (nx * ny);
        }
        else
        {
            i = row % nx;
            j = (row / nx) % ny;
            k = h->slab_lo + row / (nx * ny);
        }
        // expected entries: (global column, value); same products as k_spmv
        int64_t ecol[7];
        double eval[7];
        int ne = 0;
        const volatile double ayz = dy[j] * dz[k], axz = dx[i] * dz[k], axy = dx[i] * dy[j];
        const volatile double cxm = ayz * gx[i], cxp = ayz * gx[i + 1];
        const volatile double cym = axz * gy[j], cyp = axz * gy[j + 1];
        const volatile double czm = axy * gz[k], czp = axy * gz[k + 1];
        const bool wy = h->per[1] && j == 0, wz = h->per[2] && k == 0;
        volatile double dg = cxm + cxp;
        dg = dg + (wy ? cyp : cym);
        dg = dg + (wy ? cym : cyp);
        if (dim == 3)
        {
            dg = dg + (wz ? czp : czm);
            dg = dg + (wz ? czm : czp);
        }
        const int64_t self = i + nx * (j + ny * k);
        ecol[ne] = self;
        eval[ne++] = -dg;
        auto add = [&](double c, int64_t ii, int64_t jj, int64_t kk) {
            if (c == 0.0) return;
            ecol[ne] = ii + nx * (jj + ny * kk);
            eval[ne++] = c;
        };
        add(cxm, (i - 1 + nx) % nx, j, k);
        add(cxp, (i + 1) % nx, j, k);
        add(cym, i, (j - 1 + ny) % ny, k);
        add(cyp, i, (j + 1) % ny, k);
        if (dim == 3)
        {
            add(czm, i, j, (k - 1 + nz) % nz);
            add(czp, i, j, (k + 1) % nz);
        }
        // compare as sets (assembled rows may carry explicit zeros)
        int matched = 0;
        for (int64_t q = rowptr[row]; q < rowptr[row + 1]; ++q)
        {
            int e = -1;
            for (int t = 0; t < ne; ++t)
                if (ecol[t] == (int64_t)col[q]) e = t;
            double diff = (e >= 0) ? std::fabs(val[q] - eval[e]) : std::fabs(val[q]);
            if (e >= 0) matched++;
            // partition-dependent accumulation order of the diagonal (see b200ls.h)
            if (e == 0 && diag_ulps > 0 && diff != 0.0 && eval[0] != 0.0 && ulps_apart(val[q], eval[0]) <= (double)diag_ulps)
                diff = 0.0;
            if (diff > worst || (diff != diff))
            {
                worst = (diff != diff) ? INFINITY : diff;
                bad_row = row;
            }
        }
        if (matched != ne)
        {
            // an expected non-zero entry is missing from the assembled row
            worst = INFINITY;
            bad_row = row;
        }
    }
    if (max_abs_diff) *max_abs_diff = worst;
    if (worst != 0.0) return fail(h, B200LS_ERR_MISMATCH, "assembled matrix differs from the separable stencil (max |diff| %.3e, e.g. local row %lld)", worst, (long long)bad_row);
    return B200LS_OK;
}
}  // namespace

int b200ls_verify_csr(b200ls_solver *h, int64_t nrows, const int64_t *rowptr, const int32_t *col, const double *val,
                      double *max_abs_diff)
{
    return verify_rows(h, nrows, nullptr, rowptr, col, val, 0, max_abs_diff);
}

int b200ls_verify_csr_rows(b200ls_solver *h, int64_t nrows, const int64_t *natural_rows, const int64_t *rowptr,
                           const int32_t *col, const double *val, int diag_ulps, double *max_abs_diff)
{
    if (!natural_rows) return B200LS_ERR_ARG;
    return verify_rows(h, nrows, natural_rows, rowptr, col, val, diag_ulps, max_abs_diff);
}

int b200ls_set_nullspace(b200ls_solver *h, int has_const, int nvecs, const double *vecs)
{
    if (!h) return B200LS_ERR_ARG;
    if (h->op == OP_NONE) return fail(h, B200LS_ERR_ARG, "set the operator before its null space");
    if (nvecs < 0 || (nvecs > 0 && !vecs)) return B200LS_ERR_ARG;
    cudaSetDevice(h->device);
    if (h->op == OP_STENCIL && nvecs > 0) return fail(h, B200LS_ERR_UNSUPPORTED, "explicit null-space vectors need the CSR operator");
    if (h->op == OP_CSR || h->op == OP_SEP) TRY(csr_set_nullvecs(h, has_const ? 1 : 0, nvecs, vecs));  // validates before anything changes
    h->has_const = has_const ? 1 : 0;
    invalidate_graph(h);
    return B200LS_OK;
}

int b200ls_apply(b200ls_solver *h, const double *x_host, double *y_host)
{
    if (!h || !x_host || !y_host) return B200LS_ERR_ARG;
    cudaSetDevice(h->device);
    if (h->op == OP_CSR) return csr_apply_host(h, x_host, y_host);
    if (h->op == OP_SEP) return sep_apply_host(h, x_host, y_host);
    if (h->op != OP_STENCIL) return fail(h, B200LS_ERR_ARG, "no operator");
    if (h->nranks > 1 && !h->connected) return fail(h, B200LS_ERR_ARG, "not connected");
    const size_t nb = sizeof(double) * (size_t)h->nlocal;
    CU(h, cudaMemcpyAsync(h->stage, x_host, nb, cudaMemcpyHostToDevice, h->stream));
    k_state_reset<<<1, 32, 0, h->stream>>>(h->d_state);
    k_scatter<<<h->num_sms * 8, 256, 0, h->stream>>>(h->g, h->stage, h->r, nullptr);
    if (h->nranks > 1)
    {
        if (h->reduce_mode == B200LS_REDUCE_P2P)
        {
            k_barrier<<<1, 32, 0, h->stream>>>(h->cm, h->d_state);
            k_push_halo<<<h->num_sms, 256, 0, h->stream>>>(h->g, h->r, h->ghost_dn, h->ghost_up);
            k_barrier<<<1, 32, 0, h->stream>>>(h->cm, h->d_state);
        }
        else
        {
            TRY(post_reduce(h, FIN_INIT_CENTRE));
            k_push_halo<<<h->num_sms, 256, 0, h->stream>>>(h->g, h->r, h->ghost_dn, h->ghost_up);
            TRY(post_reduce(h, FIN_INIT_CENTRE));
        }
    }
    VecSet v{h->r, nullptr, nullptr, h->w, nullptr, nullptr};
    TRY((launch_spmv_t<false, true>(h, v, 0)));
    if (h->nranks > 1)
    {
        if (h->reduce_mode == B200LS_REDUCE_P2P) k_barrier<<<1, 32, 0, h->stream>>>(h->cm, h->d_state);
        else TRY(post_reduce(h, FIN_INIT_CENTRE));
    }
    k_gather<<<h->num_sms * 8, 256, 0, h->stream>>>(h->g, h->w, h->stage);
    CU(h, cudaMemcpyAsync(y_host, h->stage, nb, cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaMemcpyAsync(&h->h_state[0], h->d_state, sizeof(DevState), cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    CU(h, cudaGetLastError());
    return check_state_err(h, h->h_state[0]);
}

// ---- solve ------------------------------------------------------------------------------
int b200ls_solve_device(b200ls_solver *h, const double *b_dev, double *x_dev)
{
    if (!h || !b_dev || !x_dev) return B200LS_ERR_ARG;
    cudaSetDevice(h->device);
    int rc;
    if (h->opt.ksp_type == B200LS_KSP_PREONLY)
        rc = dense_solve(h, b_dev, x_dev);
    else if (h->op == OP_STENCIL)
    {
        if (h->opt.ksp_type != B200LS_KSP_CG)
            return fail(h, B200LS_ERR_UNSUPPORTED, "the separable stencil operator is solved with cg; bcgs needs the CSR operator");
        rc = (h->opt.pc_type == B200LS_PC_MG) ? solve_stencil_pcg_mg(h, b_dev, x_dev) : solve_stencil_cg(h, b_dev, x_dev);
    }
    else if (h->op == OP_CSR)
        rc = csr_solve(h, b_dev, x_dev);
    else if (h->op == OP_SEP)
        rc = sep_solve(h, b_dev, x_dev);
    else
        return fail(h, B200LS_ERR_ARG, "no operator set");
    if (rc != B200LS_OK) return rc;
    if (h->reason < 0) return fail(h, B200LS_ERR_DIVERGED, "diverged: KSPConvergedReason %d after %d iterations, residual %.6e", h->reason, h->its, h->rnorm);
    return B200LS_OK;
}

int b200ls_solve(b200ls_solver *h, const double *b_host, double *x_host)
{
    if (!h || !b_host || !x_host) return B200LS_ERR_ARG;
    if (h->op == OP_NONE) return fail(h, B200LS_ERR_ARG, "no operator set");
    cudaSetDevice(h->device);
    const size_t nb = sizeof(double) * (size_t)h->nlocal;
    CU(h, cudaEventRecord(h->ev_e0, h->stream));
    CU(h, cudaMemcpyAsync(h->stage, b_host, nb, cudaMemcpyHostToDevice, h->stream));
    const int rc = b200ls_solve_device(h, h->stage, h->stage);
    if (rc != B200LS_OK && rc != B200LS_ERR_DIVERGED) return rc;
    CU(h, cudaMemcpyAsync(x_host, h->stage, nb, cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaEventRecord(h->ev_e1, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, h->ev_e0, h->ev_e1);
    h->e2e_ms = ms;
    return rc;
}

int b200ls_get_iters(const b200ls_solver *h, int *its)
{
    if (!h || !its) return B200LS_ERR_ARG;
    *its = h->its;
    return B200LS_OK;
}
int b200ls_get_residual(const b200ls_solver *h, double *rnorm)
{
    if (!h || !rnorm) return B200LS_ERR_ARG;
    *rnorm = h->rnorm;
    return B200LS_OK;
}
int b200ls_get_reason(const b200ls_solver *h, int *reason)
{
    if (!h || !reason) return B200LS_ERR_ARG;
    *reason = h->reason;
    return B200LS_OK;
}
int b200ls_get_history(const b200ls_solver *h, double *buf, int cap, int *n)
{
    if (!h || !n) return B200LS_ERR_ARG;
    *n = (int)h->history.size();
    if (buf && cap > 0) memcpy(buf, h->history.data(), sizeof(double) * (size_t)std::min<int>(cap, *n));
    return B200LS_OK;
}

// ---- measurement ------------------------------------------------------------------------
int b200ls_get_timing(const b200ls_solver *h, double *solve_ms, double *loop_ms, int64_t *launches)
{
    if (!h) return B200LS_ERR_ARG;
    if (solve_ms) *solve_ms = h->solve_ms;
    if (loop_ms) *loop_ms = h->loop_ms;
    if (launches) *launches = h->launches;
    return B200LS_OK;
}
int b200ls_get_e2e_ms(const b200ls_solver *h, double *e2e_ms)
{
    if (!h || !e2e_ms) return B200LS_ERR_ARG;
    *e2e_ms = h->e2e_ms;
    return B200LS_OK;
}
int b200ls_set_profile(b200ls_solver *h, int enable)
{
    if (!h) return B200LS_ERR_ARG;
    h->profile = enable ? 1 : 0;
    for (int q = 0; q < 4; ++q)
    {
        h->prof_ms[q] = 0.0;
        h->prof_cnt[q] = 0;
    }
    return B200LS_OK;
}
int b200ls_get_profile(const b200ls_solver *h, int kclass, double *total_ms, int64_t *count)
{
    if (!h || kclass < 0 || kclass > 3) return B200LS_ERR_ARG;
    if (total_ms) *total_ms = h->prof_ms[kclass];
    if (count) *count = h->prof_cnt[kclass];
    return B200LS_OK;
}

int b200ls_time_kernel(b200ls_solver *h, int kclass, int reps, int flush_l2, double *avg_ms)
{
    if (!h || !avg_ms || reps <= 0) return B200LS_ERR_ARG;
    if (h->op != OP_STENCIL) return fail(h, B200LS_ERR_ARG, "kernel timing needs the stencil operator");
    if (h->nranks > 1) return fail(h, B200LS_ERR_UNSUPPORTED, "kernel timing is single-GPU");
    cudaSetDevice(h->device);
    TRY(ensure_hist(h));
    TRY(ensure_jacobi(h));
    if (flush_l2 && !h->flush_buf)
    {
        h->flush_elems = (size_t)40 << 20;  // 320 MB > 126 MB L2
        CU(h, cudaMalloc(&h->flush_buf, sizeof(double) * h->flush_elems));
    }
    // a benign state: never "done", a = 0 so the vectors keep their values, b = 0
    DevState s{};
    s.betaold = 1.0;
    s.a = 0.0;
    s.b = 0.0;
    s.pending = 1;
    double total = 0.0;
    const b200ls_options keep = h->opt;
    h->opt.max_it = 1 << 30;
    h->opt.rtol = 0.0;
    h->opt.atol = 0.0;
    h->opt.divtol = 1e300;
    cudaEvent_t e0, e1;
    CU(h, cudaEventCreate(&e0));
    CU(h, cudaEventCreate(&e1));
    const bool jac = h->opt.pc_type == B200LS_PC_JACOBI;
    for (int q = -2; q < reps; ++q)
    {
        CU(h, cudaMemcpyAsync(h->d_state, &s, sizeof s, cudaMemcpyHostToDevice, h->stream));
        if (flush_l2) k_fill<<<h->num_sms * 8, 256, 0, h->stream>>>(h->flush_buf, (long long)h->flush_elems, 0.0);
        CU(h, cudaEventRecord(e0, h->stream));
        if (kclass == 0)
        {
            VecSet v{h->r, h->p[0], h->p[1], h->w, h->x, h->dinv};
            if (jac) TRY((launch_spmv_t<true, false>(h, v, 0)));
            else TRY((launch_spmv_t<false, false>(h, v, 0)));
        }
        else if (kclass == 1)
        {
            if (jac) launch_update_t<true, false>(h, FIN_UPDATE, false);
            else launch_update_t<false, false>(h, FIN_UPDATE, false);
        }
        else
        {
            VecSet v{h->r, nullptr, nullptr, h->w, nullptr, nullptr};
            TRY((launch_spmv_t<false, true>(h, v, 0)));
        }
        CU(h, cudaEventRecord(e1, h->stream));
        CU(h, cudaEventSynchronize(e1));
        float ms = 0.f;
        CU(h, cudaEventElapsedTime(&ms, e0, e1));
        if (q >= 0) total += ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    h->opt = keep;
    CU(h, cudaGetLastError());
    *avg_ms = total / reps;
    return B200LS_OK;
}

int b200ls_set_trace(b200ls_solver *h, int capacity)
{
    if (!h || capacity < 0) return B200LS_ERR_ARG;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    if (h->ws.trace) cudaFree(h->ws.trace);
    h->ws.trace = nullptr;
    invalidate_graph(h);
    if (capacity == 0) return B200LS_OK;
    const size_t n = 8 + 5 * (size_t)capacity;
    CU(h, cudaMalloc(&h->ws.trace, sizeof(unsigned long long) * n));
    CU(h, cudaMemset(h->ws.trace, 0, sizeof(unsigned long long) * n));
    const unsigned long long cap = (unsigned long long)capacity;
    CU(h, cudaMemcpy(h->ws.trace + 2, &cap, sizeof cap, cudaMemcpyHostToDevice));
    return B200LS_OK;
}

int b200ls_get_trace(b200ls_solver *h, unsigned long long *buf, int capacity, int *n)
{
    if (!h || !n) return B200LS_ERR_ARG;
    *n = 0;
    if (!h->ws.trace) return B200LS_OK;
    cudaSetDevice(h->device);
    CU(h, cudaStreamSynchronize(h->stream));
    unsigned long long head[8];
    CU(h, cudaMemcpy(head, h->ws.trace, sizeof head, cudaMemcpyDeviceToHost));
    const int have = (int)std::min<unsigned long long>(head[0], head[2]);
    *n = have;
    const int take = std::min(have, capacity);
    if (buf && take > 0)
        CU(h, cudaMemcpy(buf, h->ws.trace + 8, sizeof(unsigned long long) * 5 * (size_t)take, cudaMemcpyDeviceToHost));
    const unsigned long long zero = 0;
    CU(h, cudaMemcpy(h->ws.trace, &zero, sizeof zero, cudaMemcpyHostToDevice));
    return B200LS_OK;
}

void *b200ls_stream(b200ls_solver *h) { return h ? (void *)h->stream : nullptr; }

}  // extern "C"

#include "ops_solver.inc"
