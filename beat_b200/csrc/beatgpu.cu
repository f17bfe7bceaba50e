// beatgpu.cu -- host side of libbeatgpu.so: context, operand upload, kernel launches, the extern "C" ABI
//               declared in include/beatgpu.h.  sm_100a only; no CPU fallback anywhere.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "beatgpu.h"
#include "sweep.cuh"
#include "stack.cuh"
#include "aux.cuh"
#include "gemm.cuh"
#include "probe.cuh"
#include "geom.cuh"

using namespace beatgpu;

namespace {

thread_local std::string g_create_error;

struct WaveMap {
    int nt = 0, ns = 0, interp = 0;
    bool has_station = false;
    int* d_station_idx = nullptr;
    int* d_hyper_idx = nullptr;
    int* d_nsamp = nullptr;
    // library
    void* G[BEATGPU_MAX_SLIPVARS] = {nullptr, nullptr, nullptr};
    bool G_owned[BEATGPU_MAX_SLIPVARS] = {false, false, false};
    int store_dtype = -1;
    int64_t dims[5] = {0, 0, 0, 0, 0};
    long ld = 0;
    double dur_min = 0, dur_step = 1, st_min = 0, st_step = 1;
    bool axes_set = false;
    // data / weights
    double* d_data = nullptr;
    double* d_W = nullptr;
    size_t W_bytes = 0;
    double* d_slog_pdet = nullptr;
    int misfit_mode = -1, bw = 0, dense_upper = 1;
    int out_ofs = 0;
    int geom = -1;                  // >= 0: geometry-mode wavemap (index into ctx->gwmaps); no GF library of its own
};

// geometry mode (BASELINE config 2): a pyrocko-style GF store and the per-wavemap static operands
struct GeomStore {
    GeomStoreDev dev;
    float* d_traces = nullptr;
    int* d_itmin = nullptr;
    int* d_nsamp = nullptr;
    int max_nsamp = 0;
};

struct GeomWaveMap {
    int store_id = 0, nt = 0, nr = 0, nraw_max = 0, n4 = 0;
    double *d_rcv_lat = nullptr, *d_rcv_lon = nullptr;
    int *d_rcv_itmin = nullptr, *d_rcv_nraw = nullptr, *d_rcv_first = nullptr, *d_tgt_of = nullptr;
    float* d_tgt_f = nullptr;
    int *d_tgt_nraw = nullptr, *d_tgt_ibeg = nullptr;
    double* d_taper = nullptr;      // nullptr: all factors are 1 (chop between b and c)
    // station corrections (time_shift hierarchical): windows follow arrival + shift of the chain
    bool has_station = false;
    double *d_rcv_arrival = nullptr, *d_tgt_arrival = nullptr;
    int *d_rcv_station = nullptr, *d_tgt_station = nullptr, *d_tgt_rcv = nullptr;
    double abcd[4] = {0, 0, 0, 0};
    int chop_lo = 1, chop_hi = 2, nraw_cap = 0;
    int nsec = 1, ord = 1, demean = 0;
    double fb[kGeomMaxSec][kGeomMaxOrder + 1], fa[kGeomMaxSec][kGeomMaxOrder + 1];
};

struct Geodetic {
    bool set = false;
    int nobs = 0, nds = 0, max_n = 0;
    double* G[BEATGPU_MAX_SLIPVARS] = {nullptr, nullptr, nullptr};
    double *d_data = nullptr, *d_odw = nullptr, *d_UT = nullptr, *d_slog = nullptr;
    long* d_UT_ofs = nullptr;
    int *d_lo = nullptr, *d_hi = nullptr, *d_upper = nullptr, *d_nsamp = nullptr, *d_hyper_idx = nullptr;
    std::vector<int> lo, hi, h_upper, h_nsamp, h_hyper_idx;
    std::vector<double> h_slog;
    std::vector<long> ut_ofs;
    long ut_total = 0;
    int out_ofs = 0;
};

struct Laplacian {
    bool set = false;
    double* d_LT = nullptr;
    double sdet = 0;
    int hyper_idx = 0;
    int out_ofs = 0;
};

}  // namespace

struct beatgpu_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = true;
    std::string err;
    cudaDeviceProp prop;
    // fault
    int nsf = 0, np_total = 0, max_np_sf = 0, max_diag = 0;
    int sweep_pack = 1;             // BEATGPU_SWEEP_PACK=0: one chain per warp in the rupture sweep
    int misfit_warp = 1;            // BEATGPU_MISFIT_WARP=0: CTA-per-(chain, target) misfit pass for every trace length
    std::vector<int> h_nd, h_ns, h_pofs;
    std::vector<double> h_psize;
    int *d_nd = nullptr, *d_ns = nullptr, *d_pofs = nullptr;
    double* d_psize = nullptr;
    // layout
    bool layout_set = false;
    beatgpu_layout layout;
    double* d_fixed = nullptr;
    long canon_slip[BEATGPU_MAX_SLIPVARS], canon_dur = 0, canon_vel = 0, canon_nstr = 0, canon_ndip = 0, canon_time = 0,
         canon_hyp = 0, canon_ts = 0, canon_len = 0;
    // composites
    std::vector<WaveMap> wmaps;
    Geodetic geo;
    Laplacian lap;
    // scratch (grown on demand)
    int cap_B = 0;
    double* d_q = nullptr;          // [cap_B, n_params]  (host-entry staging)
    double* d_logpts = nullptr;     // [cap_B, n_out]
    double* d_like = nullptr;       // [cap_B]
    double* d_t0 = nullptr;         // [cap_B, np_total]
    unsigned char* d_bad = nullptr; // [cap_B]
    int cap_n_out = 0, cap_n_params = 0;
    unsigned long long* d_viol = nullptr;
    // generic scratch for the unit entries
    void* d_tmp[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    size_t tmp_bytes[6] = {0, 0, 0, 0, 0, 0};
    // accounting
    long long n_launches = 0;
    bool persistent = false;        // BEATGPU_PERSISTENT=1: fused kernel with one resident CTA wave walking the items
    int stack_mode = 1;             // 0 = fused kernel (CTA per target x chain), 1 = patch-chunked warps + misfit pass
    int chunk_patches = 32;         // BEATGPU_CHUNK: target patches per chunk (<= 32)
    int chunk_occ = 5;              // BEATGPU_CHUNK_OCC: 5 / 6 / 7 CTAs per SM guaranteed by the chunk kernel's register allocation
    int geo_mode = 1;               // 1 = FP64 tensor-core GEMM tiles (BEATGPU_GEO_MODE=mma), 0 = one CTA per (chain, dataset)
    double* d_partial = nullptr;    // [B, nt, nchunk, ns] scratch of the chunked path
    size_t partial_bytes = 0;
    void* d_plan = nullptr;         // per-(chain, patch) plans of the chunked path (plan_cache_kernel)
    size_t plan_bytes = 0;
    int plan_cache_on = 1;          // BEATGPU_PLAN_CACHE=0: every (target, chunk, chain) warp plans its own patches
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;      // event pair of the LAST evaluation (aliases into the ring below)
    bool ev_valid = false;
    // ring of event pairs: one per fused evaluation, so the kernels' share of a timed loop can be summed afterwards
    std::vector<cudaEvent_t> tev;                  // 2 * kTimingRing events, created on first use
    int tev_next = 0, tev_pending = 0;             // next slot; evaluations recorded since the last reset (<= kTimingRing)
    double l2_frac = 0.6;                          // BEATGPU_L2_FRAC: share of L2 one patch chunk of the library may span (nominal bytes; chains touch part of a block)
    bool chunk_forced = false;                     // BEATGPU_CHUNK given: no L2-derived chunk
    cudaStream_t copy_stream = nullptr;            // second stream of the host-pointer entry (q columns behind the sweep)
    cudaEvent_t copy_done = nullptr, copy_go = nullptr;
    int split_h2d = 1;                             // BEATGPU_SPLIT_H2D=0: one plain copy of q
    cudaEvent_t wait_before_stack = nullptr;       // set by the host-pointer entry: the columns of q the stacking needs arrive on copy_stream
    // geometry mode
    std::vector<GeomStore> gstores;
    std::vector<GeomWaveMap> gwmaps;
    bool glayout_set = false;
    beatgpu_geom_layout glayout;
    double* d_gfixed = nullptr;
    double ev_lat = 0, ev_lon = 0, stf_anchor = -1.0;
    int stf_type = BEATGPU_STF_HALFSINUSOID, off_peak_ratio = -1;   // beatgpu_geom_set_stf
    double* d_peak_ratio = nullptr;                                 // fixed TriangularSTF.peak_ratio (one value)
    void *d_rplan = nullptr, *d_cplan = nullptr, *d_rawT = nullptr, *d_gmean = nullptr;
    size_t g_rplan_bytes = 0, g_cplan_bytes = 0, g_raw_bytes = 0, g_mean_bytes = 0;
    unsigned int* d_gerr = nullptr;
    int filter_cap = 80;            // BEATGPU_FILTER_CAP: 80 = register-capped filter kernel (6 CTAs per SM, measured faster), 0 = uncapped
    int geom_half = 3;              // BEATGPU_GEOM_HALF: rows per pipeline half of the delay-and-sum kernel (3 or 4; 3 measured faster)
};

namespace {

int fail(beatgpu_ctx* c, int code, const char* fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (c) c->err = buf; else g_create_error = buf;
    return code;
}

#define CK(call)                                                                                      \
    do {                                                                                              \
        cudaError_t e__ = (call);                                                                     \
        if (e__ != cudaSuccess) {                                                                     \
            (void)cudaGetLastError(); /* clear the (non-sticky) error so later launch checks are not poisoned */ \
            return fail(ctx, BEATGPU_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), \
                        __FILE__, __LINE__);                                                          \
        }                                                                                             \
    } while (0)

#define CKL()                                                                                          \
    do {                                                                                               \
        cudaError_t e__ = cudaGetLastError();                                                          \
        if (e__ != cudaSuccess)                                                                        \
            return fail(ctx, BEATGPU_E_CUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__), \
                        __FILE__, __LINE__);                                                           \
        ctx->n_launches++;                                                                             \
    } while (0)

template <typename T>
int upload_vec(beatgpu_ctx* ctx, T** dptr, const T* h, size_t n)
{
    if (*dptr) { cudaFree(*dptr); *dptr = nullptr; }
    if (n == 0) return BEATGPU_OK;
    CK(cudaMalloc((void**)dptr, n * sizeof(T)));
    CK(cudaMemcpyAsync(*dptr, h, n * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return BEATGPU_OK;
}

int ensure_tmp(beatgpu_ctx* ctx, int slot, size_t bytes)
{
    if (ctx->tmp_bytes[slot] >= bytes) return BEATGPU_OK;
    if (ctx->d_tmp[slot]) cudaFree(ctx->d_tmp[slot]);
    ctx->d_tmp[slot] = nullptr;
    ctx->tmp_bytes[slot] = 0;
    CK(cudaMalloc(&ctx->d_tmp[slot], bytes));
    ctx->tmp_bytes[slot] = bytes;
    return BEATGPU_OK;
}

int n_outputs(const beatgpu_ctx* ctx)
{
    int n = 0;
    for (const auto& w : ctx->wmaps) n += w.nt;
    if (ctx->geo.set) n += ctx->geo.nds;
    if (ctx->lap.set) n += 1;
    return n;
}

void assign_out_offsets(beatgpu_ctx* ctx)
{
    int o = 0;
    for (auto& w : ctx->wmaps) { w.out_ofs = o; o += w.nt; }
    ctx->geo.out_ofs = o;
    if (ctx->geo.set) o += ctx->geo.nds;
    ctx->lap.out_ofs = o;
}

int ensure_scratch(beatgpu_ctx* ctx, int B)
{
    const int n_out = std::max(1, n_outputs(ctx));
    const int n_par = std::max(ctx->layout_set ? ctx->layout.n_params : 1, ctx->glayout_set ? ctx->glayout.n_params : 1);
    if (B <= ctx->cap_B && n_out <= ctx->cap_n_out && n_par <= ctx->cap_n_params) return BEATGPU_OK;
    const int nb = std::max(B, ctx->cap_B);
    cudaFree(ctx->d_q); cudaFree(ctx->d_logpts); cudaFree(ctx->d_like); cudaFree(ctx->d_t0); cudaFree(ctx->d_bad);
    ctx->d_q = ctx->d_logpts = ctx->d_like = ctx->d_t0 = nullptr;
    ctx->d_bad = nullptr;
    ctx->cap_B = 0;
    CK(cudaMalloc((void**)&ctx->d_q, (size_t)nb * n_par * sizeof(double)));
    CK(cudaMalloc((void**)&ctx->d_logpts, (size_t)nb * n_out * sizeof(double)));
    CK(cudaMalloc((void**)&ctx->d_like, (size_t)nb * sizeof(double)));
    CK(cudaMalloc((void**)&ctx->d_t0, (size_t)nb * std::max(1, ctx->np_total) * sizeof(double)));
    CK(cudaMalloc((void**)&ctx->d_bad, (size_t)nb));
    ctx->cap_B = nb;
    ctx->cap_n_out = n_out;
    ctx->cap_n_params = n_par;
    return BEATGPU_OK;
}

// pointer + per-chain stride of a model variable inside q (or the fixed vector)
struct VarRef { const double* p; long stride; };
VarRef var_ref(const beatgpu_ctx* ctx, const double* q_dev, int off, long canon)
{
    if (off >= 0) return {q_dev + off, (long)ctx->layout.n_params};
    return {ctx->d_fixed + canon, 0};
}

int launch_sweep(beatgpu_ctx* ctx, SweepArgs& a, int n_items)
{
    // lanes per chain: the longest grid diagonal (min(n_dip, n_strike) cells are independent per relaxation step);
    // 32 / W chains share a warp.  BEATGPU_SWEEP_PACK=0: one chain per warp
    // Packing trades latency (a warp runs until the slowest of its chains has converged: 121 vs 100 us for 500 chains)
    // for issue slots (137 vs 172 us for 4000 chains): pack once the unpacked grid would put >= 12 warps on every SM.
    int W = 32;
    if (ctx->sweep_pack && ctx->max_diag > 0 && ctx->max_diag <= 16 && n_items >= 12 * ctx->prop.multiProcessorCount) W = ctx->max_diag;
    const int cpw = 32 / W;
    a.group_width = W;
    const int n_warps = (n_items + cpw - 1) / cpw;
    size_t per_warp = (size_t)4 * a.max_np_sf * sizeof(double) * cpw;
    int wpb = (int)std::min<size_t>(8, std::max<size_t>(1, (48 * 1024) / per_warp));
    // latency-bound kernel: spread the warps over all SMs before stacking several on one
    wpb = std::max(1, std::min(wpb, n_warps / (2 * ctx->prop.multiProcessorCount)));
    size_t smem = per_warp * wpb;
    if (smem > 48 * 1024) {
        if (smem > (size_t)ctx->prop.sharedMemPerBlockOptin)
            return fail(ctx, BEATGPU_E_ARG, "subfault of %d patches needs %zu B shared memory per warp (> %zu)",
                        a.max_np_sf, smem, (size_t)ctx->prop.sharedMemPerBlockOptin);
        CK(cudaFuncSetAttribute(chain_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    n_items = n_warps;
    int blocks = (n_items + wpb - 1) / wpb;
    chain_sweep_kernel<<<blocks, wpb * 32, smem, ctx->stream>>>(a, wpb);
    CKL();
    return BEATGPU_OK;
}

template <typename T, int K, bool WS>
int launch_stack_nvar(beatgpu_ctx* ctx, const StackArgs& a)
{
    const size_t smem = WS ? 0 : (size_t)a.ns * sizeof(double);
    const long grid = (long)a.nt * a.B;
    if (grid > 2147483647L) return fail(ctx, BEATGPU_E_ARG, "grid too large: nt*B = %ld", grid);
#define LAUNCH(NV)                                                                                                   \
    do {                                                                                                             \
        if (smem > 48 * 1024)                                                                                        \
            CK(cudaFuncSetAttribute(gf_stack_misfit_kernel<T, K, NV, WS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        long g = grid;                                                                                               \
        if (ctx->persistent) {                                                                                       \
            int per_sm = 0;                                                                                          \
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gf_stack_misfit_kernel<T, K, NV, WS>, kStackThreads, smem)); \
            g = std::min<long>(grid, (long)std::max(1, per_sm) * ctx->prop.multiProcessorCount);                     \
        }                                                                                                            \
        gf_stack_misfit_kernel<T, K, NV, WS><<<(unsigned)g, kStackThreads, smem, ctx->stream>>>(a);                 \
    } while (0)
    switch (a.nvar) {
        case 1: LAUNCH(1); break;
        case 2: LAUNCH(2); break;
        case 3: LAUNCH(3); break;
        default: return fail(ctx, BEATGPU_E_ARG, "n_slipvars must be 1..3, got %d", a.nvar);
    }
#undef LAUNCH
    CKL();
    return BEATGPU_OK;
}

template <bool WS>
int launch_stack(beatgpu_ctx* ctx, const WaveMap& w, const StackArgs& a)
{
    if (a.ns > 16384) return fail(ctx, BEATGPU_E_ARG, "nsamples %d too large", a.ns);
    if (w.store_dtype == BEATGPU_F32) {
        return (w.interp == BEATGPU_NEAREST) ? launch_stack_nvar<float, 1, WS>(ctx, a) : launch_stack_nvar<float, 4, WS>(ctx, a);
    } else {
        return (w.interp == BEATGPU_NEAREST) ? launch_stack_nvar<double, 1, WS>(ctx, a) : launch_stack_nvar<double, 4, WS>(ctx, a);
    }
}

template <typename T, int K>
int launch_chunk_nvar(beatgpu_ctx* ctx, const ChunkArgs& ca)
{
    const long n_items = (long)ca.s.nt * ca.nchunk * ca.s.B;
    const long grid = (n_items + kChunkWarps - 1) / kChunkWarps;
    if (grid > 2147483647L) return fail(ctx, BEATGPU_E_ARG, "grid too large: %ld", grid);
#define LAUNCH_CHUNK(NV)                                                                                            \
    do {                                                                                                            \
        if (ctx->chunk_occ >= 7) gf_stack_chunk_kernel<T, K, NV, 7><<<(unsigned)grid, kChunkWarps * 32, 0, ctx->stream>>>(ca);      \
        else if (ctx->chunk_occ == 6) gf_stack_chunk_kernel<T, K, NV, 6><<<(unsigned)grid, kChunkWarps * 32, 0, ctx->stream>>>(ca); \
        else gf_stack_chunk_kernel<T, K, NV, 5><<<(unsigned)grid, kChunkWarps * 32, 0, ctx->stream>>>(ca);                          \
    } while (0)
    switch (ca.s.nvar) {
        case 1: LAUNCH_CHUNK(1); break;
        case 2: LAUNCH_CHUNK(2); break;
        case 3: LAUNCH_CHUNK(3); break;
        default: return fail(ctx, BEATGPU_E_ARG, "n_slipvars must be 1..3, got %d", ca.s.nvar);
    }
#undef LAUNCH_CHUNK
    CKL();
    return BEATGPU_OK;
}

// plan of every (chain, patch) once per evaluation (see plan_cache_kernel): valid when start times do not depend on the target
template <typename T, int K>
int launch_plan_cache(beatgpu_ctx* ctx, ChunkArgs& ca)
{
    const StackArgs& a = ca.s;
    const long n = (long)a.B * a.np;
    size_t esz = 0;
    switch (a.nvar) {
        case 1: esz = sizeof(PatchPlan<T, K, 1>); break;
        case 2: esz = sizeof(PatchPlan<T, K, 2>); break;
        case 3: esz = sizeof(PatchPlan<T, K, 3>); break;
        default: return fail(ctx, BEATGPU_E_ARG, "n_slipvars must be 1..3, got %d", a.nvar);
    }
    const size_t need = (size_t)n * esz;
    if (ctx->plan_bytes < need) {
        if (ctx->d_plan) cudaFree(ctx->d_plan);
        ctx->d_plan = nullptr; ctx->plan_bytes = 0;
        CK(cudaMalloc(&ctx->d_plan, need));
        ctx->plan_bytes = need;
    }
    const unsigned grid = (unsigned)((n + 127) / 128);
    switch (a.nvar) {
        case 1: plan_cache_kernel<T, K, 1><<<grid, 128, 0, ctx->stream>>>(ca, (PatchPlan<T, K, 1>*)ctx->d_plan); break;
        case 2: plan_cache_kernel<T, K, 2><<<grid, 128, 0, ctx->stream>>>(ca, (PatchPlan<T, K, 2>*)ctx->d_plan); break;
        default: plan_cache_kernel<T, K, 3><<<grid, 128, 0, ctx->stream>>>(ca, (PatchPlan<T, K, 3>*)ctx->d_plan); break;
    }
    CKL();
    ca.plan_cache = ctx->d_plan;
    return BEATGPU_OK;
}

constexpr int kTimingRing = 1024;

bool stream_capturing(beatgpu_ctx* ctx)
{
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(ctx->stream, &st) != cudaSuccess) { (void)cudaGetLastError(); return false; }
    return st != cudaStreamCaptureStatusNone;
}

// event pair around the dominant kernels of one evaluation (skipped while the stream is being captured into a graph)
int timing_begin(beatgpu_ctx* ctx)
{
    if (stream_capturing(ctx)) { ctx->ev0 = ctx->ev1 = nullptr; return BEATGPU_OK; }
    if (ctx->tev.empty()) {
        ctx->tev.assign(2 * kTimingRing, nullptr);
        for (auto& e : ctx->tev) CK(cudaEventCreate(&e));
    }
    ctx->ev0 = ctx->tev[2 * ctx->tev_next];
    ctx->ev1 = ctx->tev[2 * ctx->tev_next + 1];
    CK(cudaEventRecord(ctx->ev0, ctx->stream));
    return BEATGPU_OK;
}

int timing_end(beatgpu_ctx* ctx)
{
    if (!ctx->ev1) return BEATGPU_OK;
    CK(cudaEventRecord(ctx->ev1, ctx->stream));
    ctx->tev_next = (ctx->tev_next + 1) % kTimingRing;
    ctx->tev_pending = std::min(ctx->tev_pending + 1, kTimingRing);
    ctx->ev_valid = true;
    return BEATGPU_OK;
}

// Patches per chunk of the chunked stacking pass.  All chains stream through one (target, chunk) slice of the library
// while it is L2 resident, so the slice -- chunk * nvar * ndur * nst * ld * sizeof(T) bytes -- must stay well below the
// L2 size: 40 % of cudaDeviceProp.l2CacheSize (the partial-synthetic stores and the other operands share the cache),
// at most kChunkMax (a lane plans up to two patches), and not so small that the partials scratch explodes (<= 24 chunks).
void derive_chunk(const beatgpu_ctx* ctx, const WaveMap& w, int nvar, int np, int* chunk, int* nchunk, int64_t* chunk_bytes)
{
    const int64_t esz = (w.store_dtype == BEATGPU_F32) ? 4 : 8;
    const int64_t per_patch = (int64_t)nvar * w.dims[2] * w.dims[3] * w.ld * esz;
    int target = ctx->chunk_patches;
    if (!ctx->chunk_forced) {
        const double budget = ctx->l2_frac * (double)ctx->prop.l2CacheSize;
        target = (int)std::max<int64_t>(1, std::min<int64_t>(kChunkMax, (int64_t)(budget / (double)std::max<int64_t>(1, per_patch))));
        const int min_chunk = (np + 23) / 24;                   // keep the [B, nt, nchunk, ns] scratch bounded
        target = std::max(target, std::min(kChunkMax, min_chunk));
    }
    const int nch0 = (np + target - 1) / target;
    *chunk = (np + nch0 - 1) / nch0;                            // balanced chunks
    *nchunk = (np + *chunk - 1) / *chunk;
    if (chunk_bytes) *chunk_bytes = (int64_t)(*chunk) * per_patch;
}

// chunked path: partial synthetics per (chain, target, patch chunk), then residual + misfit
// residual + misfit + logpt of every (chain, target): warp-per-item kernel for short traces with diagonal / banded
// weights (BEATGPU_MISFIT_WARP=0 disables it), CTA-per-item kernel otherwise
int launch_misfit(beatgpu_ctx* ctx, const MisfitArgs& m)
{
    const long n_items = (long)m.nt * m.B;
    if (ctx->misfit_warp && m.misfit_mode != MISFIT_DENSE && m.ns <= kMisfitWarpMaxNs) {
        const unsigned grid = (unsigned)((n_items + kStackWarps - 1) / kStackWarps);
        if (m.ns <= 128) misfit_warp_kernel<4><<<grid, kStackThreads, 0, ctx->stream>>>(m);
        else misfit_warp_kernel<8><<<grid, kStackThreads, 0, ctx->stream>>>(m);
        CKL();
        return BEATGPU_OK;
    }
    const size_t smem = (size_t)m.ns * sizeof(double);
    if (smem > 48 * 1024) CK(cudaFuncSetAttribute(misfit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    misfit_kernel<<<(unsigned)n_items, kStackThreads, smem, ctx->stream>>>(m);
    CKL();
    return BEATGPU_OK;
}

int launch_stack_chunked(beatgpu_ctx* ctx, const WaveMap& w, const StackArgs& a)
{
    ChunkArgs ca;
    ca.s = a;
    ca.zero_mask = 0u;
    derive_chunk(ctx, w, a.nvar, a.np, &ca.chunk, &ca.nchunk, nullptr);
    const size_t need = (size_t)a.B * a.nt * ca.nchunk * a.ns * sizeof(double);
    if (ctx->partial_bytes < need) {
        if (ctx->d_partial) cudaFree(ctx->d_partial);
        ctx->d_partial = nullptr; ctx->partial_bytes = 0;
        CK(cudaMalloc((void**)&ctx->d_partial, need));
        ctx->partial_bytes = need;
    }
    ca.partial = ctx->d_partial;
    int rc;
    const bool dense_gemm = a.misfit_mode == MISFIT_DENSE && ctx->geo_mode == 1 && a.B >= 32;
    if (dense_gemm) {       // scratch first: ensure_tmp may free + allocate (implicitly synchronising)
        const int mt = (a.ns + kGemmBM - 1) / kGemmBM;
        if ((rc = ensure_tmp(ctx, 4, (size_t)a.nt * a.B * a.ns * sizeof(double)))) return rc;
        if ((rc = ensure_tmp(ctx, 5, (size_t)a.nt * a.B * mt * sizeof(double)))) return rc;
    }
    ca.plan_cache = nullptr;
    if (ctx->plan_cache_on && a.st_st == 0 && a.corr == nullptr && a.nt > 1) {     // start times independent of the target
        if (w.store_dtype == BEATGPU_F32)
            rc = (w.interp == BEATGPU_NEAREST) ? launch_plan_cache<float, 1>(ctx, ca) : launch_plan_cache<float, 4>(ctx, ca);
        else
            rc = (w.interp == BEATGPU_NEAREST) ? launch_plan_cache<double, 1>(ctx, ca) : launch_plan_cache<double, 4>(ctx, ca);
        if (rc) return rc;
    }
    if (w.store_dtype == BEATGPU_F32)
        rc = (w.interp == BEATGPU_NEAREST) ? launch_chunk_nvar<float, 1>(ctx, ca) : launch_chunk_nvar<float, 4>(ctx, ca);
    else
        rc = (w.interp == BEATGPU_NEAREST) ? launch_chunk_nvar<double, 1>(ctx, ca) : launch_chunk_nvar<double, 4>(ctx, ca);
    if (rc) return rc;
    if (dense_gemm) {
        // full (non-Toeplitz) covariance: |U_t r|^2 for all chains is a GEMM per target -> FP64 tensor cores
        const int mt = (a.ns + kGemmBM - 1) / kGemmBM;
        double* R = (double*)ctx->d_tmp[4];
        double* qpart = (double*)ctx->d_tmp[5];
        residual_from_partials_kernel<<<(unsigned)((long)a.nt * a.B), 128, 0, ctx->stream>>>(ctx->d_partial, a.data, R, a.B, a.nt, a.ns, ca.nchunk);
        CKL();
        GemmArgs g;
        memset(&g, 0, sizeof(g));
        g.M = a.ns; g.N = a.B; g.K = a.ns; g.n_parts = 1;
        g.A[0] = a.W; g.a_sm = 1; g.a_sk = a.ns; g.a_batch = (long)a.ns * a.ns;        // U_t(m, k) = W[t][k*ns + m]
        g.B[0] = R; g.b_sk = 1; g.b_sn[0] = a.ns; g.b_batch = (long)a.B * a.ns;          // R_t(k, c) = R[t][c][k]
        g.upper = a.dense_upper;
        g.qpart = qpart; g.n_mtiles = mt; g.q_batch = (long)a.B * mt;
        CK(launch_dgemm<1>(g, a.nt, ctx->stream)); ctx->n_launches++;
        SeisFinishArgs f;
        memset(&f, 0, sizeof(f));
        f.B = a.B; f.nt = a.nt; f.n_mtiles = mt; f.qpart = qpart;
        f.slog_pdet = a.slog_pdet; f.nsamp = a.nsamp; f.hyper_idx = a.hyper_idx;
        f.hyp = a.hyp; f.hyp_sc = a.hyp_sc; f.chain_bad = a.chain_bad;
        f.logpts = a.logpts; f.logpts_sc = a.logpts_sc; f.out_ofs = a.out_ofs;
        seismic_finish_kernel<<<(unsigned)(((long)a.B * a.nt + 127) / 128), 128, 0, ctx->stream>>>(f);
        CKL();
        return BEATGPU_OK;
    }
    MisfitArgs m;
    memset(&m, 0, sizeof(m));
    m.B = a.B; m.nt = a.nt; m.ns = a.ns;
    m.resid = nullptr; m.partial = ctx->d_partial; m.nchunk = ca.nchunk; m.data = a.data; m.chain_bad = a.chain_bad;
    m.hyp = a.hyp; m.hyp_sc = a.hyp_sc; m.hyper_idx = a.hyper_idx;
    m.misfit_mode = a.misfit_mode; m.bw = a.bw; m.dense_upper = a.dense_upper;
    m.W = a.W; m.slog_pdet = a.slog_pdet; m.nsamp = a.nsamp;
    m.logpts = a.logpts; m.logpts_sc = a.logpts_sc; m.out_ofs = a.out_ofs;
    return launch_misfit(ctx, m);
}

void fill_static(const WaveMap& w, StackArgs& a, int nvar)
{
    a.nvar = nvar;
    for (int v = 0; v < BEATGPU_MAX_SLIPVARS; ++v) a.G[v] = w.G[v];
    a.nt = w.nt;
    a.np = (int)w.dims[1];
    a.ndur = (int)w.dims[2];
    a.nst = (int)w.dims[3];
    a.ns = w.ns;
    a.ld = w.ld;
    a.dur_min = w.dur_min; a.dur_step = w.dur_step; a.st_min = w.st_min; a.st_step = w.st_step;
    a.data = w.d_data;
    a.misfit_mode = w.misfit_mode; a.bw = w.bw; a.dense_upper = w.dense_upper;
    a.W = w.d_W; a.slog_pdet = w.d_slog_pdet; a.nsamp = w.d_nsamp;
    a.hyper_idx = w.d_hyper_idx;
    a.station_idx = w.d_station_idx;
}

int check_lib(beatgpu_ctx* ctx, const WaveMap& w, int nvar)
{
    for (int v = 0; v < nvar; ++v)
        if (!w.G[v]) return fail(ctx, BEATGPU_E_NOTREADY, "GF library for slip variable %d not uploaded", v);
    return BEATGPU_OK;
}

long row_stride_for(int ns, int store_dtype)
{
    long align_bytes = 16;
    if (const char* e = getenv("BEATGPU_ROW_ALIGN_BYTES")) { long v = atol(e); if (v >= 16 && (v % 16) == 0) align_bytes = v; }
    const long esz = (store_dtype == BEATGPU_F32) ? 4 : 8;
    const long bytes = ((long)ns * esz + align_bytes - 1) / align_bytes * align_bytes;
    return bytes / esz;
}

int set_lib_meta(beatgpu_ctx* ctx, WaveMap& w, int store_dtype, const int64_t dims[5], double dmin, double dstep,
                 double smin, double sstep)
{
    if (dims[0] != w.nt || dims[4] != w.ns)
        return fail(ctx, BEATGPU_E_ARG, "library dims (%ld targets, %ld samples) do not match wavemap (%d, %d)",
                    (long)dims[0], (long)dims[4], w.nt, w.ns);
    if (ctx->np_total && dims[1] != ctx->np_total)
        return fail(ctx, BEATGPU_E_ARG, "library has %ld patches, fault has %d", (long)dims[1], ctx->np_total);
    if (dims[0] * dims[1] * dims[2] * dims[3] > 2147483647LL)
        return fail(ctx, BEATGPU_E_ARG, "library has more than 2^31 rows");
    if (!(dstep > 0) || !(sstep > 0)) return fail(ctx, BEATGPU_E_ARG, "axis steps must be positive");
    if (w.axes_set) {
        bool same = w.store_dtype == store_dtype && w.dur_min == dmin && w.dur_step == dstep && w.st_min == smin && w.st_step == sstep;
        for (int i = 0; i < 5; ++i) same = same && (w.dims[i] == dims[i]);
        if (!same) return fail(ctx, BEATGPU_E_ARG, "all slip components of a wavemap must share dims, axes and storage dtype");
    }
    {   // the kernels address a row as a 32-bit offset in 16-byte units (PatchPlan::off): < 64 GB per slip component
        const long ld = row_stride_for(w.ns, store_dtype);
        const long long units = (long long)dims[0] * dims[1] * dims[2] * dims[3] * (ld * (store_dtype == BEATGPU_F32 ? 4 : 8) / 16);
        if (units > 0xFFFFFFFFLL)
            return fail(ctx, BEATGPU_E_ARG, "library of %.1f GB per slip component exceeds the 64 GB the row offsets can address",
                        (double)units * 16.0 / 1e9);
    }
    w.store_dtype = store_dtype;
    for (int i = 0; i < 5; ++i) w.dims[i] = dims[i];
    w.ld = row_stride_for(w.ns, store_dtype);
    w.dur_min = dmin; w.dur_step = dstep; w.st_min = smin; w.st_step = sstep;
    w.axes_set = true;
    return BEATGPU_OK;
}

}  // namespace

// =========================================================================================================
extern "C" {

int beatgpu_version(void) { return BEATGPU_VERSION; }

const char* beatgpu_last_error(const beatgpu_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int beatgpu_ctx_create(int device, beatgpu_ctx** out)
{
    if (!out) return fail(nullptr, BEATGPU_E_ARG, "out is NULL");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, BEATGPU_E_CUDA, "no CUDA device available (%s); libbeatgpu has no CPU fallback",
                    cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(nullptr, BEATGPU_E_ARG, "device %d out of range (0..%d)", device, ndev - 1);
    beatgpu_ctx* c = new beatgpu_ctx();
    c->device = device;
    if ((e = cudaSetDevice(device)) != cudaSuccess || (e = cudaGetDeviceProperties(&c->prop, device)) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaMalloc((void**)&c->d_viol, sizeof(unsigned long long))) != cudaSuccess ||
        (e = cudaMemset(c->d_viol, 0, sizeof(unsigned long long))) != cudaSuccess) {
        int rc = fail(nullptr, BEATGPU_E_CUDA, "context setup failed: %s", cudaGetErrorString(e));
        delete c;
        return rc;
    }
    if (c->prop.major != 10)
        fprintf(stderr, "libbeatgpu: warning: built for sm_100a, device is sm_%d%d\n", c->prop.major, c->prop.minor);
    for (int v = 0; v < BEATGPU_MAX_SLIPVARS; ++v) c->canon_slip[v] = 0;
    memset(&c->layout, 0, sizeof(c->layout));
    if (const char* e = getenv("BEATGPU_PERSISTENT")) c->persistent = atoi(e) != 0;
    if (const char* e = getenv("BEATGPU_STACK_MODE")) c->stack_mode = (strcmp(e, "fused") == 0 || strcmp(e, "0") == 0) ? 0 : 1;
    if (const char* e = getenv("BEATGPU_GEO_MODE")) c->geo_mode = (strcmp(e, "simple") == 0 || strcmp(e, "0") == 0) ? 0 : 1;
    if (const char* e = getenv("BEATGPU_FILTER_CAP")) { int v = atoi(e); if (v == 80 || v == 0) c->filter_cap = v; }
    if (const char* e = getenv("BEATGPU_GEOM_HALF")) { int v = atoi(e); if (v >= 2 && v <= 4) c->geom_half = v; }
    if (const char* e = getenv("BEATGPU_CHUNK")) { int v = atoi(e); if (v >= 1 && v <= kChunkMax) { c->chunk_patches = v; c->chunk_forced = true; } }
    if (const char* e = getenv("BEATGPU_CHUNK_OCC")) { int v = atoi(e); if (v >= 5 && v <= 7) c->chunk_occ = v; }
    if (const char* e = getenv("BEATGPU_L2_FRAC")) { double v = atof(e); if (v > 0.0 && v <= 4.0) c->l2_frac = v; }
    if (const char* e = getenv("BEATGPU_SPLIT_H2D")) c->split_h2d = atoi(e) != 0;
    if (const char* e = getenv("BEATGPU_SWEEP_PACK")) c->sweep_pack = atoi(e) != 0;
    if (const char* e = getenv("BEATGPU_MISFIT_WARP")) c->misfit_warp = atoi(e) != 0;
    if (const char* e = getenv("BEATGPU_PLAN_CACHE")) c->plan_cache_on = atoi(e) != 0;
    *out = c;
    return BEATGPU_OK;
}

void beatgpu_ctx_destroy(beatgpu_ctx* ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (auto& w : ctx->wmaps) {
        for (int v = 0; v < BEATGPU_MAX_SLIPVARS; ++v) if (w.G[v] && w.G_owned[v]) cudaFree(w.G[v]);
        cudaFree(w.d_station_idx); cudaFree(w.d_hyper_idx); cudaFree(w.d_nsamp); cudaFree(w.d_data); cudaFree(w.d_W);
        cudaFree(w.d_slog_pdet);
    }
    Geodetic& g = ctx->geo;
    for (int v = 0; v < BEATGPU_MAX_SLIPVARS; ++v) cudaFree(g.G[v]);
    cudaFree(g.d_data); cudaFree(g.d_odw); cudaFree(g.d_UT); cudaFree(g.d_slog); cudaFree(g.d_UT_ofs); cudaFree(g.d_lo);
    cudaFree(g.d_hi); cudaFree(g.d_upper); cudaFree(g.d_nsamp); cudaFree(g.d_hyper_idx);
    cudaFree(ctx->lap.d_LT);
    cudaFree(ctx->d_nd); cudaFree(ctx->d_ns); cudaFree(ctx->d_pofs); cudaFree(ctx->d_psize); cudaFree(ctx->d_fixed);
    cudaFree(ctx->d_q); cudaFree(ctx->d_logpts); cudaFree(ctx->d_like); cudaFree(ctx->d_t0); cudaFree(ctx->d_bad);
    cudaFree(ctx->d_viol);
    cudaFree(ctx->d_partial); cudaFree(ctx->d_plan);
    for (int i = 0; i < 6; ++i) cudaFree(ctx->d_tmp[i]);
    for (auto& st : ctx->gstores) { cudaFree(st.d_traces); cudaFree(st.d_itmin); cudaFree(st.d_nsamp); }
    for (auto& g : ctx->gwmaps) {
        cudaFree(g.d_rcv_lat); cudaFree(g.d_rcv_lon); cudaFree(g.d_rcv_itmin); cudaFree(g.d_rcv_nraw); cudaFree(g.d_rcv_first);
        cudaFree(g.d_tgt_of); cudaFree(g.d_tgt_f); cudaFree(g.d_tgt_nraw); cudaFree(g.d_tgt_ibeg); cudaFree(g.d_taper);
        cudaFree(g.d_rcv_arrival); cudaFree(g.d_tgt_arrival); cudaFree(g.d_rcv_station); cudaFree(g.d_tgt_station); cudaFree(g.d_tgt_rcv);
    }
    cudaFree(ctx->d_gfixed); cudaFree(ctx->d_peak_ratio); cudaFree(ctx->d_rplan); cudaFree(ctx->d_cplan); cudaFree(ctx->d_rawT); cudaFree(ctx->d_gmean);
    cudaFree(ctx->d_gerr);
    for (auto e : ctx->tev) if (e) cudaEventDestroy(e);
    if (ctx->copy_done) cudaEventDestroy(ctx->copy_done);
    if (ctx->copy_go) cudaEventDestroy(ctx->copy_go);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int beatgpu_sync(beatgpu_ctx* ctx)
{
    if (!ctx) return BEATGPU_E_ARG;
    CK(cudaStreamSynchronize(ctx->stream));
    return BEATGPU_OK;
}

int beatgpu_set_stream(beatgpu_ctx* ctx, void* cuda_stream, int external)
{
    if (!ctx) return BEATGPU_E_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    if (external) {
        // use the caller's stream handle as is; 0 is the (legacy) default stream, which is what torch's
        // current stream is unless the caller entered a torch.cuda.stream() context
        ctx->stream = (cudaStream_t)cuda_stream;
        ctx->own_stream = false;
    } else {
        CK(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
        ctx->own_stream = true;
    }
    return BEATGPU_OK;
}

int beatgpu_host_register(beatgpu_ctx* ctx, void* ptr, int64_t bytes)
{
    if (!ctx) return BEATGPU_E_ARG;
    if (!ptr || bytes <= 0) return fail(ctx, BEATGPU_E_ARG, "host_register: bad arguments");
    CK(cudaSetDevice(ctx->device));
    CK(cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterDefault));
    return BEATGPU_OK;
}

int beatgpu_host_unregister(beatgpu_ctx* ctx, void* ptr)
{
    if (!ctx) return BEATGPU_E_ARG;
    if (!ptr) return fail(ctx, BEATGPU_E_ARG, "host_unregister: NULL");
    CK(cudaHostUnregister(ptr));
    return BEATGPU_OK;
}

int beatgpu_device_info(beatgpu_ctx* ctx, int* n_sm, char* name, int name_len)
{
    if (!ctx) return BEATGPU_E_ARG;
    if (n_sm) *n_sm = ctx->prop.multiProcessorCount;
    if (name && name_len > 0) { strncpy(name, ctx->prop.name, name_len - 1); name[name_len - 1] = 0; }
    return BEATGPU_OK;
}

// ---------------------------------------------------------------------------------------------------------
int beatgpu_set_fault(beatgpu_ctx* ctx, int nsf, const int32_t* nd, const int32_t* ns, const double* psize)
{
    if (!ctx) return BEATGPU_E_ARG;
    if (nsf <= 0 || !nd || !ns || !psize) return fail(ctx, BEATGPU_E_ARG, "set_fault: bad arguments");
    CK(cudaSetDevice(ctx->device));
    ctx->h_nd.assign(nd, nd + nsf);
    ctx->h_ns.assign(ns, ns + nsf);
    ctx->h_psize.assign(psize, psize + nsf);
    ctx->h_pofs.resize(nsf);
    int tot = 0, mx = 0, mxd = 0;
    for (int i = 0; i < nsf; ++i) {
        if (nd[i] <= 0 || ns[i] <= 0 || !(psize[i] > 0)) return fail(ctx, BEATGPU_E_ARG, "set_fault: subfault %d has an empty grid", i);
        ctx->h_pofs[i] = tot;
        tot += nd[i] * ns[i];
        mx = std::max(mx, nd[i] * ns[i]);
        mxd = std::max(mxd, std::min(nd[i], ns[i]));
    }
    ctx->nsf = nsf; ctx->np_total = tot; ctx->max_np_sf = mx; ctx->max_diag = mxd;
    int rc;
    if ((rc = upload_vec(ctx, &ctx->d_nd, ctx->h_nd.data(), nsf))) return rc;
    if ((rc = upload_vec(ctx, &ctx->d_ns, ctx->h_ns.data(), nsf))) return rc;
    if ((rc = upload_vec(ctx, &ctx->d_pofs, ctx->h_pofs.data(), nsf))) return rc;
    if ((rc = upload_vec(ctx, &ctx->d_psize, ctx->h_psize.data(), nsf))) return rc;
    ctx->layout_set = false;
    return BEATGPU_OK;
}

int beatgpu_set_layout(beatgpu_ctx* ctx, const beatgpu_layout* L, const double* fixed)
{
    if (!ctx || !L) return BEATGPU_E_ARG;
    if (!ctx->nsf) return fail(ctx, BEATGPU_E_NOTREADY, "set_layout: call set_fault first");
    if (L->n_slipvars < 1 || L->n_slipvars > BEATGPU_MAX_SLIPVARS) return fail(ctx, BEATGPU_E_ARG, "set_layout: n_slipvars %d", L->n_slipvars);
    if (L->n_params <= 0 || L->n_hypers < 0 || L->n_time_shifts < 0) return fail(ctx, BEATGPU_E_ARG, "set_layout: bad sizes");
    CK(cudaSetDevice(ctx->device));
    const int np = ctx->np_total, nsf = ctx->nsf;
    long o = 0;
    for (int v = 0; v < L->n_slipvars; ++v) { ctx->canon_slip[v] = o; o += np; }
    ctx->canon_dur = o; o += np;
    ctx->canon_vel = o; o += np;
    ctx->canon_nstr = o; o += nsf;
    ctx->canon_ndip = o; o += nsf;
    ctx->canon_time = o; o += nsf;
    ctx->canon_hyp = o; o += L->n_hypers;
    ctx->canon_ts = o; o += L->n_time_shifts;
    ctx->canon_len = o;
    struct { int off; int len; const char* name; } vars[16];
    int nv = 0;
    for (int v = 0; v < L->n_slipvars; ++v) vars[nv++] = {L->off_slip[v], np, "slip"};
    vars[nv++] = {L->off_durations, np, "durations"};
    vars[nv++] = {L->off_velocities, np, "velocities"};
    vars[nv++] = {L->off_nucleation_strike, nsf, "nucleation_strike"};
    vars[nv++] = {L->off_nucleation_dip, nsf, "nucleation_dip"};
    vars[nv++] = {L->off_time, nsf, "time"};
    if (L->n_hypers) vars[nv++] = {L->off_hypers, L->n_hypers, "hypers"};
    if (L->n_time_shifts) vars[nv++] = {L->off_time_shifts, L->n_time_shifts, "time_shifts"};
    bool any_fixed = false;
    for (int i = 0; i < nv; ++i) {
        if (vars[i].off < 0) { any_fixed = true; continue; }
        if (vars[i].off + vars[i].len > L->n_params)
            return fail(ctx, BEATGPU_E_ARG, "set_layout: %s [%d, %d) exceeds n_params %d", vars[i].name, vars[i].off,
                        vars[i].off + vars[i].len, L->n_params);
    }
    if (any_fixed && !fixed) return fail(ctx, BEATGPU_E_ARG, "set_layout: variables with offset -1 need the `fixed` vector");
    std::vector<double> fx((size_t)std::max<long>(1, ctx->canon_len), 0.0);
    if (fixed) std::copy(fixed, fixed + ctx->canon_len, fx.begin());
    int rc = upload_vec(ctx, &ctx->d_fixed, fx.data(), fx.size());
    if (rc) return rc;
    ctx->layout = *L;
    ctx->layout_set = true;
    return BEATGPU_OK;
}

int beatgpu_add_wavemap(beatgpu_ctx* ctx, int nt, int ns, int interp, const int32_t* station_idx, const int32_t* hyper_idx,
                        const int32_t* nsamples, int* wmap_id)
{
    if (!ctx) return BEATGPU_E_ARG;
    if (nt <= 0 || ns <= 0 || !hyper_idx || !nsamples || !wmap_id) return fail(ctx, BEATGPU_E_ARG, "add_wavemap: bad arguments");
    if (interp != BEATGPU_NEAREST && interp != BEATGPU_MULTILINEAR)
        return fail(ctx, BEATGPU_E_ARG, "add_wavemap: interpolation scheme %d not implemented", interp);
    if (!ctx->gwmaps.empty()) return fail(ctx, BEATGPU_E_ARG, "add_wavemap: the context holds geometry-mode wavemaps; use one context per mode");
    CK(cudaSetDevice(ctx->device));
    WaveMap w;
    w.nt = nt; w.ns = ns; w.interp = interp;
    int rc;
    if (station_idx) { w.has_station = true; if ((rc = upload_vec(ctx, &w.d_station_idx, station_idx, nt))) return rc; }
    if ((rc = upload_vec(ctx, &w.d_hyper_idx, hyper_idx, nt))) return rc;
    if ((rc = upload_vec(ctx, &w.d_nsamp, nsamples, nt))) return rc;
    ctx->wmaps.push_back(w);
    *wmap_id = (int)ctx->wmaps.size() - 1;
    assign_out_offsets(ctx);
    return BEATGPU_OK;
}

#define GET_WMAP(id)                                                                                   \
    if (!ctx) return BEATGPU_E_ARG;                                                                    \
    if ((id) < 0 || (id) >= (int)ctx->wmaps.size()) return fail(ctx, BEATGPU_E_ARG, "unknown wavemap id %d", (id)); \
    WaveMap& w = ctx->wmaps[(id)];                                                                     \
    CK(cudaSetDevice(ctx->device))

int beatgpu_alloc_gflib(beatgpu_ctx* ctx, int wmap_id, int var, int store_dtype, const int64_t dims[5], double dmin,
                        double dstep, double smin, double sstep, void** device_ptr, int64_t* row_stride)
{
    GET_WMAP(wmap_id);
    if (var < 0 || var >= BEATGPU_MAX_SLIPVARS) return fail(ctx, BEATGPU_E_ARG, "slip variable index %d", var);
    if (store_dtype != BEATGPU_F32 && store_dtype != BEATGPU_F64) return fail(ctx, BEATGPU_E_ARG, "store dtype %d", store_dtype);
    int rc = set_lib_meta(ctx, w, store_dtype, dims, dmin, dstep, smin, sstep);
    if (rc) return rc;
    if (w.G[var] && w.G_owned[var]) { cudaFree(w.G[var]); w.G[var] = nullptr; }
    const size_t esz = store_dtype == BEATGPU_F32 ? 4 : 8;
    const size_t rows = (size_t)dims[0] * dims[1] * dims[2] * dims[3];
    CK(cudaMalloc(&w.G[var], rows * w.ld * esz));
    w.G_owned[var] = true;
    if (device_ptr) *device_ptr = w.G[var];
    if (row_stride) *row_stride = w.ld;
    return BEATGPU_OK;
}

int beatgpu_upload_gflib(beatgpu_ctx* ctx, int wmap_id, int var, const void* traces, int src_dtype, int store_dtype,
                         const int64_t dims[5], double dmin, double dstep, double smin, double sstep)
{
    if (!ctx) return BEATGPU_E_ARG;
    if (!traces) return fail(ctx, BEATGPU_E_ARG, "upload_gflib: traces is NULL");
    if (src_dtype != BEATGPU_F32 && src_dtype != BEATGPU_F64) return fail(ctx, BEATGPU_E_ARG, "source dtype %d", src_dtype);
    void* dptr = nullptr;
    int64_t ld = 0;
    int rc = beatgpu_alloc_gflib(ctx, wmap_id, var, store_dtype, dims, dmin, dstep, smin, sstep, &dptr, &ld);
    if (rc) return rc;
    const int ns = (int)dims[4];
    const size_t rows = (size_t)dims[0] * dims[1] * dims[2] * dims[3];
    const size_t ssz = src_dtype == BEATGPU_F32 ? 4 : 8, dsz = store_dtype == BEATGPU_F32 ? 4 : 8;
    // Stream the (possibly memory-mapped, pageable) source in 64 MiB chunks through two pinned host buffers and two
    // device staging buffers: host threads copy chunk i+1 into pinned memory while the DMA engine moves chunk i and the
    // repack kernel (dtype conversion + row padding) of chunk i runs behind it.  A source the caller has already
    // page-locked (beatgpu_host_register) is copied from directly.
    const size_t row_bytes = (size_t)ns * ssz;
    size_t chunk_target = (size_t)64 << 20;
    if (const char* e = getenv("BEATGPU_UPLOAD_CHUNK_KB")) { long v = atol(e); if (v >= 1) chunk_target = (size_t)v << 10; }   // tests: force many chunks
    const size_t chunk_rows = std::max<size_t>(1, chunk_target / row_bytes);
    const size_t chunk_bytes = std::min(rows, chunk_rows) * row_bytes;
    if ((rc = ensure_tmp(ctx, 0, chunk_bytes)) || (rc = ensure_tmp(ctx, 1, chunk_bytes))) return rc;
    cudaPointerAttributes attr;
    bool src_pinned = cudaPointerGetAttributes(&attr, traces) == cudaSuccess && attr.type == cudaMemoryTypeHost;
    (void)cudaGetLastError();
    if (getenv("BEATGPU_UPLOAD_DIRECT")) src_pinned = true;       // measurement knob: plain cudaMemcpyAsync from the pageable source
    void* pin[2] = {nullptr, nullptr};
    cudaEvent_t ev[2] = {nullptr, nullptr};
    auto cleanup = [&]() { for (int b = 0; b < 2; ++b) { if (pin[b]) cudaFreeHost(pin[b]); if (ev[b]) cudaEventDestroy(ev[b]); } };
    if (!src_pinned) {
        for (int b = 0; b < 2; ++b) {
            if (cudaHostAlloc(&pin[b], chunk_bytes, cudaHostAllocDefault) != cudaSuccess) { (void)cudaGetLastError(); cleanup(); pin[0] = pin[1] = nullptr; src_pinned = true; break; }
        }
    }
    for (int b = 0; b < 2; ++b) if (cudaEventCreateWithFlags(&ev[b], cudaEventDisableTiming) != cudaSuccess) { cleanup(); return fail(ctx, BEATGPU_E_CUDA, "upload_gflib: cudaEventCreate failed"); }
    const unsigned n_thr = std::max(1u, std::min(8u, std::thread::hardware_concurrency() / 2));
    size_t i = 0;
    for (size_t r0 = 0; r0 < rows; r0 += chunk_rows, ++i) {
        const int b = (int)(i & 1);
        const size_t nr = std::min(chunk_rows, rows - r0), nbytes = nr * row_bytes;
        const char* src = (const char*)traces + r0 * row_bytes;
        if (i >= 2 && cudaEventSynchronize(ev[b]) != cudaSuccess) { cleanup(); return fail(ctx, BEATGPU_E_CUDA, "upload_gflib: staging buffer wait failed"); }
        const void* h_src = src;
        if (pin[b]) {                                              // pageable source: parallel host copy into pinned memory
            std::vector<std::thread> pool;
            const size_t slice = (nbytes + n_thr - 1) / n_thr;
            for (unsigned t = 0; t < n_thr; ++t) {
                const size_t o = (size_t)t * slice;
                if (o >= nbytes) break;
                pool.emplace_back([=]() { memcpy((char*)pin[b] + o, src + o, std::min(slice, nbytes - o)); });
            }
            for (auto& th : pool) th.join();
            h_src = pin[b];
        }
        cudaError_t e = cudaMemcpyAsync(ctx->d_tmp[b], h_src, nbytes, cudaMemcpyHostToDevice, ctx->stream);
        if (e != cudaSuccess) { (void)cudaGetLastError(); cleanup(); return fail(ctx, BEATGPU_E_CUDA, "upload_gflib: H2D copy failed: %s", cudaGetErrorString(e)); }
        char* dst = (char*)dptr + r0 * ld * dsz;
        const int blocks = (int)std::min<size_t>(148 * 16, (nr * ld + 255) / 256);
        if (src_dtype == BEATGPU_F64 && store_dtype == BEATGPU_F32)
            repack_rows_kernel<double, float><<<blocks, 256, 0, ctx->stream>>>((const double*)ctx->d_tmp[b], (float*)dst, (long)nr, ns, (long)ld);
        else if (src_dtype == BEATGPU_F64 && store_dtype == BEATGPU_F64)
            repack_rows_kernel<double, double><<<blocks, 256, 0, ctx->stream>>>((const double*)ctx->d_tmp[b], (double*)dst, (long)nr, ns, (long)ld);
        else if (src_dtype == BEATGPU_F32 && store_dtype == BEATGPU_F32)
            repack_rows_kernel<float, float><<<blocks, 256, 0, ctx->stream>>>((const float*)ctx->d_tmp[b], (float*)dst, (long)nr, ns, (long)ld);
        else
            repack_rows_kernel<float, double><<<blocks, 256, 0, ctx->stream>>>((const float*)ctx->d_tmp[b], (double*)dst, (long)nr, ns, (long)ld);
        e = cudaGetLastError();
        if (e != cudaSuccess) { cleanup(); return fail(ctx, BEATGPU_E_CUDA, "upload_gflib: repack launch failed: %s", cudaGetErrorString(e)); }
        ctx->n_launches++;
        cudaEventRecord(ev[b], ctx->stream);
    }
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    cleanup();
    if (e != cudaSuccess) { (void)cudaGetLastError(); return fail(ctx, BEATGPU_E_CUDA, "upload_gflib: %s", cudaGetErrorString(e)); }
    return BEATGPU_OK;
}

int beatgpu_upload_data(beatgpu_ctx* ctx, int wmap_id, const double* data)
{
    GET_WMAP(wmap_id);
    if (!data) return fail(ctx, BEATGPU_E_ARG, "upload_data: NULL");
    return upload_vec(ctx, &w.d_data, data, (size_t)w.nt * w.ns);
}

int beatgpu_update_weights(beatgpu_ctx* ctx, int wmap_id, const double* U, const double* slog_pdet, double band_rtol)
{
    GET_WMAP(wmap_id);
    if (!U || !slog_pdet) return fail(ctx, BEATGPU_E_ARG, "update_weights: NULL");
    if (band_rtol < 0) band_rtol = 1e-13;
    const int nt = w.nt, ns = w.ns;
    // structure detection over all targets of the wavemap
    bool lower = false;
    int bw = 0;
    for (int t = 0; t < nt; ++t) {
        const double* Ut = U + (size_t)t * ns * ns;
        double mx = 0.0;
        for (size_t i = 0; i < (size_t)ns * ns; ++i) {
            const double v = std::fabs(Ut[i]);
            if (!(v == v)) return fail(ctx, BEATGPU_E_ARG, "update_weights: NaN in weight matrix of target %d", t);
            mx = std::max(mx, v);
        }
        const double thr = band_rtol * mx;
        for (int i = 0; i < ns; ++i) {
            const double* row = Ut + (size_t)i * ns;
            for (int j = 0; j < i && !lower; ++j) if (std::fabs(row[j]) > thr) lower = true;
            for (int j = ns - 1; j > i + bw; --j) if (std::fabs(row[j]) > thr) { bw = j - i; break; }
        }
    }
    int mode;
    if (lower) mode = MISFIT_DENSE;
    else if (bw == 0) mode = MISFIT_DIAG;
    else if (bw <= 32 && bw + 1 < ns / 2) mode = MISFIT_BAND;
    else mode = MISFIT_DENSE;
    std::vector<double> W;
    if (mode == MISFIT_DIAG) {
        W.resize((size_t)nt * ns);
        for (int t = 0; t < nt; ++t) for (int k = 0; k < ns; ++k) W[(size_t)t * ns + k] = U[(size_t)t * ns * ns + (size_t)k * ns + k];
    } else if (mode == MISFIT_BAND) {
        W.assign((size_t)nt * (bw + 1) * ns, 0.0);
        for (int t = 0; t < nt; ++t)
            for (int j = 0; j <= bw; ++j)
                for (int k = 0; k + j < ns; ++k)
                    W[((size_t)t * (bw + 1) + j) * ns + k] = U[(size_t)t * ns * ns + (size_t)k * ns + (k + j)];
    } else {
        W.resize((size_t)nt * ns * ns);
        for (int t = 0; t < nt; ++t)
            for (int k = 0; k < ns; ++k)
                for (int j = 0; j < ns; ++j)
                    W[(size_t)t * ns * ns + (size_t)j * ns + k] = U[(size_t)t * ns * ns + (size_t)k * ns + j];
    }
    // in-place update when the footprint is unchanged (between SMC stages): no reallocation
    const size_t bytes = W.size() * sizeof(double);
    if (w.d_W && w.W_bytes == bytes) {
        CK(cudaMemcpyAsync(w.d_W, W.data(), bytes, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    } else {
        int rc = upload_vec(ctx, &w.d_W, W.data(), W.size());
        if (rc) return rc;
        w.W_bytes = bytes;
    }
    int rc = upload_vec(ctx, &w.d_slog_pdet, slog_pdet, nt);
    if (rc) return rc;
    w.misfit_mode = mode; w.bw = bw; w.dense_upper = lower ? 0 : 1;
    return BEATGPU_OK;
}

int beatgpu_update_weights_dev(beatgpu_ctx* ctx, int wmap_id, const double* U_dev, const double* slog_pdet_dev, double band_rtol)
{
    GET_WMAP(wmap_id);
    if (!U_dev || !slog_pdet_dev) return fail(ctx, BEATGPU_E_ARG, "update_weights_dev: NULL");
    if (band_rtol < 0) band_rtol = 1e-13;
    const int nt = w.nt, ns = w.ns;
    int rc;
    if ((rc = ensure_tmp(ctx, 4, (size_t)nt * sizeof(double) + 4 * sizeof(int)))) return rc;
    double* d_amax = (double*)ctx->d_tmp[4];
    int* d_flags = (int*)(d_amax + nt);                               // [has_nan, lower, bw]
    CK(cudaMemsetAsync(d_flags, 0, 4 * sizeof(int), ctx->stream));
    weights_absmax_kernel<<<nt, 256, 0, ctx->stream>>>(U_dev, ns, d_amax, d_flags);
    CKL();
    weights_structure_kernel<<<nt, 256, 0, ctx->stream>>>(U_dev, ns, d_amax, band_rtol, d_flags + 1, d_flags + 2);
    CKL();
    int flags[4] = {0, 0, 0, 0};
    CK(cudaMemcpyAsync(flags, d_flags, sizeof(flags), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (flags[0]) return fail(ctx, BEATGPU_E_ARG, "update_weights_dev: NaN in a weight matrix");
    const bool lower = flags[1] != 0;
    const int bw = flags[2];
    int mode;                                                         // same rules as the host path
    if (lower) mode = MISFIT_DENSE;
    else if (bw == 0) mode = MISFIT_DIAG;
    else if (bw <= 32 && bw + 1 < ns / 2) mode = MISFIT_BAND;
    else mode = MISFIT_DENSE;
    const size_t n = mode == MISFIT_DIAG ? (size_t)nt * ns : (mode == MISFIT_BAND ? (size_t)nt * (bw + 1) * ns : (size_t)nt * ns * ns);
    const size_t bytes = n * sizeof(double);
    if (!w.d_W || w.W_bytes != bytes) {                               // footprint changed: reallocate (stage boundary only)
        if (w.d_W) { cudaFree(w.d_W); w.d_W = nullptr; w.W_bytes = 0; }
        CK(cudaMalloc((void**)&w.d_W, bytes));
        w.W_bytes = bytes;
    }
    const int blocks = (int)std::min<size_t>((size_t)ctx->prop.multiProcessorCount * 8, (n + 255) / 256);
    weights_repack_kernel<<<blocks, 256, 0, ctx->stream>>>(U_dev, w.d_W, nt, ns, mode == MISFIT_DIAG ? 0 : (mode == MISFIT_BAND ? 1 : 2), bw);
    CKL();
    if (!w.d_slog_pdet) CK(cudaMalloc((void**)&w.d_slog_pdet, (size_t)nt * sizeof(double)));
    CK(cudaMemcpyAsync(w.d_slog_pdet, slog_pdet_dev, (size_t)nt * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));                           // the caller may free U_dev right after the call
    w.misfit_mode = mode; w.bw = bw; w.dense_upper = lower ? 0 : 1;
    return BEATGPU_OK;
}

int beatgpu_set_geodetic(beatgpu_ctx* ctx, int nobs, int nds, const int32_t* slo, const int32_t* shi, const double* const* G,
                         const double* data, const double* odw, const double* U_concat, const double* slog_pdet,
                         const int32_t* nsamples, const int32_t* hyper_idx)
{
    if (!ctx) return BEATGPU_E_ARG;
    if (!ctx->layout_set) return fail(ctx, BEATGPU_E_NOTREADY, "set_geodetic: call set_fault and set_layout first");
    if (nobs <= 0 || nds <= 0 || !slo || !shi || !G || !data || !odw || !U_concat || !slog_pdet || !nsamples || !hyper_idx)
        return fail(ctx, BEATGPU_E_ARG, "set_geodetic: bad arguments");
    CK(cudaSetDevice(ctx->device));
    Geodetic& g = ctx->geo;
    g.nobs = nobs; g.nds = nds;
    g.lo.assign(slo, slo + nds); g.hi.assign(shi, shi + nds);
    g.ut_ofs.resize(nds);
    long tot = 0; int mx = 0;
    for (int d = 0; d < nds; ++d) {
        const int n = shi[d] - slo[d];
        if (n <= 0 || slo[d] < 0 || shi[d] > nobs) return fail(ctx, BEATGPU_E_ARG, "set_geodetic: bad slice %d", d);
        g.ut_ofs[d] = tot; tot += (long)n * n; mx = std::max(mx, n);
    }
    g.ut_total = tot; g.max_n = mx;
    int rc;
    for (int v = 0; v < ctx->layout.n_slipvars; ++v) {
        if (!G[v]) return fail(ctx, BEATGPU_E_ARG, "set_geodetic: G[%d] is NULL", v);
        if ((rc = upload_vec(ctx, &g.G[v], G[v], (size_t)ctx->np_total * nobs))) return rc;
    }
    if ((rc = upload_vec(ctx, &g.d_data, data, nobs))) return rc;
    if ((rc = upload_vec(ctx, &g.d_odw, odw, nobs))) return rc;
    if ((rc = upload_vec(ctx, &g.d_lo, g.lo.data(), nds))) return rc;
    if ((rc = upload_vec(ctx, &g.d_hi, g.hi.data(), nds))) return rc;
    if ((rc = upload_vec(ctx, &g.d_UT_ofs, g.ut_ofs.data(), nds))) return rc;
    if ((rc = upload_vec(ctx, &g.d_nsamp, nsamples, nds))) return rc;
    if ((rc = upload_vec(ctx, &g.d_hyper_idx, hyper_idx, nds))) return rc;
    g.h_nsamp.assign(nsamples, nsamples + nds);
    g.h_hyper_idx.assign(hyper_idx, hyper_idx + nds);
    for (int d = 0; d < nds; ++d)
        if (hyper_idx[d] < 0 || hyper_idx[d] >= ctx->layout.n_hypers) return fail(ctx, BEATGPU_E_ARG, "set_geodetic: hyper_idx[%d] out of range", d);
    g.set = true;
    assign_out_offsets(ctx);
    return beatgpu_update_geodetic_weights(ctx, U_concat, slog_pdet);
}

int beatgpu_update_geodetic_weights(beatgpu_ctx* ctx, const double* U_concat, const double* slog_pdet)
{
    if (!ctx) return BEATGPU_E_ARG;
    Geodetic& g = ctx->geo;
    if (!g.set) return fail(ctx, BEATGPU_E_NOTREADY, "update_geodetic_weights: geodetic composite not set");
    if (!U_concat || !slog_pdet) return fail(ctx, BEATGPU_E_ARG, "update_geodetic_weights: NULL");
    CK(cudaSetDevice(ctx->device));
    std::vector<double> UT((size_t)g.ut_total);
    std::vector<int> upper(g.nds, 1);
    for (int d = 0; d < g.nds; ++d) {
        const int n = g.hi[d] - g.lo[d];
        const double* U = U_concat + g.ut_ofs[d];
        double* T = UT.data() + g.ut_ofs[d];
        for (int k = 0; k < n; ++k)
            for (int j = 0; j < n; ++j) {
                const double v = U[(size_t)k * n + j];
                T[(size_t)j * n + k] = v;
                if (j < k && v != 0.0) upper[d] = 0;
            }
    }
    int rc;
    if ((rc = upload_vec(ctx, &g.d_UT, UT.data(), UT.size()))) return rc;
    if ((rc = upload_vec(ctx, &g.d_upper, upper.data(), g.nds))) return rc;
    if ((rc = upload_vec(ctx, &g.d_slog, slog_pdet, g.nds))) return rc;
    g.h_upper = upper;
    g.h_slog.assign(slog_pdet, slog_pdet + g.nds);
    return BEATGPU_OK;
}

int beatgpu_set_laplacian(beatgpu_ctx* ctx, const double* L, double sdet, int hyper_idx)
{
    if (!ctx) return BEATGPU_E_ARG;
    if (!ctx->layout_set) return fail(ctx, BEATGPU_E_NOTREADY, "set_laplacian: call set_fault and set_layout first");
    if (!L || hyper_idx < 0 || hyper_idx >= ctx->layout.n_hypers) return fail(ctx, BEATGPU_E_ARG, "set_laplacian: bad arguments");
    CK(cudaSetDevice(ctx->device));
    const int np = ctx->np_total;
    std::vector<double> LT((size_t)np * np);
    for (int i = 0; i < np; ++i) for (int j = 0; j < np; ++j) LT[(size_t)j * np + i] = L[(size_t)i * np + j];
    int rc = upload_vec(ctx, &ctx->lap.d_LT, LT.data(), LT.size());
    if (rc) return rc;
    ctx->lap.sdet = sdet; ctx->lap.hyper_idx = hyper_idx; ctx->lap.set = true;
    assign_out_offsets(ctx);
    return BEATGPU_OK;
}

int beatgpu_n_outputs(beatgpu_ctx* ctx, int* n_out)
{
    if (!ctx || !n_out) return BEATGPU_E_ARG;
    *n_out = n_outputs(ctx);
    return BEATGPU_OK;
}

// ---------------------------------------------------------------------------------------------------------
int beatgpu_fast_sweep_batch(beatgpu_ctx* ctx, int sf, int B, const double* slowness, const int32_t* nuc_dip_idx,
                             const int32_t* nuc_strike_idx, double* starttimes, int32_t* n_iter)
{
    if (!ctx) return BEATGPU_E_ARG;
    if (!ctx->nsf) return fail(ctx, BEATGPU_E_NOTREADY, "fast_sweep_batch: call set_fault first");
    if (sf < 0 || sf >= ctx->nsf || B <= 0 || !slowness || !nuc_dip_idx || !nuc_strike_idx || !starttimes)
        return fail(ctx, BEATGPU_E_ARG, "fast_sweep_batch: bad arguments");
    CK(cudaSetDevice(ctx->device));
    const int n = ctx->h_nd[sf] * ctx->h_ns[sf];
    // the reference wrapper raises on bad input rather than reading outside the grid
    for (int b = 0; b < B; ++b)
        if (nuc_dip_idx[b] < 0 || nuc_dip_idx[b] >= ctx->h_nd[sf] || nuc_strike_idx[b] < 0 || nuc_strike_idx[b] >= ctx->h_ns[sf])
            return fail(ctx, BEATGPU_E_INDEX, "fast_sweep_batch: nucleation index (%d, %d) of chain %d outside the %dx%d grid",
                        nuc_dip_idx[b], nuc_strike_idx[b], b, ctx->h_nd[sf], ctx->h_ns[sf]);
    int rc;
    const size_t nb = (size_t)B * n * sizeof(double);
    if ((rc = ensure_tmp(ctx, 0, nb))) return rc;
    if ((rc = ensure_tmp(ctx, 1, nb))) return rc;
    if ((rc = ensure_tmp(ctx, 2, (size_t)B * 3 * sizeof(int)))) return rc;
    int* d_idx = (int*)ctx->d_tmp[2];
    CK(cudaMemcpyAsync(ctx->d_tmp[0], slowness, nb, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(d_idx, nuc_dip_idx, (size_t)B * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(d_idx + B, nuc_strike_idx, (size_t)B * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    SweepArgs a;
    memset(&a, 0, sizeof(a));
    a.n_dip = ctx->d_nd; a.n_strike = ctx->d_ns; a.patch_ofs = ctx->d_pofs; a.patch_size = ctx->d_psize;
    a.n_subfaults = ctx->nsf; a.n_patches_total = ctx->np_total; a.max_np_sf = ctx->max_np_sf; a.B = B;
    a.vel = (const double*)ctx->d_tmp[0]; a.vel_stride = n; a.is_slowness = 1;
    a.nuc_dip_idx = d_idx; a.nuc_strike_idx = d_idx + B; a.only_sf = sf;
    a.time = nullptr; a.t0 = (double*)ctx->d_tmp[1]; a.n_iter = n_iter ? d_idx + 2 * B : nullptr;
    a.violations = ctx->d_viol; a.chain_bad = nullptr;
    if ((rc = launch_sweep(ctx, a, B))) return rc;
    CK(cudaMemcpyAsync(starttimes, ctx->d_tmp[1], nb, cudaMemcpyDeviceToHost, ctx->stream));
    if (n_iter) CK(cudaMemcpyAsync(n_iter, d_idx + 2 * B, (size_t)B * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return BEATGPU_OK;
}

static int check_violations(beatgpu_ctx* ctx, const char* what)
{
    unsigned long long v = 0;
    CK(cudaMemcpyAsync(&v, ctx->d_viol, sizeof(v), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (v) {
        CK(cudaMemsetAsync(ctx->d_viol, 0, sizeof(v), ctx->stream));
        return fail(ctx, BEATGPU_E_INDEX, "%s: %llu GF-library / patch-grid indices out of range (affected outputs are NaN)", what, v);
    }
    return BEATGPU_OK;
}

int beatgpu_stack_batch(beatgpu_ctx* ctx, int wmap_id, int B, int nvar, const double* durations, const double* starttimes,
                        const double* slips, double* synthetics)
{
    GET_WMAP(wmap_id);
    if (B <= 0 || !durations || !starttimes || !slips || !synthetics) return fail(ctx, BEATGPU_E_ARG, "stack_batch: bad arguments");
    int rc = check_lib(ctx, w, nvar);
    if (rc) return rc;
    const int np = (int)w.dims[1];
    const size_t b_dur = (size_t)B * np * 8, b_st = (size_t)B * w.nt * np * 8, b_sl = (size_t)nvar * B * np * 8,
                 b_out = (size_t)B * w.nt * w.ns * 8;
    if ((rc = ensure_tmp(ctx, 0, b_dur)) || (rc = ensure_tmp(ctx, 1, b_st)) || (rc = ensure_tmp(ctx, 2, b_sl)) ||
        (rc = ensure_tmp(ctx, 3, b_out)))
        return rc;
    CK(cudaMemsetAsync(ctx->d_viol, 0, sizeof(unsigned long long), ctx->stream));      // violations are reported per call
    CK(cudaMemcpyAsync(ctx->d_tmp[0], durations, b_dur, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_tmp[1], starttimes, b_st, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_tmp[2], slips, b_sl, cudaMemcpyHostToDevice, ctx->stream));
    StackArgs a;
    memset(&a, 0, sizeof(a));
    fill_static(w, a, nvar);
    a.B = B;
    a.dur = (const double*)ctx->d_tmp[0]; a.dur_sc = np;
    for (int v = 0; v < nvar; ++v) { a.slip[v] = (const double*)ctx->d_tmp[2] + (size_t)v * B * np; a.slip_sc[v] = np; }
    a.st = (const double*)ctx->d_tmp[1]; a.st_sc = (long)w.nt * np; a.st_st = np;
    a.corr = nullptr; a.station_idx = nullptr;
    a.synth = (double*)ctx->d_tmp[3];
    a.violations = ctx->d_viol;
    if ((rc = launch_stack<true>(ctx, w, a))) return rc;
    CK(cudaMemcpyAsync(synthetics, ctx->d_tmp[3], b_out, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return check_violations(ctx, "stack_batch");
}

// d_hyp + c*hyp_stride + hyper_idx[t] is the hyperparameter of (chain c, target t); logpts[c*logpts_sc + out_ofs + t]
static int misfit_batch_core(beatgpu_ctx* ctx, WaveMap& w, int B, const double* d_resid, const double* d_hyp, long hyp_stride,
                             double* d_logpts, long logpts_sc = -1, int out_ofs = 0, const unsigned char* chain_bad = nullptr)
{
    int rc;
    const long n_hypers = hyp_stride;
    if (logpts_sc < 0) logpts_sc = w.nt;
    if (w.misfit_mode == MISFIT_DENSE && ctx->geo_mode == 1 && B >= 32) {
        // dense weights: Z_t = U_t R_t for all chains on the FP64 tensor cores, straight from the caller's layout
        const int mt = (w.ns + kGemmBM - 1) / kGemmBM;
        if ((rc = ensure_tmp(ctx, 5, (size_t)w.nt * B * mt * sizeof(double)))) return rc;
        GemmArgs g;
        memset(&g, 0, sizeof(g));
        g.M = w.ns; g.N = B; g.K = w.ns; g.n_parts = 1;
        g.A[0] = w.d_W; g.a_sm = 1; g.a_sk = w.ns; g.a_batch = (long)w.ns * w.ns;
        g.B[0] = d_resid; g.b_sk = 1; g.b_sn[0] = (long)w.nt * w.ns; g.b_batch = w.ns;   // resid[c][t][k]
        g.upper = w.dense_upper;
        g.qpart = (double*)ctx->d_tmp[5]; g.n_mtiles = mt; g.q_batch = (long)B * mt;
        CK(launch_dgemm<1>(g, w.nt, ctx->stream)); ctx->n_launches++;
        SeisFinishArgs f;
        memset(&f, 0, sizeof(f));
        f.B = B; f.nt = w.nt; f.n_mtiles = mt; f.qpart = (const double*)ctx->d_tmp[5];
        f.slog_pdet = w.d_slog_pdet; f.nsamp = w.d_nsamp; f.hyper_idx = w.d_hyper_idx;
        f.hyp = d_hyp; f.hyp_sc = n_hypers; f.chain_bad = chain_bad;
        f.logpts = d_logpts; f.logpts_sc = logpts_sc; f.out_ofs = out_ofs;
        seismic_finish_kernel<<<(unsigned)(((long)B * w.nt + 127) / 128), 128, 0, ctx->stream>>>(f);
        CKL();
    } else {
        MisfitArgs a;
        memset(&a, 0, sizeof(a));
        a.B = B; a.nt = w.nt; a.ns = w.ns;
        a.resid = d_resid;
        a.hyp = d_hyp; a.hyp_sc = n_hypers; a.hyper_idx = w.d_hyper_idx;
        a.misfit_mode = w.misfit_mode; a.bw = w.bw; a.dense_upper = w.dense_upper;
        a.W = w.d_W; a.slog_pdet = w.d_slog_pdet; a.nsamp = w.d_nsamp;
        a.logpts = d_logpts; a.logpts_sc = logpts_sc; a.out_ofs = out_ofs; a.chain_bad = chain_bad;
        if ((rc = launch_misfit(ctx, a))) return rc;
    }
    return BEATGPU_OK;
}

int beatgpu_misfit_batch(beatgpu_ctx* ctx, int wmap_id, int B, const double* residuals, const double* hypers, int n_hypers,
                         double* logpts)
{
    GET_WMAP(wmap_id);
    if (B <= 0 || !residuals || !hypers || n_hypers <= 0 || !logpts) return fail(ctx, BEATGPU_E_ARG, "misfit_batch: bad arguments");
    if (w.misfit_mode < 0) return fail(ctx, BEATGPU_E_NOTREADY, "misfit_batch: weights not uploaded (update_weights)");
    int rc;
    const size_t b_r = (size_t)B * w.nt * w.ns * 8, b_h = (size_t)B * n_hypers * 8, b_o = (size_t)B * w.nt * 8;
    if ((rc = ensure_tmp(ctx, 0, b_r)) || (rc = ensure_tmp(ctx, 1, b_h)) || (rc = ensure_tmp(ctx, 2, b_o))) return rc;
    CK(cudaMemcpyAsync(ctx->d_tmp[0], residuals, b_r, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_tmp[1], hypers, b_h, cudaMemcpyHostToDevice, ctx->stream));
    if ((rc = misfit_batch_core(ctx, w, B, (const double*)ctx->d_tmp[0], (const double*)ctx->d_tmp[1], n_hypers, (double*)ctx->d_tmp[2]))) return rc;
    CK(cudaMemcpyAsync(logpts, ctx->d_tmp[2], b_o, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return BEATGPU_OK;
}

int beatgpu_misfit_batch_dev(beatgpu_ctx* ctx, int wmap_id, int B, const double* residuals_dev, const double* hypers_dev,
                             int n_hypers, double* logpts_dev)
{
    GET_WMAP(wmap_id);
    if (B <= 0 || !residuals_dev || !hypers_dev || n_hypers <= 0 || !logpts_dev) return fail(ctx, BEATGPU_E_ARG, "misfit_batch_dev: bad arguments");
    if (w.misfit_mode < 0) return fail(ctx, BEATGPU_E_NOTREADY, "misfit_batch_dev: weights not uploaded (update_weights)");
    return misfit_batch_core(ctx, w, B, residuals_dev, hypers_dev, n_hypers, logpts_dev);
}

// rupture onset times for all chains / subfaults into ctx->d_t0 (seismic.py:1253-1272)
static int run_sweep(beatgpu_ctx* ctx, int B, const double* q)
{
    const beatgpu_layout& L = ctx->layout;
    CK(cudaMemsetAsync(ctx->d_bad, 0, (size_t)B, ctx->stream));
    SweepArgs s;
    memset(&s, 0, sizeof(s));
    s.n_dip = ctx->d_nd; s.n_strike = ctx->d_ns; s.patch_ofs = ctx->d_pofs; s.patch_size = ctx->d_psize;
    s.n_subfaults = ctx->nsf; s.n_patches_total = ctx->np_total; s.max_np_sf = ctx->max_np_sf; s.B = B;
    VarRef vel = var_ref(ctx, q, L.off_velocities, ctx->canon_vel);
    VarRef nd = var_ref(ctx, q, L.off_nucleation_dip, ctx->canon_ndip);
    VarRef nst = var_ref(ctx, q, L.off_nucleation_strike, ctx->canon_nstr);
    VarRef tm = var_ref(ctx, q, L.off_time, ctx->canon_time);
    s.vel = vel.p; s.vel_stride = vel.stride; s.is_slowness = 0;
    s.nuc_dip = nd.p; s.nuc_dip_stride = nd.stride; s.nuc_strike = nst.p; s.nuc_strike_stride = nst.stride;
    s.nuc_dip_idx = nullptr; s.nuc_strike_idx = nullptr; s.only_sf = -1;
    s.time = tm.p; s.time_stride = tm.stride;
    s.t0 = ctx->d_t0; s.n_iter = nullptr; s.violations = ctx->d_viol; s.chain_bad = ctx->d_bad;
    return launch_sweep(ctx, s, B * ctx->nsf);
}

// per-chain inputs of the stacking kernels taken from q through the layout (fused-mode wiring)
static void wire_chain_inputs(beatgpu_ctx* ctx, const WaveMap& w, const double* q, StackArgs& a)
{
    const beatgpu_layout& L = ctx->layout;
    VarRef dur = var_ref(ctx, q, L.off_durations, ctx->canon_dur);
    a.dur = dur.p; a.dur_sc = dur.stride;
    for (int v = 0; v < L.n_slipvars; ++v) {
        VarRef sl = var_ref(ctx, q, L.off_slip[v], ctx->canon_slip[v]);
        a.slip[v] = sl.p; a.slip_sc[v] = sl.stride;
    }
    a.st = ctx->d_t0; a.st_sc = ctx->np_total; a.st_st = 0;
    if (w.has_station) {
        VarRef ts = var_ref(ctx, q, L.off_time_shifts, ctx->canon_ts);
        a.corr = ts.p; a.corr_sc = ts.stride;
    } else {
        a.corr = nullptr; a.station_idx = nullptr;
    }
}

// Host-pointer entry: q [B, n_params] -> ctx->d_q in two phases.  The rupture sweep reads only velocities, nucleation
// point and time (a quarter of a chain's row at C3), so those columns go first on the ctx stream (strided 2-D copy) and
// the sweep starts behind them while the slips / durations / hypers columns are still crossing PCIe on a second stream;
// the stacking pass waits for them through an event.
static int upload_q_split(beatgpu_ctx* ctx, int B, const double* q)
{
    const beatgpu_layout& L = ctx->layout;
    const int np = ctx->np_total, nsf = ctx->nsf, n = L.n_params;
    const size_t pitch = (size_t)n * sizeof(double);
    int lo = n, hi = 0;
    const int offs[4] = {L.off_velocities, L.off_nucleation_strike, L.off_nucleation_dip, L.off_time};
    const int lens[4] = {np, nsf, nsf, nsf};
    for (int i = 0; i < 4; ++i) if (offs[i] >= 0) { lo = std::min(lo, offs[i]); hi = std::max(hi, offs[i] + lens[i]); }
    const bool split = ctx->split_h2d && !ctx->wmaps.empty() && hi > lo && (hi - lo) * 2 <= n && B >= 64;
    if (!split) {
        CK(cudaMemcpyAsync(ctx->d_q, q, (size_t)B * pitch, cudaMemcpyHostToDevice, ctx->stream));
        return BEATGPU_OK;
    }
    if (!ctx->copy_stream) {
        CK(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&ctx->copy_done, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&ctx->copy_go, cudaEventDisableTiming));
    }
    CK(cudaEventRecord(ctx->copy_go, ctx->stream));                      // d_q is free once earlier work on the ctx stream is done
    CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->copy_go, 0));
    CK(cudaMemcpy2DAsync(ctx->d_q + lo, pitch, q + lo, pitch, (size_t)(hi - lo) * sizeof(double), (size_t)B,
                         cudaMemcpyHostToDevice, ctx->stream));
    if (lo > 0)
        CK(cudaMemcpy2DAsync(ctx->d_q, pitch, q, pitch, (size_t)lo * sizeof(double), (size_t)B, cudaMemcpyHostToDevice, ctx->copy_stream));
    if (hi < n)
        CK(cudaMemcpy2DAsync(ctx->d_q + hi, pitch, q + hi, pitch, (size_t)(n - hi) * sizeof(double), (size_t)B, cudaMemcpyHostToDevice,
                             ctx->copy_stream));
    CK(cudaEventRecord(ctx->copy_done, ctx->copy_stream));
    ctx->wait_before_stack = ctx->copy_done;
    return BEATGPU_OK;
}

// the fused path, everything on the device
int beatgpu_ffi_loglike_batch_dev(beatgpu_ctx* ctx, int B, const double* q, double* logpts, double* like)
{
    if (!ctx) return BEATGPU_E_ARG;
    if (B <= 0 || !q || !logpts) return fail(ctx, BEATGPU_E_ARG, "ffi_loglike_batch: bad arguments");
    if (!ctx->layout_set) return fail(ctx, BEATGPU_E_NOTREADY, "ffi_loglike_batch: set_fault / set_layout not called");
    if (ctx->wmaps.empty() && !ctx->geo.set && !ctx->lap.set) return fail(ctx, BEATGPU_E_NOTREADY, "ffi_loglike_batch: no composite configured");
    CK(cudaSetDevice(ctx->device));
    int rc;
    if ((rc = ensure_scratch(ctx, B))) return rc;
    const beatgpu_layout& L = ctx->layout;
    const int n_out = n_outputs(ctx);
    for (auto& w : ctx->wmaps) {
        if (w.geom >= 0) return fail(ctx, BEATGPU_E_ARG, "ffi_loglike_batch: the context holds geometry-mode wavemaps; use beatgpu_geom_loglike_batch");
        if ((rc = check_lib(ctx, w, L.n_slipvars))) return rc;
        if (!w.d_data) return fail(ctx, BEATGPU_E_NOTREADY, "ffi_loglike_batch: data of a wavemap not uploaded");
        if (w.misfit_mode < 0) return fail(ctx, BEATGPU_E_NOTREADY, "ffi_loglike_batch: weights of a wavemap not uploaded");
        if (w.has_station && !L.n_time_shifts) return fail(ctx, BEATGPU_E_ARG, "ffi_loglike_batch: wavemap has station corrections but the layout has no time_shifts");
    }
    VarRef hyp = var_ref(ctx, q, L.n_hypers ? L.off_hypers : -1, ctx->canon_hyp);

    if (!ctx->wmaps.empty()) {
        if ((rc = run_sweep(ctx, B, q))) return rc;
    }
    if (ctx->wait_before_stack) {                    // host-pointer entry: the remaining columns of q arrive on the copy stream
        cudaEvent_t ev = ctx->wait_before_stack;
        ctx->wait_before_stack = nullptr;
        CK(cudaStreamWaitEvent(ctx->stream, ev, 0));
    }
    if (!ctx->wmaps.empty()) {
        // ---- per wavemap: gather + stack + residual + misfit (seismic.py:1275-1343)
        if ((rc = timing_begin(ctx))) return rc;
        for (auto& w : ctx->wmaps) {
            StackArgs a;
            memset(&a, 0, sizeof(a));
            fill_static(w, a, L.n_slipvars);
            a.B = B;
            wire_chain_inputs(ctx, w, q, a);
            a.hyp = hyp.p; a.hyp_sc = hyp.stride;
            a.logpts = logpts; a.logpts_sc = n_out; a.out_ofs = w.out_ofs;
            a.synth = nullptr; a.chain_bad = ctx->d_bad; a.violations = ctx->d_viol;
            // the chunked path needs a scratch of B*nt*nchunk*ns doubles; fall back to the fused kernel beyond 8 GiB
            int chunk_p = 0, nchunk_p = 0;
            derive_chunk(ctx, w, L.n_slipvars, ctx->np_total, &chunk_p, &nchunk_p, nullptr);
            const bool chunked = ctx->stack_mode == 1 && nchunk_p > 1 &&
                                 (size_t)B * w.nt * nchunk_p * w.ns * sizeof(double) <= ((size_t)8 << 30);
            if ((rc = chunked ? launch_stack_chunked(ctx, w, a) : launch_stack<false>(ctx, w, a))) return rc;
        }
        if ((rc = timing_end(ctx))) return rc;
    }

    if (ctx->geo.set && ctx->geo_mode == 1) {
        // batched over chains the geodetic composite is two GEMMs: FP64 tensor-core tiles (gemm.cuh)
        Geodetic& g = ctx->geo;
        const int mt_max = (g.max_n + kGemmBM - 1) / kGemmBM;
        if ((rc = ensure_tmp(ctx, 4, (size_t)B * g.nobs * sizeof(double)))) return rc;          // R [B, nobs]
        if ((rc = ensure_tmp(ctx, 5, (size_t)B * mt_max * sizeof(double)))) return rc;          // partial norms
        GemmArgs ga;
        memset(&ga, 0, sizeof(ga));
        ga.M = g.nobs; ga.N = B; ga.K = ctx->np_total; ga.n_parts = L.n_slipvars;
        ga.a_sm = 1; ga.a_sk = g.nobs; ga.b_sk = 1;
        for (int v = 0; v < L.n_slipvars; ++v) {
            ga.A[v] = g.G[v];                                   // G_v [np, nobs]: A(m=obs, k=patch) = G_v[k*nobs + m]
            VarRef sl = var_ref(ctx, q, L.off_slip[v], ctx->canon_slip[v]);
            ga.B[v] = sl.p; ga.b_sn[v] = sl.stride;             // slip_v(k=patch, n=chain) = q[n*stride + off + k]
        }
        ga.upper = 0; ga.data = g.d_data; ga.odw = g.d_odw; ga.R = (double*)ctx->d_tmp[4]; ga.ldr = g.nobs;
        CK(launch_dgemm<0>(ga, 1, ctx->stream)); ctx->n_launches++;
        for (int d = 0; d < g.nds; ++d) {
            const int n = g.hi[d] - g.lo[d];
            const int mt = (n + kGemmBM - 1) / kGemmBM;
            GemmArgs gb;
            memset(&gb, 0, sizeof(gb));
            gb.M = n; gb.N = B; gb.K = n; gb.n_parts = 1;
            gb.A[0] = g.d_UT + g.ut_ofs[d]; gb.a_sm = 1; gb.a_sk = n;          // U(m, k) = UT[k*n + m]
            gb.B[0] = (const double*)ctx->d_tmp[4] + g.lo[d]; gb.b_sk = 1; gb.b_sn[0] = g.nobs;
            gb.upper = g.h_upper[d];
            gb.qpart = (double*)ctx->d_tmp[5]; gb.n_mtiles = mt;
            CK(launch_dgemm<1>(gb, 1, ctx->stream)); ctx->n_launches++;
            GeoFinishArgs f;
            memset(&f, 0, sizeof(f));
            f.B = B; f.n_mtiles = mt; f.qpart = (const double*)ctx->d_tmp[5];
            f.slog_pdet = g.h_slog[d]; f.nsamp = g.h_nsamp[d]; f.hyper_idx = g.h_hyper_idx[d];
            f.hyp = hyp.p; f.hyp_sc = hyp.stride;
            f.logpts = logpts; f.logpts_sc = n_out; f.out_col = g.out_ofs + d;
            geodetic_finish_kernel<<<(B + 127) / 128, 128, 0, ctx->stream>>>(f);
            CKL();
        }
    } else if (ctx->geo.set) {
        Geodetic& g = ctx->geo;
        GeoArgs a;
        memset(&a, 0, sizeof(a));
        a.B = B; a.np = ctx->np_total; a.nobs = g.nobs; a.ndatasets = g.nds; a.nvar = L.n_slipvars;
        for (int v = 0; v < L.n_slipvars; ++v) {
            a.G[v] = g.G[v];
            VarRef sl = var_ref(ctx, q, L.off_slip[v], ctx->canon_slip[v]);
            a.slip[v] = sl.p; a.slip_sc[v] = sl.stride;
        }
        a.data = g.d_data; a.odw = g.d_odw; a.lo = g.d_lo; a.hi = g.d_hi; a.UT = g.d_UT; a.UT_ofs = g.d_UT_ofs;
        a.upper = g.d_upper; a.slog_pdet = g.d_slog; a.nsamp = g.d_nsamp; a.hyper_idx = g.d_hyper_idx;
        a.hyp = hyp.p; a.hyp_sc = hyp.stride;
        a.logpts = logpts; a.logpts_sc = n_out; a.out_ofs = g.out_ofs; a.max_n = g.max_n;
        const size_t smem = ((size_t)L.n_slipvars * ctx->np_total + g.max_n) * sizeof(double);
        if (smem > (size_t)ctx->prop.sharedMemPerBlockOptin) return fail(ctx, BEATGPU_E_ARG, "geodetic dataset too large for shared memory (%zu B)", smem);
        if (smem > 48 * 1024) CK(cudaFuncSetAttribute(geodetic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        geodetic_kernel<<<(unsigned)((long)g.nds * B), 256, smem, ctx->stream>>>(a);
        CKL();
    }

    if (ctx->lap.set && ctx->geo_mode == 1 && B >= 32) {
        // batched over chains |L u_v|^2 is a GEMM per slip variable with a column-norm epilogue: FP64 tensor-core tiles
        const int np = ctx->np_total, mt = (np + kGemmBM - 1) / kGemmBM;
        if ((rc = ensure_tmp(ctx, 5, (size_t)L.n_slipvars * B * mt * sizeof(double)))) return rc;
        for (int v = 0; v < L.n_slipvars; ++v) {
            GemmArgs gl;
            memset(&gl, 0, sizeof(gl));
            gl.M = np; gl.N = B; gl.K = np; gl.n_parts = 1;
            gl.A[0] = ctx->lap.d_LT; gl.a_sm = 1; gl.a_sk = np;                       // L(m, k) = LT[k*np + m]
            VarRef sl = var_ref(ctx, q, L.off_slip[v], ctx->canon_slip[v]);
            gl.B[0] = sl.p; gl.b_sk = 1; gl.b_sn[0] = sl.stride;
            gl.qpart = (double*)ctx->d_tmp[5] + (size_t)v * B * mt; gl.n_mtiles = mt;
            CK(launch_dgemm<1>(gl, 1, ctx->stream)); ctx->n_launches++;
        }
        LapFinishArgs f;
        memset(&f, 0, sizeof(f));
        f.B = B; f.n_mtiles = mt; f.nvar = L.n_slipvars; f.np = np; f.qpart = (const double*)ctx->d_tmp[5];
        f.sdet = ctx->lap.sdet; f.hyper_idx = ctx->lap.hyper_idx; f.hyp = hyp.p; f.hyp_sc = hyp.stride;
        f.logpts = logpts; f.logpts_sc = n_out; f.out_col = ctx->lap.out_ofs;
        laplacian_finish_kernel<<<(B + 127) / 128, 128, 0, ctx->stream>>>(f);
        CKL();
    } else if (ctx->lap.set) {
        LapArgs a;
        memset(&a, 0, sizeof(a));
        a.B = B; a.np = ctx->np_total; a.nvar = L.n_slipvars;
        a.LT = ctx->lap.d_LT; a.sdet = ctx->lap.sdet; a.hyper_idx = ctx->lap.hyper_idx;
        for (int v = 0; v < L.n_slipvars; ++v) {
            VarRef sl = var_ref(ctx, q, L.off_slip[v], ctx->canon_slip[v]);
            a.slip[v] = sl.p; a.slip_sc[v] = sl.stride;
        }
        a.hyp = hyp.p; a.hyp_sc = hyp.stride;
        a.logpts = logpts; a.logpts_sc = n_out; a.out_ofs = ctx->lap.out_ofs;
        laplacian_kernel<<<B, 256, (size_t)ctx->np_total * sizeof(double), ctx->stream>>>(a);
        CKL();
    }

    if (like) {
        sum_like_kernel<<<(B + 127) / 128, 128, 0, ctx->stream>>>(logpts, like, B, n_out);
        CKL();
    }
    return BEATGPU_OK;
}

int beatgpu_ffi_loglike_batch(beatgpu_ctx* ctx, int B, const double* q, double* logpts, double* like)
{
    if (!ctx) return BEATGPU_E_ARG;
    if (B <= 0 || !q || !logpts) return fail(ctx, BEATGPU_E_ARG, "ffi_loglike_batch: bad arguments");
    if (!ctx->layout_set) return fail(ctx, BEATGPU_E_NOTREADY, "ffi_loglike_batch: set_fault / set_layout not called");
    CK(cudaSetDevice(ctx->device));
    int rc;
    if ((rc = ensure_scratch(ctx, B))) return rc;
    const int n_out = n_outputs(ctx);
    // violations are reported per call: what earlier device-pointer evaluations (a sampler's rejected proposals) counted
    // is not this call's business
    CK(cudaMemsetAsync(ctx->d_viol, 0, sizeof(unsigned long long), ctx->stream));
    if ((rc = upload_q_split(ctx, B, q))) return rc;
    if ((rc = beatgpu_ffi_loglike_batch_dev(ctx, B, ctx->d_q, ctx->d_logpts, ctx->d_like))) return rc;
    CK(cudaMemcpyAsync(logpts, ctx->d_logpts, (size_t)B * n_out * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (like) CK(cudaMemcpyAsync(like, ctx->d_like, (size_t)B * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    return check_violations(ctx, "ffi_loglike_batch");
}

int beatgpu_ffi_synthetics_batch(beatgpu_ctx* ctx, int wmap_id, int B, const double* q, double* synthetics)
{
    GET_WMAP(wmap_id);
    if (B <= 0 || !q || !synthetics) return fail(ctx, BEATGPU_E_ARG, "ffi_synthetics_batch: bad arguments");
    if (!ctx->layout_set) return fail(ctx, BEATGPU_E_NOTREADY, "ffi_synthetics_batch: set_fault / set_layout not called");
    int rc;
    if ((rc = check_lib(ctx, w, ctx->layout.n_slipvars))) return rc;
    if (w.has_station && !ctx->layout.n_time_shifts) return fail(ctx, BEATGPU_E_ARG, "ffi_synthetics_batch: wavemap has station corrections but the layout has no time_shifts");
    if ((rc = ensure_scratch(ctx, B))) return rc;
    const size_t b_out = (size_t)B * w.nt * w.ns * sizeof(double);
    if ((rc = ensure_tmp(ctx, 3, b_out))) return rc;
    CK(cudaMemsetAsync(ctx->d_viol, 0, sizeof(unsigned long long), ctx->stream));      // violations are reported per call
    CK(cudaMemcpyAsync(ctx->d_q, q, (size_t)B * ctx->layout.n_params * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    if ((rc = run_sweep(ctx, B, ctx->d_q))) return rc;
    StackArgs a;
    memset(&a, 0, sizeof(a));
    fill_static(w, a, ctx->layout.n_slipvars);
    a.B = B;
    wire_chain_inputs(ctx, w, ctx->d_q, a);
    a.synth = (double*)ctx->d_tmp[3];
    a.chain_bad = ctx->d_bad; a.violations = ctx->d_viol;
    if ((rc = launch_stack<true>(ctx, w, a))) return rc;
    CK(cudaMemcpyAsync(synthetics, ctx->d_tmp[3], b_out, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return check_violations(ctx, "ffi_synthetics_batch");
}

int beatgpu_get_starttimes(beatgpu_ctx* ctx, int B, double* starttimes)
{
    if (!ctx || !starttimes) return BEATGPU_E_ARG;
    if (B <= 0 || B > ctx->cap_B || !ctx->d_t0) return fail(ctx, BEATGPU_E_NOTREADY, "get_starttimes: no evaluation of >= %d chains yet", B);
    CK(cudaMemcpyAsync(starttimes, ctx->d_t0, (size_t)B * ctx->np_total * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return BEATGPU_OK;
}

int beatgpu_index_violations(beatgpu_ctx* ctx, int64_t* count)
{
    if (!ctx || !count) return BEATGPU_E_ARG;
    unsigned long long v = 0;
    CK(cudaMemcpyAsync(&v, ctx->d_viol, sizeof(v), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemsetAsync(ctx->d_viol, 0, sizeof(v), ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    *count = (int64_t)v;
    return BEATGPU_OK;
}

int beatgpu_launch_count(beatgpu_ctx* ctx, int64_t* n)
{
    if (!ctx || !n) return BEATGPU_E_ARG;
    *n = ctx->n_launches;
    return BEATGPU_OK;
}

int beatgpu_last_stack_ms(beatgpu_ctx* ctx, float* ms)
{
    if (!ctx || !ms) return BEATGPU_E_ARG;
    if (!ctx->ev_valid || !ctx->ev1) return fail(ctx, BEATGPU_E_NOTREADY, "last_stack_ms: no (uncaptured) fused evaluation yet");
    CK(cudaEventSynchronize(ctx->ev1));
    CK(cudaEventElapsedTime(ms, ctx->ev0, ctx->ev1));
    return BEATGPU_OK;
}

int beatgpu_stack_ms_accum(beatgpu_ctx* ctx, int reset, double* sum_ms, int64_t* n_evals)
{
    if (!ctx || !sum_ms || !n_evals) return BEATGPU_E_ARG;
    *sum_ms = 0.0;
    *n_evals = 0;
    const int n = ctx->tev_pending;
    if (n > 0) {
        CK(cudaSetDevice(ctx->device));
        const int last = (ctx->tev_next + kTimingRing - 1) % kTimingRing;
        CK(cudaEventSynchronize(ctx->tev[2 * last + 1]));
        for (int i = 0; i < n; ++i) {
            const int k = (ctx->tev_next + kTimingRing - 1 - i) % kTimingRing;
            float ms = 0.f;
            CK(cudaEventElapsedTime(&ms, ctx->tev[2 * k], ctx->tev[2 * k + 1]));
            *sum_ms += (double)ms;
        }
        *n_evals = n;
    }
    if (reset) ctx->tev_pending = 0;
    return BEATGPU_OK;
}

int beatgpu_stack_blocking(beatgpu_ctx* ctx, int wmap_id, int n_slipvars, int* chunk_patches, int* n_chunks,
                           int64_t* chunk_bytes, int64_t* l2_bytes)
{
    GET_WMAP(wmap_id);
    if (!w.axes_set) return fail(ctx, BEATGPU_E_NOTREADY, "stack_blocking: no GF library declared for wavemap %d", wmap_id);
    if (n_slipvars < 1 || n_slipvars > BEATGPU_MAX_SLIPVARS) return fail(ctx, BEATGPU_E_ARG, "stack_blocking: n_slipvars %d", n_slipvars);
    int chunk = 0, nchunk = 0;
    int64_t cb = 0;
    derive_chunk(ctx, w, n_slipvars, (int)w.dims[1], &chunk, &nchunk, &cb);
    if (chunk_patches) *chunk_patches = chunk;
    if (n_chunks) *n_chunks = nchunk;
    if (chunk_bytes) *chunk_bytes = cb;
    if (l2_bytes) *l2_bytes = (int64_t)ctx->prop.l2CacheSize;
    return BEATGPU_OK;
}

#ifndef BEATGPU_SRC_HASH
#define BEATGPU_SRC_HASH "unknown"
#endif
const char* beatgpu_source_hash(void) { return BEATGPU_SRC_HASH; }

int beatgpu_probe_gather(beatgpu_ctx* ctx, int mode, int64_t ws_bytes, int row_bytes, int rows_per_warp, int n_launch,
                         float* ms_per_launch, double* bytes_per_launch)
{
    if (!ctx || !ms_per_launch || !bytes_per_launch) return BEATGPU_E_ARG;
    if (mode < 0 || mode > 8 || row_bytes < 16 || row_bytes > kProbeMaxRow || row_bytes % 16 || ws_bytes < row_bytes ||
        rows_per_warp < 1 || n_launch < 1)
        return fail(ctx, BEATGPU_E_ARG, "probe_gather: bad arguments (mode 0..8; row_bytes: multiple of 16, <= %d)", kProbeMaxRow);
    CK(cudaSetDevice(ctx->device));
    ProbeArgs a;
    memset(&a, 0, sizeof(a));
    a.row_bytes = row_bytes;
    a.row_stride = row_bytes;
    a.n_rows = 1;
    while (a.n_rows * 2 * row_bytes <= ws_bytes) a.n_rows *= 2;     // power of two: the row pick is a mask, not a modulo
    if (a.n_rows > 0x40000000L) return fail(ctx, BEATGPU_E_ARG, "probe_gather: working set too large");
    a.row_mask = (uint32_t)(a.n_rows - 1);
    a.rows_per_warp = rows_per_warp;
    const bool smem_only = mode >= 5;                              // rows live in (distributed) shared memory: no working set
    unsigned char* d_ws = nullptr;
    float* d_sink = nullptr;
    if (!smem_only) {
        CK(cudaMalloc((void**)&d_ws, (size_t)a.n_rows * row_bytes));
        CK(cudaMemsetAsync(d_ws, 0, (size_t)a.n_rows * row_bytes, ctx->stream));
    }
    CK(cudaMalloc((void**)&d_sink, 8));
    CK(cudaMemsetAsync(d_sink, 0, 8, ctx->stream));
    a.ws = d_ws; a.sink = d_sink; a.err = (unsigned int*)(d_sink + 1);
    // LDG mode: 16 CTAs x 4 warps = all 64 warp slots of every SM.  TMA modes: ring of `depth` rows (modes 1, 2) or of
    // kBatchRing batches of kBatchRows rows (modes 3, 4) per warp in dynamic shared memory, as many CTAs per SM as ~200 KB
    // of rings allow.  Shared-memory modes (5..8): 32 KB of rows per CTA, cluster of 1 / 2 / 4 / 8 CTAs.
    int per_sm = 16, cluster = 1, rows_smem = 0;
    size_t smem = 0;
    if (mode == 1 || mode == 2) {
        a.depth = (int)std::max<size_t>(1, std::min<size_t>(kProbeDepth, (size_t)(48 * 1024) / ((size_t)kProbeWarps * row_bytes)));
        smem = (size_t)kProbeWarps * a.depth * row_bytes + (size_t)kProbeWarps * a.depth * sizeof(uint64_t);
        per_sm = (int)std::max<size_t>(1, std::min<size_t>(16, (size_t)(200 * 1024) / (smem + 1024)));
        if (smem > 48 * 1024) {
            CK(cudaFuncSetAttribute(probe_tma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            CK(cudaFuncSetAttribute(probe_tma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        }
    } else if (mode == 3 || mode == 4) {
        smem = (size_t)kProbeWarps * kBatchRing * kBatchRows * row_bytes + (size_t)kProbeWarps * kBatchRing * sizeof(uint64_t);
        if (smem > (size_t)ctx->prop.sharedMemPerBlockOptin) {
            cudaFree(d_ws); cudaFree(d_sink);
            return fail(ctx, BEATGPU_E_ARG, "probe_gather: batched bulk-copy ring of %zu B exceeds shared memory (row_bytes too large)", smem);
        }
        per_sm = (int)std::max<size_t>(1, std::min<size_t>(16, (size_t)(216 * 1024) / (smem + 1024)));
        CK(cudaFuncSetAttribute(probe_tma_batch_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CK(cudaFuncSetAttribute(probe_tma_batch_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        a.rows_per_warp = std::max(kBatchRows, rows_per_warp / kBatchRows * kBatchRows);
    } else if (smem_only) {
        cluster = 1 << (mode - 5);
        rows_smem = 1;
        while (rows_smem * 2 * row_bytes <= 32 * 1024) rows_smem *= 2;      // power of two (the kernel masks)
        smem = (size_t)rows_smem * row_bytes;
        per_sm = (int)std::max<size_t>(1, std::min<size_t>(16, (size_t)(216 * 1024) / (smem + 1024)));
        CK(cudaFuncSetAttribute(probe_dsmem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    int grid = ctx->prop.multiProcessorCount * per_sm;
    grid -= grid % cluster;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    for (int i = 0; i < n_launch + 1; ++i) {                   // first launch = warm-up (fills L2)
        if (i == 1) CK(cudaEventRecord(e0, ctx->stream));
        if (mode == 0) probe_ldg_kernel<<<grid, kProbeThreads, 0, ctx->stream>>>(a);
        else if (mode == 1) probe_tma_kernel<true><<<grid, kProbeThreads, smem, ctx->stream>>>(a);
        else if (mode == 2) probe_tma_kernel<false><<<grid, kProbeThreads, smem, ctx->stream>>>(a);
        else if (mode == 3) probe_tma_batch_kernel<true><<<grid, kProbeThreads, smem, ctx->stream>>>(a);
        else if (mode == 4) probe_tma_batch_kernel<false><<<grid, kProbeThreads, smem, ctx->stream>>>(a);
        else {
            cudaLaunchConfig_t cfg;
            memset(&cfg, 0, sizeof(cfg));
            cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(kProbeThreads); cfg.dynamicSmemBytes = smem; cfg.stream = ctx->stream;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = (unsigned)cluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr; cfg.numAttrs = 1;
            CK(cudaLaunchKernelEx(&cfg, probe_dsmem_kernel, a, rows_smem, cluster));
        }
        CKL();
    }
    CK(cudaEventRecord(e1, ctx->stream));
    CK(cudaEventSynchronize(e1));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    unsigned int err = 0;
    CK(cudaMemcpy(&err, a.err, sizeof(err), cudaMemcpyDeviceToHost));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(d_ws); cudaFree(d_sink);
    if (err) return fail(ctx, BEATGPU_E_CUDA, "probe_gather: %u warps timed out waiting for a bulk copy", err);
    *ms_per_launch = ms / n_launch;
    *bytes_per_launch = (double)grid * kProbeWarps * (double)a.rows_per_warp * row_bytes;
    return BEATGPU_OK;
}

}  // extern "C"

#include "geom_host.inc"
#include "trace_io.inc"
