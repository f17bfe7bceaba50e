// stack.cuh -- the roofline kernel: fused GF-library gather + slip-weighted stack over patches
//              + residual + covariance-weighted misfit, one CTA per (chain, target).
//
// Replaces, for B chains at once, the sub-graph the reference builds in
// SeismicDistributerComposite.get_formula (beat/models/seismic.py:1283-1341):
//   starttimes[t,p] = t0[p] - corr[station(t)]                          (:1283-1296)
//   synth[t,:]     += SeismicGFLibrary.stack_all(...) for every slip var (:1317-1330; beat/ffi/base.py:607-709,
//                     index mapping :486-521,:535-568)
//   residual        = data - synth                                       (:1332)
//   logpts[t]       = multivariate_normal_chol(...)                      (:1335-1341; beat/models/distributions.py:119-138)
//
// Data layout in HBM: library G_var is (ntargets, npatches, ndurations, nstarttimes, ld) with ld = nsamples rounded
// up to a 16-byte multiple, float32 or float64.  One "row" = the ld contiguous samples of one (t, p, d, s) entry;
// the kernel only ever reads whole rows, 16 bytes per lane (one LDG.128 per lane per row for f32 with ns <= 128).
//
// Work decomposition: grid = nt * B CTAs, target-major (blockIdx = t*B + c), so CTAs resident at the same time
// work on the same target and march through its patches in near lock-step: rows of one (t, p) block
// (ndur*nst rows, ~0.5 MB) requested by different chains are served from the 126 MB L2 instead of HBM.
// Inside the CTA: (1) all threads build a per-patch "plan" in shared memory (row indices + f64 weights of the
// K taps x nvar slip components; every index/weight is computed in f64 exactly as the reference does);
// (2) each warp streams its share of patches, all K*nvar row loads of two patches in flight per lane, and
// accumulates in f64 registers; (3) cross-warp reduction, residual against the data row, (4) misfit epilogue:
// z = U r with U diagonal / upper-banded / dense, quad = z.z via warp shuffles, logpt written per (chain, dataset).
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

namespace beatgpu {

constexpr int kStackThreads = 128;
constexpr int kStackWarps = kStackThreads / 32;
constexpr int kPlanChunk = 256;          // patches planned per pass (bounds shared memory for any npatches)
constexpr int kWindow = 128;             // samples per pass (one 16-byte vector per lane for f32)

enum MisfitMode { MISFIT_DIAG = 0, MISFIT_BAND = 1, MISFIT_DENSE = 2 };

struct StackArgs {
    // library
    const void* G[BEATGPU_MAX_SLIPVARS];
    int nvar;
    int nt, np, ndur, nst, ns;
    long ld;                              // row stride in elements
    double dur_min, dur_step, st_min, st_step;
    int B;
    // per-chain inputs (pointer + stride in doubles between chains; stride 0 = shared by all chains)
    const double* dur;  long dur_sc;                                     // [np]
    const double* slip[BEATGPU_MAX_SLIPVARS]; long slip_sc[BEATGPU_MAX_SLIPVARS];   // [np]
    const double* st;   long st_sc; long st_st;                          // start times [np] per chain (st_st: stride between targets, 0 in fused mode)
    const double* corr; long corr_sc;                                    // time_shifts [n_time_shifts] or nullptr
    const int* station_idx;                                              // [nt] or nullptr
    const double* hyp;  long hyp_sc;                                     // hypers [n_hypers]
    const int* hyper_idx;                                                // [nt]
    // static per-target operands
    const double* data;                   // [nt, ns]
    int misfit_mode; int bw; int dense_upper;
    const double* W;                      // DIAG: [nt, ns]; BAND: [nt, bw+1, ns] (W[t][j][k] = U[k][k+j]); DENSE: [nt, ns, ns] transposed (W[t][j][k] = U[k][j])
    const double* slog_pdet;              // [nt]
    const int* nsamp;                     // [nt]  (M)
    // outputs
    double* logpts; long logpts_sc; int out_ofs;     // logpts[c*logpts_sc + out_ofs + t]
    double* synth;                        // optional [B, nt, ns]
    const unsigned char* chain_bad;       // optional [B]
    unsigned long long* violations;
};

// WT = float for f32 libraries (weights are rounded once, like the library values themselves), double for f64.
template <typename WT, int K, int NVAR>
struct __align__(16) PatchPlan {
    uint32_t off[4];           // offset of each tap's library row in 16-BYTE UNITS (row index * row_bytes / 16; rows are 16-byte
                               // multiples and a library is < 64 GB, checked at upload): one IMAD.WIDE per row turns it into
                               // the load address.  Taps with zero weight already point at a valid row.
    WT w[K * NVAR];            // weight of tap k, slip variable v at w[v*K + k]
};

// sample index (within the window) of element e (0..3) held by vector slot j
template <typename T> __device__ __forceinline__ int slot_sample(int j, int e);
template <> __device__ __forceinline__ int slot_sample<float>(int j, int e) { return 4 * j + e; }
// f64 rows are read as two 16-byte vectors per lane: vector j covers samples 2j,2j+1 and vector j+32 covers 64+2j,64+2j+1
template <> __device__ __forceinline__ int slot_sample<double>(int j, int e) { return (e < 2) ? (2 * j + e) : (64 + 2 * j + (e - 2)); }

// Plan of one (chain, target, patch): library row offsets + f64-computed weights, exactly the reference's index
// arithmetic (beat/ffi/base.py:506-517 start times, :553-564 durations, :676-679 multilinear weights; station
// correction beat/models/seismic.py:1283-1291).  Returns true when a tap with non-zero weight leaves the library.
template <typename T, int K, int NVAR>
__device__ __forceinline__ bool make_patch_plan(const StackArgs& a, int c, int t, int p, double dur_p, double st_p, double corr,
                                                PatchPlan<T, K, NVAR>& pl)
{
    const double x = (dur_p - a.dur_min) / a.dur_step;                        // base.py:556 / :561
    const double y = ((st_p - corr) - a.st_min) / a.st_step;                  // seismic.py:1283-1291, base.py:509 / :514
    const long rows_per_patch = (long)a.ndur * a.nst;
    const long base = ((long)t * a.np + p) * rows_per_patch;
    const uint32_t row16 = (uint32_t)(a.ld * (long)sizeof(T) / 16);           // 16-byte units per row
    bool viol;
    if (K == 1) {                                                             // nearest neighbour (base.py:506-512,553-559)
        const int di = (int)rint(x);          // round-half-even; the int16 cast of the reference cannot matter in range
        const int si = (int)rint(y);
        viol = (x != x) || (y != y) || (di < 0) || (di >= a.ndur) || (si < 0) || (si >= a.nst);
        pl.off[0] = viol ? 0u : (uint32_t)(base + (long)di * a.nst + si) * row16;
        pl.off[1] = pl.off[2] = pl.off[3] = 0u;
#pragma unroll
        for (int v = 0; v < NVAR; ++v) pl.w[v] = (T)a.slip[v][(long)c * a.slip_sc[v] + p];
    } else {                                                                  // multilinear (base.py:513-517,560-564,662-679)
        const int dc = (int)ceil(x);
        const int sc = (int)ceil(y);
        const double rf = (double)dc - x;
        const double sf = (double)sc - y;
        // a "floor" tap has weight exactly 0 when the coordinate is integral; numpy then reads a wrapped (valid) row and
        // multiplies by 0 -- we read the ceil row instead.  Any tap with non-zero weight outside the library is a violation.
        const int dfl = (rf == 0.0) ? dc : dc - 1;
        const int sfl = (sf == 0.0) ? sc : sc - 1;
        viol = (x != x) || (y != y) || (dc < 0) || (dc >= a.ndur) || (dfl < 0) || (sc < 0) || (sc >= a.nst) || (sfl < 0);
        if (viol) {
            pl.off[0] = pl.off[1] = pl.off[2] = pl.off[3] = 0u;
        } else {
            pl.off[0] = (uint32_t)(base + (long)dc * a.nst + sc) * row16;      // st ceil,  rt ceil
            pl.off[1] = (uint32_t)(base + (long)dc * a.nst + sfl) * row16;     // st floor, rt ceil
            pl.off[2] = (uint32_t)(base + (long)dfl * a.nst + sc) * row16;     // st ceil,  rt floor
            pl.off[3] = (uint32_t)(base + (long)dfl * a.nst + sfl) * row16;    // st floor, rt floor
        }
        const double w_cc = (1.0 - sf) * (1.0 - rf);
        const double w_fc = sf * (1.0 - rf);
        const double w_cf = (1.0 - sf) * rf;
        const double w_ff = sf * rf;
#pragma unroll
        for (int v = 0; v < NVAR; ++v) {
            const double u = a.slip[v][(long)c * a.slip_sc[v] + p];
            pl.w[v * K + 0] = (T)(w_cc * u);
            pl.w[v * K + 1] = (T)(w_fc * u);
            pl.w[v * K + 2] = (T)(w_cf * u);
            pl.w[v * K + 3] = (T)(w_ff * u);
        }
    }
    if (viol) {
#pragma unroll
        for (int q = 0; q < K * NVAR; ++q) pl.w[q] = (T)0;
    }
    return viol;
}

// address of 16-byte vector `vec` of the row at plan offset `off16` (one IMAD.WIDE.U32 + the load)
__device__ __forceinline__ const char* row_ptr(const char* lane_base, uint32_t off16) { return lane_base + (size_t)off16 * 16u; }

template <typename T, int K, int NVAR, bool WRITE_SYNTH>
__global__ void __launch_bounds__(kStackThreads)
gf_stack_misfit_kernel(StackArgs a)
{
    using Plan = PatchPlan<T, K, NVAR>;
    __shared__ Plan plan[kPlanChunk];
    __shared__ double red[kStackWarps][kWindow];
    __shared__ double red_q[kStackWarps];
    __shared__ int s_bad;
    extern __shared__ double resid[];            // [ns] residual (or synth) of this (chain, target)

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    // Persistent CTAs: the grid is sized to the number of co-resident CTAs and every CTA walks the (target, chain)
    // items with stride gridDim.  All CTAs start together and every item costs the same, so at any moment the
    // resident CTAs are at (nearly) the same patch of the same target: the rows they gather come from a working
    // set of a few (t, p) blocks (~1 MB each) that stays in L2, instead of being spread over the whole target.
    const long n_items = (long)a.nt * a.B;
    for (long item = blockIdx.x; item < n_items; item += gridDim.x) {
    const int c = (int)(item % a.B);
    const int t = (int)(item / a.B);

    __syncthreads();                                         // previous item fully retired (shared memory reuse)
    if (tid == 0) s_bad = (a.chain_bad && a.chain_bad[c]) ? 1 : 0;

    const double* dur = a.dur + (long)c * a.dur_sc;
    const double* st = a.st + (long)c * a.st_sc + (long)t * a.st_st;
    double corr = 0.0;
    if (a.corr) corr = a.corr[(long)c * a.corr_sc + a.station_idx[t]];

    for (int s0 = 0; s0 < a.ns; s0 += kWindow) {
        const int wlen = min(kWindow, a.ns - s0);
        const int nvec = (wlen * (int)sizeof(T) + 15) / 16;      // 16-byte vectors in this window of a row
        double acc[4] = {0.0, 0.0, 0.0, 0.0};

        for (int p0 = 0; p0 < a.np; p0 += kPlanChunk) {
            const int pn = min(kPlanChunk, a.np - p0);
            __syncthreads();                                     // previous chunk fully consumed
            // ---------------- (1) plan: indices + weights, f64, reference arithmetic ----------------
            for (int i = tid; i < pn; i += kStackThreads) {
                const int p = p0 + i;
                Plan pl;
                if (make_patch_plan<T, K, NVAR>(a, c, t, p, dur[p], st[p], corr, pl)) {
                    if (s0 == 0) atomicAdd(a.violations, 1ULL);
                    s_bad = 1;
                }
                plan[i] = pl;
            }
            __syncthreads();

            // ---------------- (2) stream rows: K*NVAR rows per patch, two patches in flight ----------------
            if (sizeof(T) == 4) {
                const bool active = lane < nvec;
                for (int i = warp; i < pn; i += 2 * kStackWarps) {
                    const int i2 = i + kStackWarps;
                    const bool has2 = i2 < pn;
                    float4 g[2][K * NVAR];
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        const int ii = (u == 0) ? i : (has2 ? i2 : i);
                        const bool ld_on = active && (u == 0 || has2);
#pragma unroll
                        for (int v = 0; v < NVAR; ++v)
#pragma unroll
                            for (int k = 0; k < K; ++k) {
                                const char* row = row_ptr(reinterpret_cast<const char*>(a.G[v]) + (size_t)s0 * sizeof(float), plan[ii].off[k]);
                                g[u][v * K + k] = ld_on ? __ldg(reinterpret_cast<const float4*>(row) + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
                            }
                    }
                    // f32 library: products and the 2*K*NVAR-term partial sum in f32 (FFMA pipe), one f32->f64
                    // conversion per element per patch pair; the running sum over patches stays f64.
                    float4 part = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        if (u == 1 && !has2) break;
                        const int ii = (u == 0) ? i : i2;
#pragma unroll
                        for (int q = 0; q < K * NVAR; ++q) {
                            const float w = plan[ii].w[q];
                            part.x = fmaf(w, g[u][q].x, part.x);
                            part.y = fmaf(w, g[u][q].y, part.y);
                            part.z = fmaf(w, g[u][q].z, part.z);
                            part.w = fmaf(w, g[u][q].w, part.w);
                        }
                    }
                    acc[0] += (double)part.x;
                    acc[1] += (double)part.y;
                    acc[2] += (double)part.z;
                    acc[3] += (double)part.w;
                }
            } else {
                // f64 storage: a window of 128 samples = 64 16-byte vectors; lane reads vectors `lane` and `lane+32`
                const bool act0 = lane < nvec, act1 = (lane + 32) < nvec;
                for (int i = warp; i < pn; i += kStackWarps) {
                    double2 g0[K * NVAR], g1[K * NVAR];
#pragma unroll
                    for (int v = 0; v < NVAR; ++v)
#pragma unroll
                        for (int k = 0; k < K; ++k) {
                            const char* row = row_ptr(reinterpret_cast<const char*>(a.G[v]) + (size_t)s0 * sizeof(double), plan[i].off[k]);
                            g0[v * K + k] = act0 ? __ldg(reinterpret_cast<const double2*>(row) + lane) : make_double2(0.0, 0.0);
                            g1[v * K + k] = act1 ? __ldg(reinterpret_cast<const double2*>(row) + lane + 32) : make_double2(0.0, 0.0);
                        }
#pragma unroll
                    for (int q = 0; q < K * NVAR; ++q) {
                        const double w = plan[i].w[q];
                        acc[0] = fma(w, g0[q].x, acc[0]);
                        acc[1] = fma(w, g0[q].y, acc[1]);
                        acc[2] = fma(w, g1[q].x, acc[2]);
                        acc[3] = fma(w, g1[q].y, acc[3]);
                    }
                }
            }
        }

        // ---------------- (3) cross-warp reduction of this window, residual ----------------
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int sidx = slot_sample<T>(lane, e);
            if (sidx < kWindow) red[warp][sidx] = acc[e];
        }
        __syncthreads();
        for (int k = tid; k < wlen; k += kStackThreads) {
            double s = red[0][k];
#pragma unroll
            for (int w = 1; w < kStackWarps; ++w) s += red[w][k];
            if (WRITE_SYNTH) {
                a.synth[((long)c * a.nt + t) * a.ns + s0 + k] = s;
            } else {
                resid[s0 + k] = a.data[(long)t * a.ns + s0 + k] - s;                     // seismic.py:1332
            }
        }
        __syncthreads();
    }

    if (WRITE_SYNTH) {
        if (s_bad) for (int k = tid; k < a.ns; k += kStackThreads) a.synth[((long)c * a.nt + t) * a.ns + k] = CUDART_NAN;
        continue;
    }

    // ---------------- (4) misfit: quad = |U r|^2  (distributions.py:128,136) ----------------
    double q = 0.0;
    const int ns = a.ns;
    if (a.misfit_mode == MISFIT_DIAG) {
        const double* Wt = a.W + (long)t * ns;
        for (int k = tid; k < ns; k += kStackThreads) { const double z = Wt[k] * resid[k]; q = fma(z, z, q); }
    } else if (a.misfit_mode == MISFIT_BAND) {
        const double* Wt = a.W + (long)t * (a.bw + 1) * ns;
        for (int k = tid; k < ns; k += kStackThreads) {
            double z = 0.0;
            const int jmax = min(a.bw, ns - 1 - k);
            for (int j = 0; j <= jmax; ++j) z = fma(Wt[(long)j * ns + k], resid[k + j], z);
            q = fma(z, z, q);
        }
    } else {
        const double* Wt = a.W + (long)t * ns * ns;
        for (int k = tid; k < ns; k += kStackThreads) {
            double z = 0.0;
            for (int j = a.dense_upper ? k : 0; j < ns; ++j) z = fma(Wt[(long)j * ns + k], resid[j], z);
            q = fma(z, z, q);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    if (lane == 0) red_q[warp] = q;
    __syncthreads();
    if (tid == 0) {
        double quad = red_q[0];
#pragma unroll
        for (int w = 1; w < kStackWarps; ++w) quad += red_q[w];
        const double hp = a.hyp[(long)c * a.hyp_sc + a.hyper_idx[t]];
        const double M = (double)(short)a.nsamp[t];                                       // tt.cast(..., "int16") distributions.py:120
        const double norm = M * (2.0 * hp + 1.8378770664093453);                          // log(2*pi), distributions.py:13,129
        double lp = (-0.5) * (a.slog_pdet[t] + norm + (1.0 / exp(hp * 2.0)) * quad);      // distributions.py:132-137
        if (s_bad) lp = CUDART_NAN;
        a.logpts[(long)c * a.logpts_sc + a.out_ofs + t] = lp;
    }
    }   // item loop
}

// ---------------------------------------------------------------------------------------------------------
// standalone multivariate_normal_chol over residuals [B, nt, ns] (beat/models/distributions.py:72-140)
// ---------------------------------------------------------------------------------------------------------
struct MisfitArgs {
    int B, nt, ns;
    const double* resid;                  // [B, nt, ns]  (explicit residuals), or nullptr when `partial` is given
    const double* partial; int nchunk;    // [B, nt, nchunk, ns] partial synthetics of gf_stack_chunk_kernel
    const double* data;                   // [nt, ns] (with `partial`)
    const unsigned char* chain_bad;       // optional [B]
    const double* hyp; long hyp_sc; const int* hyper_idx;
    int misfit_mode; int bw; int dense_upper;
    const double* W; const double* slog_pdet; const int* nsamp;
    double* logpts; long logpts_sc; int out_ofs;
};

__global__ void __launch_bounds__(kStackThreads) misfit_kernel(MisfitArgs a)
{
    extern __shared__ double resid[];
    __shared__ double red_q[kStackWarps];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int c = blockIdx.x % a.B, t = blockIdx.x / a.B;
    const int ns = a.ns;
    if (a.partial) {
        // synth = sum of the patch-chunk partials in fixed chunk order (deterministic); residual = data - synth
        const double* pp = a.partial + ((long)c * a.nt + t) * a.nchunk * ns;
        for (int k = tid; k < ns; k += kStackThreads) {
            double s = __ldcs(pp + k);
            for (int j = 1; j < a.nchunk; ++j) s += __ldcs(pp + (long)j * ns + k);
            resid[k] = a.data[(long)t * ns + k] - s;                                     // seismic.py:1332
        }
    } else {
        const double* r = a.resid + ((long)c * a.nt + t) * ns;
#pragma unroll 8
        for (int k = tid; k < ns; k += kStackThreads) resid[k] = __ldg(r + k);      // several loads in flight per thread
    }
    __syncthreads();
    double q = 0.0;
    if (a.misfit_mode == MISFIT_DIAG) {
        const double* Wt = a.W + (long)t * ns;
        for (int k = tid; k < ns; k += kStackThreads) { const double z = Wt[k] * resid[k]; q = fma(z, z, q); }
    } else if (a.misfit_mode == MISFIT_BAND) {
        const double* Wt = a.W + (long)t * (a.bw + 1) * ns;
        for (int k = tid; k < ns; k += kStackThreads) {
            double z = 0.0;
            const int jmax = min(a.bw, ns - 1 - k);
            for (int j = 0; j <= jmax; ++j) z = fma(Wt[(long)j * ns + k], resid[k + j], z);
            q = fma(z, z, q);
        }
    } else {
        const double* Wt = a.W + (long)t * ns * ns;
        for (int k = tid; k < ns; k += kStackThreads) {
            double z = 0.0;
            for (int j = a.dense_upper ? k : 0; j < ns; ++j) z = fma(Wt[(long)j * ns + k], resid[j], z);
            q = fma(z, z, q);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    if (lane == 0) red_q[warp] = q;
    __syncthreads();
    if (tid == 0) {
        double quad = red_q[0];
        for (int w = 1; w < kStackWarps; ++w) quad += red_q[w];
        const double hp = a.hyp[(long)c * a.hyp_sc + a.hyper_idx[t]];
        const double M = (double)(short)a.nsamp[t];
        const double norm = M * (2.0 * hp + 1.8378770664093453);
        double lp = (-0.5) * (a.slog_pdet[t] + norm + (1.0 / exp(hp * 2.0)) * quad);
        if (a.chain_bad && a.chain_bad[c]) lp = CUDART_NAN;
        a.logpts[(long)c * a.logpts_sc + a.out_ofs + t] = lp;
    }
}

// Short traces (ns <= 32 * NI) with diagonal or banded weights: one WARP per (chain, target), four items per CTA, no
// block-level barrier.  A lane owns samples lane, lane + 32, ...; the loads of all its samples of up to four chunk
// partials are in flight together (the pass is a pure HBM stream of the partials: 1.7 GB at C3 / 4000 chains).  Same
// per-sample arithmetic as misfit_kernel (partials summed in chunk order, banded dot product in j order); only the
// order in which the squared terms are added differs (deterministic).
constexpr int kMisfitWarpMaxNs = 256;

template <int NI>
__global__ void __launch_bounds__(kStackThreads) misfit_warp_kernel(MisfitArgs a)
{
    __shared__ double rs[kStackWarps][32 * NI + 8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long item = (long)blockIdx.x * kStackWarps + warp;
    if (item >= (long)a.B * a.nt) return;
    const int c = (int)(item % a.B), t = (int)(item / a.B);
    const int ns = a.ns;
    double* r = rs[warp];
    double s[NI];
    if (a.partial) {
        const double* pp = a.partial + ((long)c * a.nt + t) * a.nchunk * ns;
#pragma unroll
        for (int i = 0; i < NI; ++i) { const int k = lane + 32 * i; s[i] = (k < ns) ? __ldcs(pp + k) : 0.0; }
#pragma unroll 4
        for (int j = 1; j < a.nchunk; ++j) {
            double v[NI];
#pragma unroll
            for (int i = 0; i < NI; ++i) { const int k = lane + 32 * i; v[i] = (k < ns) ? __ldcs(pp + (long)j * ns + k) : 0.0; }
#pragma unroll
            for (int i = 0; i < NI; ++i) s[i] += v[i];
        }
#pragma unroll
        for (int i = 0; i < NI; ++i) { const int k = lane + 32 * i; if (k < ns) r[k] = a.data[(long)t * ns + k] - s[i]; }   // seismic.py:1332
    } else {
        const double* rr = a.resid + ((long)c * a.nt + t) * ns;
#pragma unroll
        for (int i = 0; i < NI; ++i) { const int k = lane + 32 * i; if (k < ns) r[k] = __ldg(rr + k); }
    }
    __syncwarp();
    double q = 0.0;
    if (a.misfit_mode == MISFIT_DIAG) {
        const double* Wt = a.W + (long)t * ns;
#pragma unroll
        for (int i = 0; i < NI; ++i) { const int k = lane + 32 * i; if (k < ns) { const double z = Wt[k] * r[k]; q = fma(z, z, q); } }
    } else {
        const double* Wt = a.W + (long)t * (a.bw + 1) * ns;
#pragma unroll
        for (int i = 0; i < NI; ++i) {
            const int k = lane + 32 * i;
            if (k < ns) {
                double z = 0.0;
                const int jmax = min(a.bw, ns - 1 - k);
                for (int j = 0; j <= jmax; ++j) z = fma(Wt[(long)j * ns + k], r[k + j], z);
                q = fma(z, z, q);
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    if (lane == 0) {
        const double hp = a.hyp[(long)c * a.hyp_sc + a.hyper_idx[t]];
        const double M = (double)(short)a.nsamp[t];
        const double norm = M * (2.0 * hp + 1.8378770664093453);
        double lp = (-0.5) * (a.slog_pdet[t] + norm + (1.0 / exp(hp * 2.0)) * q);
        if (a.chain_bad && a.chain_bad[c]) lp = CUDART_NAN;
        a.logpts[(long)c * a.logpts_sc + a.out_ofs + t] = lp;
    }
}

// ---------------------------------------------------------------------------------------------------------
// Patch-chunked stacking: one WARP per (target, patch-chunk, chain) item, items ordered (t, chunk, c) so that all
// warps resident at a time gather from the same few (t, p) library blocks (chunk * nvar * ndur*nst*ld*sizeof(T)
// bytes, ~26 MB at C3 with 25-patch chunks) which therefore stay in L2 while every chain streams through them:
// HBM traffic tends to one pass over the touched library instead of one pass per ~wave of chains.  No block-level
// barrier at all (plan, stream, store are warp-private).  Partial synthetics go to a [B, nt, nchunk, ns] f64
// scratch; misfit_kernel sums them in chunk order and finishes residual + misfit.
// ---------------------------------------------------------------------------------------------------------
constexpr int kChunkMax = 64;            // patches per chunk (a lane plans patches lane, lane + 32)
constexpr int kChunkWarps = 4;

struct ChunkArgs {
    StackArgs s;                          // library, axes and per-chain inputs as for the fused kernel
    int chunk;                            // patches per chunk (<= kChunkMax)
    int nchunk;
    uint32_t zero_mask;                   // always 0; opaque to the compiler (load-scheduling dependence, see the kernel)
    double* partial;                      // [B, nt, nchunk, ns]
    const void* plan_cache;               // optional PatchPlan<T,K,NVAR> [B, np] of target 0 (plan_cache_kernel), or nullptr
};

// Without station corrections the plan of a (chain, patch) -- library indices and weights -- is the same for every target
// (start times do not depend on the target, seismic.py:1283-1296; only the target's block offset differs), yet every
// (target, chunk, chain) warp would recompute it: 64 times at C3.  This kernel computes it once per (chain, patch) for
// target 0; the chunk kernel then loads its patches' plans (one coalesced read) and adds the target offset.  A violating
// tap is counted once per target (as the per-target plans would) and its weights become NaN so that the partial sums and
// the logpt of that chain do.
template <typename T, int K, int NVAR>
__global__ void __launch_bounds__(128) plan_cache_kernel(ChunkArgs ca, PatchPlan<T, K, NVAR>* __restrict__ cache)
{
    const StackArgs& a = ca.s;
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long)a.B * a.np) return;
    const int c = (int)(idx / a.np), p = (int)(idx % a.np);
    PatchPlan<T, K, NVAR> pl;
    const bool viol = make_patch_plan<T, K, NVAR>(a, c, 0, p, a.dur[(long)c * a.dur_sc + p], a.st[(long)c * a.st_sc + p], 0.0, pl);
    if (viol) {
        atomicAdd(a.violations, (unsigned long long)a.nt);
#pragma unroll
        for (int q = 0; q < K * NVAR; ++q) pl.w[q] = (T)CUDART_NAN;
    }
    cache[idx] = pl;
}

// MINB: CTAs per SM the register allocation must allow (__launch_bounds__): 5 leaves the multilinear kernels their natural
// ~96 registers (16 row loads of two patches in flight per lane), 6 / 7 cap them at 80 / 72 (more resident warps, fewer
// loads in flight each); selected at run time (BEATGPU_CHUNK_OCC), default from measurements (profiles/README.md).
template <typename T, int K, int NVAR, int MINB>
__global__ void __launch_bounds__(kChunkWarps * 32, MINB)
gf_stack_chunk_kernel(ChunkArgs ca)
{
    using Plan = PatchPlan<T, K, NVAR>;
    __shared__ Plan plan_s[kChunkWarps][kChunkMax];
    const StackArgs& a = ca.s;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const long n_items = (long)a.nt * ca.nchunk * a.B;
    const long item = (long)blockIdx.x * kChunkWarps + warp;
    if (item >= n_items) return;
    const int c = (int)(item % a.B);
    const int kc = (int)((item / a.B) % ca.nchunk);
    const int t = (int)(item / ((long)a.B * ca.nchunk));
    Plan* plan = plan_s[warp];

    const int p0 = kc * ca.chunk;
    const int pn = min(ca.chunk, a.np - p0);
    const double* dur = a.dur + (long)c * a.dur_sc;
    const double* st = a.st + (long)c * a.st_sc + (long)t * a.st_st;
    double corr = 0.0;
    if (a.corr) corr = a.corr[(long)c * a.corr_sc + a.station_idx[t]];
    // ---- plan (one lane per patch; same arithmetic as the fused kernel / ffi/base.py:506-517,553-564,676-679)
    bool viol = false;
    if (ca.plan_cache) {
        const Plan* pc = reinterpret_cast<const Plan*>(ca.plan_cache) + (long)c * a.np + p0;
        const uint32_t toff = (uint32_t)((long)t * a.np * a.ndur * a.nst) * (uint32_t)(a.ld * (long)sizeof(T) / 16);   // target block, 16-byte units
        for (int i = lane; i < pn; i += 32) {
            Plan pl = pc[i];
#pragma unroll
            for (int k = 0; k < 4; ++k) pl.off[k] += toff;
            plan[i] = pl;
        }
    } else {
        for (int i = lane; i < pn; i += 32) {
            const int p = p0 + i;
            Plan pl;
            const bool v = make_patch_plan<T, K, NVAR>(a, c, t, p, dur[p], st[p], corr, pl);
            if (v) atomicAdd(a.violations, 1ULL);
            viol = viol || v;
            plan[i] = pl;
        }
    }
    const bool any_viol = __any_sync(0xffffffffu, viol);
    __syncwarp();

    double* out = ca.partial + (((long)c * a.nt + t) * ca.nchunk + kc) * a.ns;
    for (int s0 = 0; s0 < a.ns; s0 += kWindow) {
        const int wlen = min(kWindow, a.ns - s0);
        const int nvec = (wlen * (int)sizeof(T) + 15) / 16;
        double acc[4] = {0.0, 0.0, 0.0, 0.0};
        // Every lane loads unconditionally: lanes past the end of the row re-read its last vector (valid memory, the
        // value is never stored), so the row loop carries no predicates and no zero fills; the address of a row is one
        // IMAD.WIDE (plan offset in 16-byte units) on a per-lane base pointer.
        if (sizeof(T) == 4) {
            // rows of PF patches in flight per lane.  Measured at C3, nearest neighbour: PF = 2 / 4 / 8 -> 789 k / 840 k /
            // 732 k evals/s (depth vs registers/occupancy); multilinear already has 8*NVAR loads per patch.
            constexpr int PF = (K == 1) ? 4 : 2;
            const char* gb[NVAR];
#pragma unroll
            for (int v = 0; v < NVAR; ++v)
                gb[v] = reinterpret_cast<const char*>(a.G[v]) + (size_t)s0 * sizeof(float) + (size_t)min(lane, nvec - 1) * 16u;
            int i = 0;
            for (; i + PF <= pn; i += PF) {
                float4 g[PF][K * NVAR];
#pragma unroll
                for (int u = 0; u < PF; ++u)
#pragma unroll
                    for (int v = 0; v < NVAR; ++v)
#pragma unroll
                        for (int k = 0; k < K; ++k)
                            g[u][v * K + k] = __ldg(reinterpret_cast<const float4*>(row_ptr(gb[v], plan[i + u].off[k])));
                // Every row load of the group must be in flight before the first product is formed: left alone, ptxas pairs
                // each load with its use to save registers and the in-order issue then stalls on the first product with
                // two loads outstanding instead of sixteen.  The partial sums therefore START from a zero that is data
                // dependent on all loads (OR of one word of each, masked with a kernel argument that is 0): exact (+0.0f),
                // a handful of LOP3s, and no product can issue before the last load has.
                uint32_t dep = 0u;
#pragma unroll
                for (int u = 0; u < PF; ++u)
#pragma unroll
                    for (int q = 0; q < K * NVAR; ++q) dep |= __float_as_uint(g[u][q].x);
                const float zero = __uint_as_float(dep & ca.zero_mask);
                // f32 library: products and the PF*K*NVAR-term partial sum in f32 (FFMA pipe), one f32->f64 conversion per
                // element per patch group; the running sum over patches stays f64.
                float4 part = make_float4(zero, zero, zero, zero);
#pragma unroll
                for (int u = 0; u < PF; ++u)
#pragma unroll
                    for (int q = 0; q < K * NVAR; ++q) {
                        const float w = plan[i + u].w[q];
                        part.x = fmaf(w, g[u][q].x, part.x);
                        part.y = fmaf(w, g[u][q].y, part.y);
                        part.z = fmaf(w, g[u][q].z, part.z);
                        part.w = fmaf(w, g[u][q].w, part.w);
                    }
                acc[0] += (double)part.x;
                acc[1] += (double)part.y;
                acc[2] += (double)part.z;
                acc[3] += (double)part.w;
            }
            if (i < pn) {                                        // remainder (< PF patches): one group, same arithmetic
                float4 part = make_float4(0.f, 0.f, 0.f, 0.f);
                for (; i < pn; ++i) {
                    float4 g[K * NVAR];
#pragma unroll
                    for (int v = 0; v < NVAR; ++v)
#pragma unroll
                        for (int k = 0; k < K; ++k)
                            g[v * K + k] = __ldg(reinterpret_cast<const float4*>(row_ptr(gb[v], plan[i].off[k])));
#pragma unroll
                    for (int q = 0; q < K * NVAR; ++q) {
                        const float w = plan[i].w[q];
                        part.x = fmaf(w, g[q].x, part.x);
                        part.y = fmaf(w, g[q].y, part.y);
                        part.z = fmaf(w, g[q].z, part.z);
                        part.w = fmaf(w, g[q].w, part.w);
                    }
                }
                acc[0] += (double)part.x;
                acc[1] += (double)part.y;
                acc[2] += (double)part.z;
                acc[3] += (double)part.w;
            }
        } else {
            // f64 storage: a window of 128 samples = 64 16-byte vectors; lane reads vectors `lane` and `lane+32`
            const char* gb0[NVAR];
            const char* gb1[NVAR];
#pragma unroll
            for (int v = 0; v < NVAR; ++v) {
                const char* b = reinterpret_cast<const char*>(a.G[v]) + (size_t)s0 * sizeof(double);
                gb0[v] = b + (size_t)min(lane, nvec - 1) * 16u;
                gb1[v] = b + (size_t)min(lane + 32, nvec - 1) * 16u;
            }
            for (int i = 0; i < pn; ++i) {
                double2 g0[K * NVAR], g1[K * NVAR];
#pragma unroll
                for (int v = 0; v < NVAR; ++v)
#pragma unroll
                    for (int k = 0; k < K; ++k) {
                        g0[v * K + k] = __ldg(reinterpret_cast<const double2*>(row_ptr(gb0[v], plan[i].off[k])));
                        g1[v * K + k] = __ldg(reinterpret_cast<const double2*>(row_ptr(gb1[v], plan[i].off[k])));
                    }
                // all 2*K*NVAR loads in flight before the first FMA (see the f32 branch): the first weight carries a data
                // dependence on every load (its high word OR 0), and each accumulator chain starts with that weight
                uint32_t dep = 0u;
#pragma unroll
                for (int q = 0; q < K * NVAR; ++q) dep |= (uint32_t)__double2hiint(g0[q].x) | (uint32_t)__double2hiint(g1[q].x);
                dep &= ca.zero_mask;
#pragma unroll
                for (int q = 0; q < K * NVAR; ++q) {
                    double w = plan[i].w[q];
                    if (q == 0) w = __hiloint2double(__double2hiint(w) | (int)dep, __double2loint(w));
                    acc[0] = fma(w, g0[q].x, acc[0]);
                    acc[1] = fma(w, g0[q].y, acc[1]);
                    acc[2] = fma(w, g1[q].x, acc[2]);
                    acc[3] = fma(w, g1[q].y, acc[3]);
                }
            }
        }
        if (any_viol) acc[0] = acc[1] = acc[2] = acc[3] = CUDART_NAN;
        if ((a.ns & 1) == 0) {
            // (e0,e1) and (e2,e3) are adjacent samples for both storage types: two 16-byte stores per lane
#pragma unroll
            for (int e = 0; e < 4; e += 2) {
                const int sidx = slot_sample<T>(lane, e);
                // streaming store: the partials are read once by the misfit pass and must not evict library rows from L2
                if (sidx + 1 < wlen) __stcs(reinterpret_cast<double2*>(out + s0 + sidx), make_double2(acc[e], acc[e + 1]));
                else if (sidx < wlen) out[s0 + sidx] = acc[e];
            }
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int sidx = slot_sample<T>(lane, e);
                if (sidx < wlen) out[s0 + sidx] = acc[e];
            }
        }
    }
}

}  // namespace beatgpu
