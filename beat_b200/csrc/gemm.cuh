// gemm.cuh -- FP64 tensor-core (DMMA, mma.sync.m8n8k4.f64) tiles for the two places where the hot path IS a GEMM
// once it is batched over chains: the geodetic static forward model and the dense-covariance misfit.
//
//   mu[nobs, B]   = sum_var G_var^T [nobs, np] . slip_var [np, B]          (beat/ffi/base.py:292-305, geodetic.py:1065-1070)
//   R  [nobs, B]  = (data - mu) * odw                                       (geodetic.py:1072-1074)
//   Z_d[n_d, B]   = U_d [n_d, n_d] . R[lo_d:hi_d, B]   ;  quad[c, d] = |Z_d[:, c]|^2    (distributions.py:128,136)
//
// The reference evaluates these as one gemv per chain; with all chains of a population in one call they are
// [500 x 200] x [200 x B] and [500 x 500] x [500 x B] products (config 4: full non-Toeplitz data covariance), and
// [2048 x 2048] x [2048 x B] per dataset for a dense covariance at the config-2 trace length.  f64 throughout (the
// residual is a difference of nearly equal numbers -- tf32/bf16 tensor formats are not accurate enough), IEEE FMA
// inside the MMA.
//
// Kernel shape (round 2): CTA tile 128 (M) x 64 (N), 8 warps of 32 x 32 (4 x 4 m8n8 fragments), K step 16, operands
// staged by `cp.async` through a 3-stage shared-memory ring (one CTA barrier per K step, loads two steps ahead), 2 CTAs
// per SM.  A is M-contiguous in memory (weights are stored transposed: A(m, k) = A[k*a_sk + m]) and lands as As[k][m];
// B is K-contiguous (one chain's residual / slip vector per row: B(k, n) = B[n*b_sn + k]) and lands as Bs[n][k]; the
// paddings (130 / 24 doubles) make every fragment read a conflict-free LDS.128: a thread fetches TWO consecutive m of
// one k for A and TWO consecutive k of one n for B, i.e. the fragments use permuted row / k assignments (rows 2g, 2g+1
// instead of g, g+8; k = {0,2,4,6} then {1,3,5,7}) -- a GEMM is invariant under both as long as A and B agree on k and
// the epilogue knows the row of each accumulator.  Upper-triangular U: K steps left of the CTA's rows are never loaded,
// K steps left of a warp's rows are not multiplied.  Summation order is fixed (deterministic results).
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

namespace beatgpu {

constexpr int kGemmBM = 128;                             // CTA tile rows (M)
constexpr int kGemmBN = 64;                              // CTA tile columns (N = chains)
constexpr int kGemmK = 16;
constexpr int kGemmStages = 3;
constexpr int kGemmThreads = 256;
constexpr int kGemmLDA = kGemmBM + 2;                    // As[k][m]: 16-byte units of a quarter-warp's LDS.128 fall into 8 distinct banks groups
constexpr int kGemmLDB = kGemmK + 8;                     // Bs[n][k]
constexpr int kGemmStageDoubles = kGemmK * kGemmLDA + kGemmBN * kGemmLDB;
constexpr size_t kGemmSmemBytes = (size_t)kGemmStages * kGemmStageDoubles * sizeof(double);

__device__ __forceinline__ void dmma_m8n8k4(double& c0, double& c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// asynchronous global -> shared copy of BYTES (8 or 16); an invalid source is replaced by zeros (src-size 0)
template <int BYTES>
__device__ __forceinline__ void cp_async_zfill(double* smem_dst, const double* gsrc, bool valid)
{
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int n = valid ? BYTES : 0;
    if (BYTES == 16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" :: "r"(s), "l"(gsrc), "r"(n) : "memory");
    else             asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" :: "r"(s), "l"(gsrc), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" :: "n"(N) : "memory"); }

struct GemmArgs {
    int M, N, K;                          // C[M, N] = A[M, K] B[K, N]
    int n_parts;                          // K is the concatenation of n_parts blocks (slip variables), each with its own A / B
    const double* A[BEATGPU_MAX_SLIPVARS]; long a_sm, a_sk;     // A(m, k) = A[p][m*a_sm + k*a_sk], a_sm == 1
    const double* B[BEATGPU_MAX_SLIPVARS]; long b_sk, b_sn[BEATGPU_MAX_SLIPVARS];   // B(k, n) = B[p][k*b_sk + n*b_sn[p]], b_sk == 1
    int upper;                            // A is upper triangular: skip k-tiles with k < m
    // epilogue 0: residual  R[n*ldr + m] = (data[m] - acc) * odw[m]
    const double* data; const double* odw; double* R; long ldr;
    // epilogue 1: partial column norms  qpart[(n*n_mtiles + mtile)] = sum over the tile's rows of acc^2
    double* qpart; int n_mtiles;
    // batching over blockIdx.z (e.g. the targets of a wavemap): element strides of A[0], B[0], qpart per batch index
    long a_batch, b_batch, q_batch;
};

// EPI: epilogue (above).  VEC: doubles per cp.async (2 when every row start and extent is 16-byte aligned, else 1).
template <int EPI, int VEC>
__global__ void __launch_bounds__(kGemmThreads, 2) dgemm_tile_kernel(GemmArgs g)
{
    extern __shared__ __align__(16) double gemm_smem[];
    __shared__ double colsum[4][kGemmBN];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int m0 = blockIdx.y * kGemmBM, n0 = blockIdx.x * kGemmBN;
    const int wm = (warp & 3) * 32, wn = (warp >> 2) * 32;
    const int gid = lane >> 2, tig = lane & 3;

    // accumulator f = blk*2 + i holds row  m0 + wm + blk*16 + 2*gid + i ; columns n0 + wn + j*8 + 2*tig + e
    double acc[4][4][2];
#pragma unroll
    for (int f = 0; f < 4; ++f)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[f][j][0] = acc[f][j][1] = 0.0;

    const int tiles_per_part = (g.K + kGemmK - 1) / kGemmK;
    const int kt_begin = g.upper ? m0 / kGemmK : 0;                  // K steps entirely left of this CTA's rows hold zeros
    const int n_kt = g.n_parts * tiles_per_part - kt_begin;

    auto load_tile = [&](int kt, int stage) {
        const int part = (kt_begin + kt) / tiles_per_part, k0 = ((kt_begin + kt) % tiles_per_part) * kGemmK;
        const double* Ap = g.A[part] + (long)blockIdx.z * g.a_batch;
        const double* Bp = g.B[part] + (long)blockIdx.z * g.b_batch;
        const long bsn = g.b_sn[part];
        double* As = gemm_smem + (size_t)stage * kGemmStageDoubles;
        double* Bs = As + kGemmK * kGemmLDA;
        constexpr int a_per_row = kGemmBM / VEC, b_per_row = kGemmK / VEC;
#pragma unroll
        for (int c = tid; c < kGemmK * a_per_row; c += kGemmThreads) {
            const int kk = c / a_per_row, mm = (c % a_per_row) * VEC;
            const int m = m0 + mm, k = k0 + kk;
            const bool ok = m < g.M && k < g.K;
            cp_async_zfill<8 * VEC>(As + kk * kGemmLDA + mm, ok ? Ap + (long)k * g.a_sk + m : Ap, ok);
        }
#pragma unroll
        for (int c = tid; c < kGemmBN * b_per_row; c += kGemmThreads) {
            const int nn = c / b_per_row, kk = (c % b_per_row) * VEC;
            const int n = n0 + nn, k = k0 + kk;
            const bool ok = n < g.N && k < g.K;
            cp_async_zfill<8 * VEC>(Bs + nn * kGemmLDB + kk, ok ? Bp + (long)n * bsn + k : Bp, ok);
        }
    };

#pragma unroll
    for (int s = 0; s < kGemmStages - 1; ++s) {
        if (s < n_kt) load_tile(s, s);
        cp_async_commit();
    }
    for (int kt = 0; kt < n_kt; ++kt) {
        cp_async_wait<kGemmStages - 2>();                            // this thread's copies of step kt have landed
        __syncthreads();                                             // ... everybody's; and step kt-1 has been consumed
        if (kt + kGemmStages - 1 < n_kt) load_tile(kt + kGemmStages - 1, (kt + kGemmStages - 1) % kGemmStages);
        cp_async_commit();
        const int k0 = ((kt_begin + kt) % tiles_per_part) * kGemmK;
        if (g.upper && k0 + kGemmK <= m0 + wm) continue;             // left of this warp's rows: zeros (warp-uniform)
        const double* As = gemm_smem + (size_t)(kt % kGemmStages) * kGemmStageDoubles;
        const double* Bs = As + kGemmK * kGemmLDA;
#pragma unroll
        for (int kk = 0; kk < kGemmK; kk += 8) {
            double2 a[2][2], b[4];
#pragma unroll
            for (int blk = 0; blk < 2; ++blk)
#pragma unroll
                for (int kp = 0; kp < 2; ++kp)
                    a[blk][kp] = *reinterpret_cast<const double2*>(As + (kk + 2 * tig + kp) * kGemmLDA + wm + blk * 16 + 2 * gid);
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = *reinterpret_cast<const double2*>(Bs + (wn + j * 8 + gid) * kGemmLDB + kk + 2 * tig);
            // all 16 products of one k set, then the 16 of the other: consecutive DMMAs never touch the same accumulator
#pragma unroll
            for (int kp = 0; kp < 2; ++kp)
#pragma unroll
                for (int blk = 0; blk < 2; ++blk)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const double bk = kp ? b[j].y : b[j].x;
                        dmma_m8n8k4(acc[blk * 2 + 0][j][0], acc[blk * 2 + 0][j][1], a[blk][kp].x, bk);
                        dmma_m8n8k4(acc[blk * 2 + 1][j][0], acc[blk * 2 + 1][j][1], a[blk][kp].y, bk);
                    }
        }
    }
    cp_async_wait<0>();

    if (EPI == 0) {
#pragma unroll
        for (int f = 0; f < 4; ++f)
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int m = m0 + wm + (f >> 1) * 16 + 2 * gid + (f & 1);
                    const int n = n0 + wn + j * 8 + tig * 2 + e;
                    if (m < g.M && n < g.N) g.R[(long)n * g.ldr + m] = (g.data[m] - acc[f][j][e]) * g.odw[m];
                }
    } else {
        // column sums of squares over this tile's 128 rows, in a fixed order (deterministic)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                double s = (acc[0][j][e] * acc[0][j][e] + acc[1][j][e] * acc[1][j][e]) + (acc[2][j][e] * acc[2][j][e] + acc[3][j][e] * acc[3][j][e]);
                s += __shfl_xor_sync(0xffffffffu, s, 4);                                   // over gid (lane bits 2..4)
                s += __shfl_xor_sync(0xffffffffu, s, 8);
                s += __shfl_xor_sync(0xffffffffu, s, 16);
                if (gid == 0) colsum[warp & 3][wn + j * 8 + tig * 2 + e] = s;
            }
        __syncthreads();
        if (tid < kGemmBN) {
            const int n = n0 + tid;
            if (n < g.N) g.qpart[(long)blockIdx.z * g.q_batch + (long)n * g.n_mtiles + blockIdx.y] = (colsum[0][tid] + colsum[1][tid]) + (colsum[2][tid] + colsum[3][tid]);
        }
    }
}

// Host launcher: picks 16-byte copies when every operand row is 16-byte aligned and M, K are even, else 8-byte copies.
template <int EPI>
static cudaError_t launch_dgemm(const GemmArgs& g, int n_batch, cudaStream_t stream)
{
    if (g.a_sm != 1 || g.b_sk != 1) return cudaErrorInvalidValue;        // operand layouts this kernel is built for
    bool vec2 = (g.M % 2 == 0) && (g.K % 2 == 0) && (g.a_sk % 2 == 0) && (g.a_batch % 2 == 0) && (g.b_batch % 2 == 0);
    for (int p = 0; p < g.n_parts; ++p)
        vec2 = vec2 && ((uintptr_t)g.A[p] % 16 == 0) && ((uintptr_t)g.B[p] % 16 == 0) && (g.b_sn[p] % 2 == 0);
    static bool attr_set[2][64] = {};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64 && !attr_set[vec2][dev]) {
        e = vec2 ? cudaFuncSetAttribute(dgemm_tile_kernel<EPI, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGemmSmemBytes)
                 : cudaFuncSetAttribute(dgemm_tile_kernel<EPI, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGemmSmemBytes);
        if (e != cudaSuccess) return e;
        attr_set[vec2][dev] = true;
    }
    dim3 grid((g.N + kGemmBN - 1) / kGemmBN, (g.M + kGemmBM - 1) / kGemmBM, n_batch);
    if (vec2) dgemm_tile_kernel<EPI, 2><<<grid, kGemmThreads, kGemmSmemBytes, stream>>>(g);
    else      dgemm_tile_kernel<EPI, 1><<<grid, kGemmThreads, kGemmSmemBytes, stream>>>(g);
    return cudaGetLastError();
}

// logpt[c, d] from the per-row-tile partial norms (fixed summation order)
struct GeoFinishArgs {
    int B, n_mtiles;
    const double* qpart;                  // [B, n_mtiles]
    double slog_pdet; int nsamp; int hyper_idx;
    const double* hyp; long hyp_sc;
    double* logpts; long logpts_sc; int out_col;
};

__global__ void geodetic_finish_kernel(GeoFinishArgs a)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= a.B) return;
    double quad = 0.0;
    for (int j = 0; j < a.n_mtiles; ++j) quad += a.qpart[(long)c * a.n_mtiles + j];
    const double hp = a.hyp[(long)c * a.hyp_sc + a.hyper_idx];
    const double M = (double)(short)a.nsamp;
    const double norm = M * (2.0 * hp + 1.8378770664093453);
    a.logpts[(long)c * a.logpts_sc + a.out_col] = (-0.5) * (a.slog_pdet + norm + (1.0 / exp(hp * 2.0)) * quad);
}

// Laplacian prior from the per-row-tile partial norms of Z_v = L u_v (one GEMM per slip variable, epilogue 1):
// sum over slip vars of -1/2 ( -sdet + np (log 2pi + 2h) + exp(-2h) |L u_v|^2 )   (beat/models/laplacian.py:88-96,128-139)
struct LapFinishArgs {
    int B, n_mtiles, nvar, np;
    const double* qpart;                  // [nvar, B, n_mtiles]
    double sdet; int hyper_idx;
    const double* hyp; long hyp_sc;
    double* logpts; long logpts_sc; int out_col;
};

__global__ void laplacian_finish_kernel(LapFinishArgs a)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= a.B) return;
    const double hp = a.hyp[(long)c * a.hyp_sc + a.hyper_idx];
    double total = 0.0;
    for (int v = 0; v < a.nvar; ++v) {
        double e = 0.0;
        for (int j = 0; j < a.n_mtiles; ++j) e += a.qpart[((long)v * a.B + c) * a.n_mtiles + j];
        total += (-0.5) * (-a.sdet + ((double)a.np * (1.8378770664093453 + 2.0 * hp)) + (1.0 / exp(hp * 2.0) * e));
    }
    a.logpts[(long)c * a.logpts_sc + a.out_col] = total;
}

// ---------------------------------------------------------------------------------------------------------
// Dense-covariance seismic misfit as a GEMM (noise structures `non-toeplitz` / `import`: U_t is a full upper
// triangle): residuals R[t][c][:] = data[t] - sum_chunks partial, Z_t = U_t R_t on the FP64 tensor cores
// (batched over targets through blockIdx.z), then logpt per (chain, target).
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) residual_from_partials_kernel(const double* __restrict__ partial, const double* __restrict__ data,
                                                                      double* __restrict__ R, int B, int nt, int ns, int nchunk)
{
    const int c = blockIdx.x % B, t = blockIdx.x / B;
    const double* pp = partial + ((long)c * nt + t) * nchunk * ns;
    double* r = R + ((long)t * B + c) * ns;
    for (int k = threadIdx.x; k < ns; k += blockDim.x) {
        double s = pp[k];
        for (int j = 1; j < nchunk; ++j) s += pp[(long)j * ns + k];
        r[k] = data[(long)t * ns + k] - s;                                               // seismic.py:1332
    }
}

struct SeisFinishArgs {
    int B, nt, n_mtiles;
    const double* qpart;                  // [nt, B, n_mtiles]
    const double* slog_pdet; const int* nsamp; const int* hyper_idx;   // [nt]
    const double* hyp; long hyp_sc;
    const unsigned char* chain_bad;
    double* logpts; long logpts_sc; int out_ofs;
};

__global__ void seismic_finish_kernel(SeisFinishArgs a)
{
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long)a.B * a.nt) return;
    const int c = (int)(i % a.B), t = (int)(i / a.B);
    double quad = 0.0;
    for (int j = 0; j < a.n_mtiles; ++j) quad += a.qpart[((long)t * a.B + c) * a.n_mtiles + j];
    const double hp = a.hyp[(long)c * a.hyp_sc + a.hyper_idx[t]];
    const double M = (double)(short)a.nsamp[t];
    const double norm = M * (2.0 * hp + 1.8378770664093453);
    double lp = (-0.5) * (a.slog_pdet[t] + norm + (1.0 / exp(hp * 2.0)) * quad);
    if (a.chain_bad && a.chain_bad[c]) lp = CUDART_NAN;
    a.logpts[(long)c * a.logpts_sc + a.out_ofs + t] = lp;
}

}  // namespace beatgpu
