// gemm.cuh -- FP64 tensor-core (DMMA, mma.sync.m8n8k4.f64) tiles for the two places where the hot path IS a GEMM
// once it is batched over chains: the geodetic static forward model and the dense-covariance misfit.
//
//   mu[nobs, B]   = sum_var G_var^T [nobs, np] . slip_var [np, B]          (beat/ffi/base.py:292-305, geodetic.py:1065-1070)
//   R  [nobs, B]  = (data - mu) * odw                                       (geodetic.py:1072-1074)
//   Z_d[n_d, B]   = U_d [n_d, n_d] . R[lo_d:hi_d, B]   ;  quad[c, d] = |Z_d[:, c]|^2    (distributions.py:128,136)
//
// The reference evaluates these as one gemv per chain; with all chains of a population in one call they are
// [500 x 200] x [200 x B] and [500 x 500] x [500 x B] products (config 4: full non-Toeplitz data covariance).  f64
// throughout (the residual is a difference of nearly equal numbers -- tf32/bf16 tensor formats are not accurate
// enough), IEEE FMA inside the MMA.  Tile 64 x 64 per CTA, 8 warps (each 16 x 32 = 2 x 4 m8n8 fragments), K step 16
// through shared memory.  Upper-triangular U skips the K tiles left of the diagonal.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

namespace beatgpu {

constexpr int kGemmTile = 64;
constexpr int kGemmK = 16;
constexpr int kGemmThreads = 256;

__device__ __forceinline__ void dmma_m8n8k4(double& c0, double& c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

struct GemmArgs {
    int M, N, K;                          // C[M, N] = A[M, K] B[K, N]
    int n_parts;                          // K is the concatenation of n_parts blocks (slip variables), each with its own A / B
    const double* A[BEATGPU_MAX_SLIPVARS]; long a_sm, a_sk;     // A(m, k) = A[p][m*a_sm + k*a_sk]
    const double* B[BEATGPU_MAX_SLIPVARS]; long b_sk, b_sn[BEATGPU_MAX_SLIPVARS];   // B(k, n) = B[p][k*b_sk + n*b_sn[p]]
    int upper;                            // A is upper triangular: skip k-tiles with k < m
    // epilogue 0: residual  R[n*ldr + m] = (data[m] - acc) * odw[m]
    const double* data; const double* odw; double* R; long ldr;
    // epilogue 1: partial column norms  qpart[(n*n_mtiles + mtile)] = sum over the tile's rows of acc^2
    double* qpart; int n_mtiles;
    // batching over blockIdx.z (e.g. the targets of a wavemap): element strides of A[0], B[0], qpart per batch index
    long a_batch, b_batch, q_batch;
};

template <int EPI>
__global__ void __launch_bounds__(kGemmThreads) dgemm_tile_kernel(GemmArgs g)
{
    __shared__ double As[kGemmK][kGemmTile + 4];
    __shared__ double Bs[kGemmK][kGemmTile + 4];
    __shared__ double colsum[4][kGemmTile];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int m0 = blockIdx.y * kGemmTile, n0 = blockIdx.x * kGemmTile;
    const int wm = (warp & 3) * 16, wn = (warp >> 2) * 32;
    const int gid = lane >> 2, tig = lane & 3;

    double acc[2][4][2];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    for (int part = 0; part < g.n_parts; ++part) {
        const double* Ap = g.A[part] + (long)blockIdx.z * g.a_batch;
        const double* Bp = g.B[part] + (long)blockIdx.z * g.b_batch;
        const long bsn = g.b_sn[part];
        for (int k0 = 0; k0 < g.K; k0 += kGemmK) {
            if (g.upper && k0 + kGemmK <= m0) continue;            // whole tile strictly left of the diagonal: zeros
            // stage A[m0.., k0..] and B[k0.., n0..] (zero padded)
            // thread->element mapping follows the contiguous axis of each operand (coalesced global reads)
            for (int e = tid; e < kGemmK * kGemmTile; e += kGemmThreads) {
                const int kk = (g.a_sm == 1) ? e / kGemmTile : e % kGemmK;
                const int mm = (g.a_sm == 1) ? e % kGemmTile : e / kGemmK;
                const int m = m0 + mm, k = k0 + kk;
                As[kk][mm] = (m < g.M && k < g.K) ? Ap[(long)m * g.a_sm + (long)k * g.a_sk] : 0.0;
            }
            for (int e = tid; e < kGemmK * kGemmTile; e += kGemmThreads) {
                const int kk = (g.b_sk == 1) ? e % kGemmK : e / kGemmTile;
                const int nn = (g.b_sk == 1) ? e / kGemmK : e % kGemmTile;
                const int n = n0 + nn, k = k0 + kk;
                Bs[kk][nn] = (n < g.N && k < g.K) ? Bp[(long)k * g.b_sk + (long)n * bsn] : 0.0;
            }
            __syncthreads();
#pragma unroll
            for (int kk = 0; kk < kGemmK; kk += 4) {
                double a[2], b[4];
#pragma unroll
                for (int i = 0; i < 2; ++i) a[i] = As[kk + tig][wm + i * 8 + gid];
#pragma unroll
                for (int j = 0; j < 4; ++j) b[j] = Bs[kk + tig][wn + j * 8 + gid];
#pragma unroll
                for (int i = 0; i < 2; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
            }
            __syncthreads();
        }
    }

    if (EPI == 0) {
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int m = m0 + wm + i * 8 + gid;
                    const int n = n0 + wn + j * 8 + tig * 2 + e;
                    if (m < g.M && n < g.N) g.R[(long)n * g.ldr + m] = (g.data[m] - acc[i][j][e]) * g.odw[m];
                }
    } else {
        // column sums of squares over this tile's 64 rows, in a fixed order (deterministic)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                double s = acc[0][j][e] * acc[0][j][e] + acc[1][j][e] * acc[1][j][e];   // rows gid and gid+8 of the warp tile
                s += __shfl_xor_sync(0xffffffffu, s, 4);                                   // over gid (lane bits 2..4)
                s += __shfl_xor_sync(0xffffffffu, s, 8);
                s += __shfl_xor_sync(0xffffffffu, s, 16);
                if (gid == 0) colsum[warp & 3][wn + j * 8 + tig * 2 + e] = s;
            }
        __syncthreads();
        if (tid < kGemmTile) {
            const int n = n0 + tid;
            if (n < g.N) g.qpart[(long)blockIdx.z * g.q_batch + (long)n * g.n_mtiles + blockIdx.y] = (colsum[0][tid] + colsum[1][tid]) + (colsum[2][tid] + colsum[3][tid]);
        }
    }
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
