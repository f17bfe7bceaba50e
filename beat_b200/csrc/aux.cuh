// aux.cuh -- small kernels around the stack kernel: library repack on upload, geodetic static stack + misfit,
//            laplacian smoothing prior, per-chain sum of the per-dataset log-likelihoods.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace beatgpu {

// ---------------------------------------------------------------------------------------------------------
// GF-library repack: src rows [n_rows, ns] (f32 or f64) -> dst rows [n_rows, ld] (f32 or f64), zero padded.
// (on-disk layout of the reference's .traces.npy, beat/ffi/base.py:364-373, is kept; only the row stride and
//  the element type change)
// ---------------------------------------------------------------------------------------------------------
template <typename TS, typename TD>
__global__ void repack_rows_kernel(const TS* __restrict__ src, TD* __restrict__ dst, long n_rows, int ns, long ld)
{
    const long total = n_rows * ld;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const long r = i / ld;
        const int k = (int)(i - r * ld);
        dst[i] = (k < ns) ? (TD)src[r * ns + k] : (TD)0;
    }
}

// ---------------------------------------------------------------------------------------------------------
// Geodetic static composite (beat/models/geodetic.py:1065-1081):
//   mu = sum_var G_var^T u_var ; r = (data - mu) * odw ; per dataset: logpt = mvn_chol(U_d, r[lo:hi])
// one CTA per (chain, dataset); G_var is [np, nobs] row-major so threads over observations read coalesced.
// ---------------------------------------------------------------------------------------------------------
struct GeoArgs {
    int B, np, nobs, ndatasets, nvar;
    const double* G[BEATGPU_MAX_SLIPVARS];           // [np, nobs]
    const double* slip[BEATGPU_MAX_SLIPVARS]; long slip_sc[BEATGPU_MAX_SLIPVARS];
    const double* data; const double* odw;           // [nobs]
    const int* lo; const int* hi;                    // [ndatasets]
    const double* UT; const long* UT_ofs;            // per dataset transposed weights UT[j*n + k] = U[k][j], offsets in doubles
    const int* upper;                                // [ndatasets] 1 if U is upper triangular
    const double* slog_pdet; const int* nsamp; const int* hyper_idx;
    const double* hyp; long hyp_sc;
    double* logpts; long logpts_sc; int out_ofs;
    int max_n;                                       // largest dataset (shared memory sizing)
};

__global__ void __launch_bounds__(256) geodetic_kernel(GeoArgs a)
{
    extern __shared__ double sm[];                   // [nvar*np] slips, then [max_n] residual
    __shared__ double red_q[8];
    double* su = sm;
    double* resid = sm + (size_t)a.nvar * a.np;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int c = blockIdx.x % a.B, d = blockIdx.x / a.B;
    for (int v = 0; v < a.nvar; ++v)
        for (int p = tid; p < a.np; p += blockDim.x) su[v * a.np + p] = a.slip[v][(long)c * a.slip_sc[v] + p];
    __syncthreads();
    const int lo = a.lo[d], n = a.hi[d] - lo;
    for (int k = tid; k < n; k += blockDim.x) {
        double mu = 0.0;
        for (int v = 0; v < a.nvar; ++v) {
            const double* Gv = a.G[v] + lo + k;
            double m = 0.0;
            for (int p = 0; p < a.np; ++p) m = fma(Gv[(long)p * a.nobs], su[v * a.np + p], m);
            mu += m;                                                                   // geodetic.py:1065-1070
        }
        resid[k] = (a.data[lo + k] - mu) * a.odw[lo + k];                              // geodetic.py:1072-1074
    }
    __syncthreads();
    const double* UT = a.UT + a.UT_ofs[d];
    const int upper = a.upper[d];
    double q = 0.0;
    for (int k = tid; k < n; k += blockDim.x) {
        double z = 0.0;
        for (int j = upper ? k : 0; j < n; ++j) z = fma(UT[(long)j * n + k], resid[j], z);
        q = fma(z, z, q);
    }
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    if (lane == 0) red_q[warp] = q;
    __syncthreads();
    if (tid == 0) {
        double quad = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) quad += red_q[w];
        const double hp = a.hyp[(long)c * a.hyp_sc + a.hyper_idx[d]];
        const double M = (double)(short)a.nsamp[d];
        const double norm = M * (2.0 * hp + 1.8378770664093453);
        a.logpts[(long)c * a.logpts_sc + a.out_ofs + d] = (-0.5) * (a.slog_pdet[d] + norm + (1.0 / exp(hp * 2.0)) * quad);
    }
}

// ---------------------------------------------------------------------------------------------------------
// Laplacian smoothing prior (beat/models/laplacian.py:88-96,128-139): one CTA per chain,
//   sum over slip vars of -1/2 ( -sdet + np (log 2pi + 2h) + exp(-2h) |L u|^2 )
// LT is L transposed (LT[j*np + i] = L[i][j]) so that thread i reads coalesced.
// ---------------------------------------------------------------------------------------------------------
struct LapArgs {
    int B, np, nvar;
    const double* LT; double sdet; int hyper_idx;
    const double* slip[BEATGPU_MAX_SLIPVARS]; long slip_sc[BEATGPU_MAX_SLIPVARS];
    const double* hyp; long hyp_sc;
    double* logpts; long logpts_sc; int out_ofs;
};

__global__ void __launch_bounds__(256) laplacian_kernel(LapArgs a)
{
    extern __shared__ double su[];                   // [np]
    __shared__ double red_q[8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int c = blockIdx.x;
    const double hp = a.hyp[(long)c * a.hyp_sc + a.hyper_idx];
    double total = 0.0;
    for (int v = 0; v < a.nvar; ++v) {
        __syncthreads();
        for (int p = tid; p < a.np; p += blockDim.x) su[p] = a.slip[v][(long)c * a.slip_sc[v] + p];
        __syncthreads();
        double q = 0.0;
        for (int i = tid; i < a.np; i += blockDim.x) {
            double z = 0.0;
            for (int j = 0; j < a.np; ++j) z = fma(a.LT[(long)j * a.np + i], su[j], z);
            q = fma(z, z, q);
        }
        for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
        if (lane == 0) red_q[warp] = q;
        __syncthreads();
        if (tid == 0) {
            double e = 0.0;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) e += red_q[w];
            total += (-0.5) * (-a.sdet + ((double)a.np * (1.8378770664093453 + 2.0 * hp)) + (1.0 / exp(hp * 2.0) * e));
        }
    }
    if (tid == 0) a.logpts[(long)c * a.logpts_sc + a.out_ofs] = total;
}

// like[c] = sum_j logpts[c][j]   (beat/models/problems.py:228-247)
__global__ void sum_like_kernel(const double* __restrict__ logpts, double* __restrict__ like, int B, int n_out)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= B) return;
    double s = 0.0;
    for (int j = 0; j < n_out; ++j) s += logpts[(long)c * n_out + j];
    like[c] = s;
}

// ---------------------------------------------------------------------------------------------------------
// Weight matrices that are already on the device (per-stage covariance update, beat/models/seismic.py:1509-1534):
// structure detection and repacking into the layouts the misfit kernels read, without a host round trip.
//   pass 1: max |U_t| per target;  pass 2: any entry below the diagonal above the threshold? largest super-diagonal
//   offset above the threshold?  (same rules as the host path of beatgpu_update_weights)
__global__ void weights_absmax_kernel(const double* __restrict__ U, int ns, double* __restrict__ amax, int* __restrict__ has_nan)
{
    __shared__ double red[32];
    const int t = blockIdx.x;
    const double* Ut = U + (long)t * ns * ns;
    double m = 0.0;
    bool nan = false;
    for (long i = threadIdx.x; i < (long)ns * ns; i += blockDim.x) {
        const double v = fabs(Ut[i]);
        if (!(v == v)) nan = true;
        m = fmax(m, v);
    }
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    if (nan) atomicExch(has_nan, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) m = fmax(m, red[w]);
        amax[t] = m;
    }
}

__global__ void weights_structure_kernel(const double* __restrict__ U, int ns, const double* __restrict__ amax, double band_rtol,
                                         int* __restrict__ lower, int* __restrict__ bw)
{
    const int t = blockIdx.x;
    const double* Ut = U + (long)t * ns * ns;
    const double thr = band_rtol * amax[t];
    int my_bw = 0, my_lower = 0;
    for (long e = threadIdx.x; e < (long)ns * ns; e += blockDim.x) {
        const int i = (int)(e / ns), j = (int)(e % ns);
        if (fabs(Ut[e]) > thr) {
            if (j < i) my_lower = 1;
            else my_bw = max(my_bw, j - i);
        }
    }
    if (my_lower) atomicExch(lower, 1);
    if (my_bw) atomicMax(bw, my_bw);
}

// mode 0: diagonal W[t][k] = U[t][k][k]; mode 1: band W[t][j][k] = U[t][k][k+j] (0 past the edge); mode 2: dense transposed
// W[t][j][k] = U[t][k][j]
__global__ void weights_repack_kernel(const double* __restrict__ U, double* __restrict__ W, int nt, int ns, int mode, int bw)
{
    const long n = mode == 0 ? (long)nt * ns : (mode == 1 ? (long)nt * (bw + 1) * ns : (long)nt * ns * ns);
    for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long)gridDim.x * blockDim.x) {
        if (mode == 0) {
            const int t = (int)(e / ns), k = (int)(e % ns);
            W[e] = U[((long)t * ns + k) * ns + k];
        } else if (mode == 1) {
            const int k = (int)(e % ns), j = (int)((e / ns) % (bw + 1)), t = (int)(e / ((long)ns * (bw + 1)));
            W[e] = (k + j < ns) ? U[((long)t * ns + k) * ns + (k + j)] : 0.0;
        } else {
            const int k = (int)(e % ns), j = (int)((e / ns) % ns), t = (int)(e / ((long)ns * ns));
            W[e] = U[((long)t * ns + k) * ns + j];
        }
    }
}

}  // namespace beatgpu
