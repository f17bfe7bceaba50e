// sweep.cuh -- batched eikonal fast sweeping (rupture onset times), one lane group (<= a warp) per (chain, subfault).
//
// Replaces, for B chains at once, Sweeper.perform -> fast_sweep_ext.fast_sweep
// (reference: beat/pytensorf.py:443-500, beat/fast_sweeping/fast_sweep_ext.c:65-206).
//
// The reference converges with a loose tolerance (sum of squared updates <= 0.1, fast_sweep_ext.c:127,151,
// 198-201), so the Gauss-Seidel UPDATE ORDER determines the answer.  Each of the four directional sweeps is a
// sequential scan whose cell (r, c) reads the already-updated neighbours "behind" it and the not-yet-updated
// ones "ahead" of it; cells on one anti-diagonal of the scan direction neither read nor write each other, so
// relaxing anti-diagonals in scan order with all cells of a diagonal in parallel (one lane per cell) reproduces
// the sequential result bit for bit.  All arithmetic is IEEE double with explicit round-to-nearest intrinsics
// (no FMA contraction), sqrt for the reference's pow(x, 0.5) (see oracle/fast_sweep_port.c for the 1-ulp note),
// the same comparison directions (NaN from inf-inf falls through exactly as in C), and the convergence sum is
// accumulated serially in flat cell order.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>

namespace beatgpu {

constexpr int kSweepMaxOuterIters = 100000;   // the reference has no cap; this only guards the GPU against a hang

__device__ __forceinline__ double lesser(double x, double y) { return (x < y) ? x : y; }

// T, Tprev, fh, c2: per-chain shared arrays of n_rows*n_cols doubles.
//   fh[k] = slowness*h,  c2[k] = 2*slowness*slowness*h*h  (both sweep-invariant; evaluated left-to-right as in C)
// A warp relaxes 32 / W grids at once: a diagonal of a 10 x 20 fault has at most 10 cells, so W = 10 lanes serve one
// chain and three chains share the warp's instruction stream (lane `gl` of its group, `valid` = the group has a chain).
// Every group keeps its own convergence state; a converged group stops updating (bit-exact with the sequential code)
// while the others finish.  Returns the group's outer iteration count.
__device__ inline int group_fast_sweep(double* __restrict__ T, double* __restrict__ Tprev,
                                       const double* __restrict__ fh, const double* __restrict__ c2,
                                       int n_rows, int n_cols, int hyp_r, int hyp_c, int gl, int W, int leader, bool valid)
{
    const unsigned full = 0xffffffffu;
    const int n = valid ? n_rows * n_cols : 0;
    for (int k = gl; k < n; k += W) T[k] = CUDART_INF;
    __syncwarp();
    if (valid && gl == 0) T[hyp_r * n_cols + hyp_c] = 0.0;
    __syncwarp();

    double err = 1.0e6;
    int iters = 0;
    const int n_diag = valid ? n_rows + n_cols - 1 : 0;
    const int n_diag_warp = __reduce_max_sync(full, n_diag);
    bool go = valid;                                    // this group still iterates
    while (__any_sync(full, go)) {
        for (int k = gl; k < (go ? n : 0); k += W) Tprev[k] = T[k];
        __syncwarp();

#pragma unroll 1
        for (int s = 0; s < 4; ++s) {
            // scan directions of the four sweeps (fast_sweep_ext.c:159-196):
            // 0: rows up, cols up   1: rows down, cols up   2: rows down, cols down   3: rows up, cols down
            const bool r_up = (s == 0) || (s == 3);
            const bool c_up = (s < 2);
#pragma unroll 1
            for (int d = 0; d < n_diag_warp; ++d) {
                const int rlo = max(0, d - (n_cols - 1));
                const int rhi = go ? min(n_rows - 1, d) : -1;
                for (int rr = rlo + gl; rr <= rhi; rr += W) {
                    const int cc = d - rr;
                    const int r = r_up ? rr : n_rows - 1 - rr;
                    const int c = c_up ? cc : n_cols - 1 - cc;
                    const int rm = max(r - 1, 0), rp = min(r + 1, n_rows - 1);
                    const int cm = max(c - 1, 0), cp = min(c + 1, n_cols - 1);
                    const int k = r * n_cols + c;
                    const double a = lesser(T[rm * n_cols + c], T[rp * n_cols + c]);
                    const double b = lesser(T[r * n_cols + cm], T[r * n_cols + cp]);
                    const double cur = T[k];
                    const double f_h = fh[k];
                    const double diff = __dsub_rn(a, b);
                    // both candidates evaluated branch-free (lanes of a diagonal diverge otherwise); the selection
                    // keeps the reference's comparison `fabs(a-b) >= f*h` (false for NaN -> quadratic branch -> NaN).
                    // x / 2.0 == x * 0.5 bit for bit in IEEE arithmetic.
                    const double lin = __dadd_rn(lesser(a, b), f_h);
                    const double rad = __dsub_rn(c2[k], __dmul_rn(diff, diff));
                    const double quad = __dmul_rn(__dadd_rn(__dadd_rn(a, b), __dsqrt_rn(rad)), 0.5);
                    const double cand = (fabs(diff) >= f_h) ? lin : quad;
                    T[k] = (cand < cur) ? cand : cur;
                }
                __syncwarp();
            }
        }

        // err = sum_k (T[k]-Tprev[k])^2 in flat order (fast_sweep_ext.c:198-201)
        for (int k = gl; k < (go ? n : 0); k += W) {
            const double dlt = __dsub_rn(T[k], Tprev[k]);
            Tprev[k] = __dmul_rn(dlt, dlt);
        }
        __syncwarp();
        if (go && gl == 0) {
            double e = 0.0;
#pragma unroll 8
            for (int k = 0; k < n; ++k) e = __dadd_rn(e, Tprev[k]);
            err = e;
        }
        err = __shfl_sync(full, err, leader);
        if (go) ++iters;
        go = go && err > 0.1 && iters < kSweepMaxOuterIters;
        __syncwarp();
    }
    return iters;
}

struct SweepArgs {
    // per-subfault geometry (device arrays of length n_subfaults)
    const int* n_dip;
    const int* n_strike;
    const int* patch_ofs;       // first global patch index of each subfault
    const double* patch_size;
    int n_subfaults;
    int n_patches_total;
    int max_np_sf;              // shared-memory sizing
    int group_width;            // lanes per chain (>= longest grid diagonal when < 32); 32 / group_width chains share a warp
    int B;
    // inputs with per-chain strides (stride 0 = same value for all chains, i.e. a fixed variable)
    const double* vel;   long vel_stride;     // [np_total] velocities (slowness = 1/vel)  or slowness if is_slowness
    int is_slowness;
    const double* nuc_dip;    long nuc_dip_stride;     // [nsf] positions [km]  (ignored if idx given)
    const double* nuc_strike; long nuc_strike_stride;
    const int* nuc_dip_idx;      // optional explicit indices [B] (raw Sweeper entry); then single subfault `only_sf`
    const int* nuc_strike_idx;
    int only_sf;                 // >= 0: process just this subfault, inputs/outputs are [B, np_sf]
    const double* time; long time_stride;    // [nsf] or nullptr (no offset)
    double* t0;                  // out [B, np_total] (or [B, np_sf] when only_sf >= 0)
    int* n_iter;                 // optional [B * n_sf_processed]
    unsigned long long* violations;
    unsigned char* chain_bad;    // optional [B]: set to 1 when the nucleation index leaves the grid
};

// block = warps_per_block warps; each warp takes 32 / group_width (chain, subfault) work items.
__global__ void chain_sweep_kernel(SweepArgs a, int warps_per_block)
{
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int W = a.group_width, cpw = 32 / W;
    const int grp = lane / W, gl = lane - grp * W;
    const int n_sf_proc = (a.only_sf >= 0) ? 1 : a.n_subfaults;
    const long n_items = (long)a.B * n_sf_proc;
    const long item = ((long)blockIdx.x * warps_per_block + warp) * cpw + grp;
    const bool valid = grp < cpw && item < n_items;
    const int c = valid ? (int)(item / n_sf_proc) : 0;
    const int sf = (a.only_sf >= 0) ? a.only_sf : (valid ? (int)(item % n_sf_proc) : 0);

    double* T = smem + ((size_t)warp * cpw + min(grp, cpw - 1)) * 4 * a.max_np_sf;
    double* Tprev = T + a.max_np_sf;
    double* fh = Tprev + a.max_np_sf;
    double* c2 = fh + a.max_np_sf;

    const int nd = a.n_dip[sf], nstr = a.n_strike[sf];
    const int n = valid ? nd * nstr : 0;
    const double h = a.patch_size[sf];
    const int pofs = (a.only_sf >= 0) ? 0 : a.patch_ofs[sf];
    const int row_len = (a.only_sf >= 0) ? nd * nstr : a.n_patches_total;

    const double* vrow = a.vel + (long)c * a.vel_stride + pofs;
    for (int k = gl; k < n; k += W) {
        const double v = vrow[k];
        const double f = a.is_slowness ? v : __ddiv_rn(1.0, v);            // seismic.py:1264
        fh[k] = __dmul_rn(f, h);
        c2[k] = __dmul_rn(__dmul_rn(__dmul_rn(__dmul_rn(2.0, f), f), h), h);   // 2.0*f*f*h*h, left to right
    }

    int hr = 0, hc = 0;
    bool bad = false;
    if (valid) {
        if (a.nuc_dip_idx) {
            hr = a.nuc_dip_idx[c];
            hc = a.nuc_strike_idx[c];
        } else {
            // positions2idxs (beat/utility.py:1542-1558): round-half-even((pos - cell/2)/cell) -> int16
            const double pd = a.nuc_dip[(long)c * a.nuc_dip_stride + sf];
            const double ps = a.nuc_strike[(long)c * a.nuc_strike_stride + sf];
            const double half = __ddiv_rn(h, 2.0);
            const double xr = __ddiv_rn(__dsub_rn(pd, half), h), xc = __ddiv_rn(__dsub_rn(ps, half), h);
            hr = (xr == xr) ? (int)rint(xr) : -1;
            hc = (xc == xc) ? (int)rint(xc) : -1;
        }
        bad = (hr < 0) || (hr >= nd) || (hc < 0) || (hc >= nstr);
        if (bad) {   // the reference would write outside its array here; report instead
            if (gl == 0) {
                atomicAdd(a.violations, 1ULL);
                if (a.chain_bad) a.chain_bad[c] = 1;
            }
            hr = min(max(hr, 0), nd - 1);
            hc = min(max(hc, 0), nstr - 1);
        }
    }
    __syncwarp();

    const int iters = group_fast_sweep(T, Tprev, fh, c2, nd, nstr, hr, hc, gl, W, min(grp, cpw - 1) * W, valid);

    double tofs = 0.0;
    const bool add_time = (a.time != nullptr);
    if (add_time && valid) tofs = a.time[(long)c * a.time_stride + sf];
    double* out = a.t0 + (long)c * row_len + pofs;
    for (int k = gl; k < n; k += W) {
        double t = T[k];
        if (add_time) t = __dadd_rn(t, tofs);                               // seismic.py:1269
        if (bad) t = CUDART_NAN;
        out[k] = t;
    }
    if (a.n_iter && valid && gl == 0) a.n_iter[item] = iters;
}

}  // namespace beatgpu
