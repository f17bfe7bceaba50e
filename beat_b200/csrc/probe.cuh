// probe.cuh -- measurement kernels (diagnostics, not on the product path): how fast can this GPU deliver
//              data-dependent "rows" (a few hundred contiguous bytes each, picked from a working set) to the SMs?
//
// gf_stack_chunk_kernel (stack.cuh) gathers one such row per (chain, target, patch, tap) and ncu shows it sitting on
// the L2->SM return path.  These probes measure that ceiling directly, so bench.py can report the stack kernel
// against a MEASURED gather peak instead of an assumed 64 B/clk/SM:
//   mode 0  one warp-wide LDG.128 per row, 8 rows in flight per warp           (the stack kernel's access pattern)
//   mode 1  one cp.async.bulk (TMA) per row into a per-warp shared-memory ring, lanes then read the row from
//           shared memory                                                      (what a TMA-staged gather would do)
//   mode 2  as mode 1 without the shared-memory read                           (pure TMA ingest ceiling)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "tma.cuh"

namespace beatgpu {

constexpr int kProbeThreads = 128;
constexpr int kProbeWarps = kProbeThreads / 32;
constexpr int kProbeDepth = 8;            // rows (LDG: 512-byte row pieces) in flight per warp
constexpr int kProbeMaxRow = 16384;       // bytes

struct ProbeArgs {
    const unsigned char* ws;              // working set, n_rows rows of row_stride bytes
    long n_rows;                          // power of two
    uint32_t row_mask;                    // n_rows - 1
    int row_bytes, row_stride;
    int rows_per_warp;
    int depth;                            // TMA modes: ring slots per warp
    float* sink;
    unsigned int* err;
};

__device__ __forceinline__ uint32_t probe_hash(uint32_t x)
{
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}

// rows <= 512 B: one LDG.128 per lane per row, kProbeDepth rows in flight; longer rows: kProbeDepth 512-byte pieces
// of the same row in flight.
__global__ void __launch_bounds__(kProbeThreads) probe_ldg_kernel(ProbeArgs a)
{
    const int lane = threadIdx.x & 31;
    const uint32_t gw = blockIdx.x * kProbeWarps + (threadIdx.x >> 5);
    float acc = 0.f;
    if (a.row_bytes <= 512) {
        const bool active = lane * 16 < a.row_bytes;
        for (int it = 0; it < a.rows_per_warp; it += kProbeDepth) {
            float4 v[kProbeDepth];
#pragma unroll
            for (int u = 0; u < kProbeDepth; ++u) {
                const long r = probe_hash(gw * 7919u + (uint32_t)(it + u)) & a.row_mask;
                v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (active) v[u] = __ldg((const float4*)(a.ws + r * a.row_stride) + lane);
            }
#pragma unroll
            for (int u = 0; u < kProbeDepth; ++u) acc += v[u].x + v[u].y + v[u].z + v[u].w;
        }
    } else if (a.row_bytes <= 1024) {
        // f64-library rows (960 B): two 16-byte loads per lane per row, kProbeDepth rows in flight (the f64 stack kernel's pattern)
        const bool act0 = lane * 16 < a.row_bytes, act1 = (lane + 32) * 16 < a.row_bytes;
        for (int it = 0; it < a.rows_per_warp; it += kProbeDepth) {
            float4 v0[kProbeDepth], v1[kProbeDepth];
#pragma unroll
            for (int u = 0; u < kProbeDepth; ++u) {
                const long r = probe_hash(gw * 7919u + (uint32_t)(it + u)) & a.row_mask;
                const float4* row = (const float4*)(a.ws + r * a.row_stride);
                v0[u] = v1[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (act0) v0[u] = __ldg(row + lane);
                if (act1) v1[u] = __ldg(row + lane + 32);
            }
#pragma unroll
            for (int u = 0; u < kProbeDepth; ++u) acc += v0[u].x + v0[u].y + v0[u].z + v0[u].w + v1[u].x + v1[u].y + v1[u].z + v1[u].w;
        }
    } else {
        const int nvec = a.row_bytes / 16;
        for (int it = 0; it < a.rows_per_warp; ++it) {
            const long r = probe_hash(gw * 7919u + (uint32_t)it) & a.row_mask;
            const float4* row = (const float4*)(a.ws + r * a.row_stride);
            for (int j0 = 0; j0 < nvec; j0 += 32 * kProbeDepth) {
                float4 v[kProbeDepth];
#pragma unroll
                for (int u = 0; u < kProbeDepth; ++u) {
                    const int j = j0 + u * 32 + lane;
                    v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (j < nvec) v[u] = __ldg(row + j);
                }
#pragma unroll
                for (int u = 0; u < kProbeDepth; ++u) acc += v[u].x + v[u].y + v[u].z + v[u].w;
            }
        }
    }
    if (acc == 123.456f) a.sink[0] = acc;      // never true for the zero-filled working set; defeats dead-code elimination
}

template <bool READ_SMEM>
__global__ void __launch_bounds__(kProbeThreads) probe_tma_kernel(ProbeArgs a)
{
    extern __shared__ __align__(128) unsigned char probe_smem[];       // [warps][depth][row_bytes] then barriers
    const int lane = threadIdx.x & 31;
    const int w = threadIdx.x >> 5;
    const int depth = a.depth;
    unsigned char* ring = probe_smem + (size_t)w * depth * a.row_bytes;
    uint64_t* bar = (uint64_t*)(probe_smem + (size_t)kProbeWarps * depth * a.row_bytes) + w * depth;
    const uint32_t gw = blockIdx.x * kProbeWarps + w;
    if (lane == 0) {
        for (int s = 0; s < depth; ++s) mbar_init(&bar[s], 1);
        mbar_fence_init();
    }
    __syncwarp();
    const int n = a.rows_per_warp;
    const int nvec = a.row_bytes / 16;
    if (lane == 0) {
        for (int s = 0; s < depth && s < n; ++s) {
            const long r = probe_hash(gw * 7919u + (uint32_t)s) & a.row_mask;
            mbar_arrive_expect_tx(&bar[s], (uint32_t)a.row_bytes);
            tma_load_1d(ring + (size_t)s * a.row_bytes, a.ws + r * a.row_stride, (uint32_t)a.row_bytes, &bar[s]);
        }
    }
    float acc = 0.f;
    for (int it = 0; it < n; ++it) {
        const int s = it % depth;
        const uint32_t parity = (uint32_t)(it / depth) & 1u;
        if (!mbar_wait_bounded(&bar[s], parity)) { if (lane == 0) atomicAdd(a.err, 1u); break; }
        if (READ_SMEM) {
            const float4* row = (const float4*)(ring + (size_t)s * a.row_bytes);
            for (int j = lane; j < nvec; j += 32) {
                const float4 v = row[j];
                acc += v.x + v.y + v.z + v.w;
            }
        }
        __syncwarp();                         // all lanes are done with slot s
        if (lane == 0 && it + depth < n) {
            const long r = probe_hash(gw * 7919u + (uint32_t)(it + depth)) & a.row_mask;
            fence_proxy_async();
            mbar_arrive_expect_tx(&bar[s], (uint32_t)a.row_bytes);
            tma_load_1d(ring + (size_t)s * a.row_bytes, a.ws + r * a.row_stride, (uint32_t)a.row_bytes, &bar[s]);
        }
    }
    if (acc == 123.456f) a.sink[0] = acc;
}

// mode 3 / 4 -- BATCHED bulk copies: one mbarrier per batch of kBatchRows rows (expect_tx = kBatchRows * row_bytes, armed once
// by lane 0), lanes 0..kBatchRows-1 each issue one cp.async.bulk, kBatchRing batches in flight per warp.  This is the fair
// TMA probe: the per-copy issue and the barrier round trip are amortised over 16 rows instead of paid per row (mode 1 / 2).
// mode 3 reads the rows back from shared memory (what a TMA-staged stacking kernel would do), mode 4 only ingests.
constexpr int kBatchRows = 16;
constexpr int kBatchRing = 3;

template <bool READ_SMEM>
__global__ void __launch_bounds__(kProbeThreads) probe_tma_batch_kernel(ProbeArgs a)
{
    extern __shared__ __align__(128) unsigned char probe_smem[];       // [warps][ring][kBatchRows][row_bytes] then barriers
    const int lane = threadIdx.x & 31;
    const int w = threadIdx.x >> 5;
    const size_t batch_bytes = (size_t)kBatchRows * a.row_bytes;
    unsigned char* ring = probe_smem + (size_t)w * kBatchRing * batch_bytes;
    uint64_t* bar = (uint64_t*)(probe_smem + (size_t)kProbeWarps * kBatchRing * batch_bytes) + w * kBatchRing;
    const uint32_t gw = blockIdx.x * kProbeWarps + w;
    if (lane == 0) {
        for (int s = 0; s < kBatchRing; ++s) mbar_init(&bar[s], 1);
        mbar_fence_init();
    }
    __syncwarp();
    const int n_batches = a.rows_per_warp / kBatchRows;
    const int nvec = a.row_bytes / 16;
    auto issue = [&](int b) {
        const int s = b % kBatchRing;
        if (lane == 0) mbar_arrive_expect_tx(&bar[s], (uint32_t)batch_bytes);
        __syncwarp();
        if (lane < kBatchRows) {
            const long r = probe_hash(gw * 7919u + (uint32_t)(b * kBatchRows + lane)) & a.row_mask;
            tma_load_1d(ring + (size_t)s * batch_bytes + (size_t)lane * a.row_bytes, a.ws + r * a.row_stride, (uint32_t)a.row_bytes, &bar[s]);
        }
    };
    for (int b = 0; b < kBatchRing && b < n_batches; ++b) issue(b);
    float acc = 0.f;
    for (int b = 0; b < n_batches; ++b) {
        const int s = b % kBatchRing;
        const uint32_t parity = (uint32_t)(b / kBatchRing) & 1u;
        if (!mbar_wait_bounded(&bar[s], parity)) { if (lane == 0) atomicAdd(a.err, 1u); break; }
        if (READ_SMEM) {
            const unsigned char* base = ring + (size_t)s * batch_bytes;
#pragma unroll 4
            for (int r = 0; r < kBatchRows; ++r) {
                const float4* row = (const float4*)(base + (size_t)r * a.row_bytes);
                for (int j = lane; j < nvec; j += 32) {
                    const float4 v = row[j];
                    acc += v.x + v.y + v.z + v.w;
                }
            }
        }
        __syncwarp();
        if (b + kBatchRing < n_batches) {
            if (lane == 0) fence_proxy_async();
            issue(b + kBatchRing);
        }
    }
    if (acc == 123.456f) a.sink[0] = acc;
}

// mode 5 / 6 -- rows served from (distributed) shared memory: every CTA of a cluster holds `rows_smem` rows in its shared
// memory; warps gather pseudo-random rows from pseudo-random CTAs of the cluster with one 16-byte ld.shared::cluster per
// lane, kProbeDepth rows in flight.  Cluster size 1 = plain local shared memory (mode 6): the ceiling a row-staged
// stacking kernel could reach if the rows it needs were already on the SM.
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t cta_rank)
{
    uint32_t out;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(out) : "r"(addr), "r"(cta_rank));
    return out;
}

__device__ __forceinline__ float4 ld_shared_cluster_v4(uint32_t addr)
{
    float4 v;
    asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}

__global__ void __launch_bounds__(kProbeThreads) probe_dsmem_kernel(ProbeArgs a, int rows_smem, int cluster_size)
{
    extern __shared__ __align__(128) unsigned char probe_smem[];       // [rows_smem][row_bytes]
    const int lane = threadIdx.x & 31;
    const uint32_t gw = blockIdx.x * kProbeWarps + (threadIdx.x >> 5);
    for (int i = threadIdx.x; i < rows_smem * a.row_bytes / 16; i += kProbeThreads) ((float4*)probe_smem)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (cluster_size > 1) {
        asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    } else {
        __syncthreads();
    }
    const uint32_t base = smem_u32(probe_smem);
    const bool active = lane * 16 < a.row_bytes;
    float acc = 0.f;
    for (int it = 0; it < a.rows_per_warp; it += kProbeDepth) {
        float4 v[kProbeDepth];
#pragma unroll
        for (int u = 0; u < kProbeDepth; ++u) {
            // cheap row pick (the probe must not be issue-bound): one multiply-add, masks instead of modulos
            // (rows_smem and cluster_size are powers of two)
            const uint32_t h = (gw * 7919u + (uint32_t)(it + u)) * 2654435761u;
            const uint32_t r = (h >> 12) & (uint32_t)(rows_smem - 1);
            uint32_t addr = base + r * (uint32_t)a.row_bytes + (uint32_t)lane * 16u;
            if (cluster_size > 1) addr = mapa_shared(addr, (h >> 4) & (uint32_t)(cluster_size - 1));
            v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (active) v[u] = ld_shared_cluster_v4(addr);
        }
#pragma unroll
        for (int u = 0; u < kProbeDepth; ++u) acc += v[u].x + v[u].y + v[u].z + v[u].w;
    }
    if (acc == 123.456f) a.sink[0] = acc;
    if (cluster_size > 1)       // no CTA may exit while a peer still reads its shared memory
        asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

}  // namespace beatgpu
