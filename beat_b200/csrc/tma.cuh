// tma.cuh -- thin inline-PTX wrappers (sm_100a) around the 1-D bulk asynchronous copy (TMA engine, SASS UBLKCP)
//            and the shared-memory mbarrier it signals.  Used by the delay-and-sum kernel (geom.cuh), which stages
//            windows of GF-store traces in shared memory, and by the gather probe (probe.cuh).
//
// Rules the callers follow: source address and byte count are multiples of 16, the destination is 16-byte aligned
// shared memory; one thread arms the barrier with the expected byte count before (or together with) issuing the
// copies; consumers spin on try_wait with a bound (a kernel must never hang a box: on time-out they raise a flag
// in global memory and fall through).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace beatgpu {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

// make barrier initialisation visible to the async proxy (the TMA engine)
__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

// generic-proxy writes/reads of shared memory ordered before a later async-proxy (TMA) write of the same bytes
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}

// Bounded wait: returns false after ~max_polls unsuccessful polls (each poll itself blocks for a
// hardware-defined interval), so a lost copy becomes an error flag instead of a hung GPU.
__device__ __forceinline__ bool mbar_wait_bounded(uint64_t* bar, uint32_t parity, uint32_t max_polls = (1u << 22),
                                                  uint32_t backoff_ns = 0)
{
    for (uint32_t i = 0; i < max_polls; ++i) {
        if (mbar_try_wait(bar, parity)) return true;
        if (backoff_ns) __nanosleep(backoff_ns);
    }
    return false;
}

// global -> shared bulk copy of `bytes` (multiple of 16), completion counted on `bar`
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

}  // namespace beatgpu
