/* lnb_bulk.cuh -- bulk-asynchronous global -> shared copies (the 1-D form of the tensor memory accelerator:
 * cp.async.bulk, SASS UBLKCP) completing on an mbarrier, for sm_100a.
 *
 * One thread issues a copy of up to a few KB with a single instruction; the bytes land in shared memory without
 * passing through registers and the consumer learns of it through the barrier's transaction count.  This is what
 * feeds the bitstream windows of the decoders (lnb_stream_v2.cuh): the serial code-word walk of a block runs on one
 * lane, and that lane keeps its own window filled a few chunks ahead at the cost of one instruction per chunk.
 *
 * Source and destination must be 16-byte aligned, the size a multiple of 16.
 */
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)

__device__ __forceinline__ uint32_t lnb_smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void lnb_mbar_init(uint64_t *bar, uint32_t arrivals)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(lnb_smem_addr(bar)), "r"(arrivals) : "memory");
}
/* make freshly initialised barriers visible to the asynchronous proxy (call once, before the CTA barrier) */
__device__ __forceinline__ void lnb_mbar_init_fence()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
/* this thread's arrival plus the announcement of `bytes` of asynchronous traffic for the current phase */
__device__ __forceinline__ void lnb_mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(lnb_smem_addr(bar)), "r"(bytes) : "memory");
}
/* true once the phase of the given parity has completed */
__device__ __forceinline__ bool lnb_mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(lnb_smem_addr(bar)), "r"(parity) : "memory");
    return ok != 0u;
}
__device__ __forceinline__ void lnb_mbar_wait(uint64_t *bar, uint32_t parity)
{
    while (!lnb_mbar_try_wait(bar, parity)) { }
}
/* global -> shared, `bytes` (multiple of 16) announced on `bar` when they have landed */
__device__ __forceinline__ void lnb_bulk_load(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(lnb_smem_addr(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(lnb_smem_addr(bar)) : "memory");
}

/* order this thread's earlier generic-proxy accesses to shared memory before asynchronous-proxy writes it issues next */
__device__ __forceinline__ void lnb_proxy_fence_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

/* ---- shared-memory accesses by 32-bit window address (no generic-address arithmetic on a hot path) ---- */
__device__ __forceinline__ void lnb_sts32(uint32_t addr, uint32_t v)
{
    asm volatile("st.shared.u32 [%0], %1;" :: "r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t lnb_lds32(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
/* index of the most significant one bit, 0xFFFFFFFF for zero (31 - clz without the subtraction) */
__device__ __forceinline__ uint32_t lnb_bfind(uint32_t x)
{
    uint32_t r;
    asm("bfind.u32 %0, %1;" : "=r"(r) : "r"(x));
    return r;
}

#endif /* __CUDACC__ */
