/* lnb_scan_v2.cuh -- block byte offsets: exclusive scan of the block sizes of one batch, one CTA.
 *
 * Replaces the single-thread scan (stage E8b).  The north star's "per-frame bit offsets from a
 * device-wide exclusive scan": each thread sums a contiguous run of blocks, runs are scanned across the
 * CTA (warp shuffles + one shared-memory pass), then every thread writes its blocks' offsets.  A batch is
 * at most a few thousand blocks (the analysis scratch bounds it), so one CTA of 1024 threads is enough;
 * the scan of SHARD totals across GPUs is done by the host (linne_b200/shard.py).
 */
#pragma once
#include "lnb_common.cuh"

#define LNB_SC_THREADS 1024

__global__ void __launch_bounds__(LNB_SC_THREADS) lnb_scan_v2_kernel(LnbEncodeBatch b)
{
    __shared__ uint32_t warp_sums[LNB_SC_THREADS / 32];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t per = (b.num_blocks + LNB_SC_THREADS - 1u) / LNB_SC_THREADS;
    const uint32_t lo = tid * per, hi = (lo + per < b.num_blocks) ? lo + per : b.num_blocks;
    uint32_t sum = 0;
    for (uint32_t i = lo; i < hi; i++) sum += b.blocks[i].byte_size;
    uint32_t inc = sum;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, off);
        if (lane >= (uint32_t)off) inc += t;
    }
    if (lane == 31u) warp_sums[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = warp_sums[lane];
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, w, off);
            if (lane >= (uint32_t)off) w += t;
        }
        warp_sums[lane] = w;                               /* inclusive scan of the warp totals */
    }
    __syncthreads();
    uint32_t off = b.out_base + (warp ? warp_sums[warp - 1u] : 0u) + inc - sum;
    for (uint32_t i = lo; i < hi; i++) { b.blocks[i].byte_off = off; off += b.blocks[i].byte_size; }
    if (tid == LNB_SC_THREADS - 1u) *b.total_size = warp_sums[LNB_SC_THREADS / 32 - 1u];
}
