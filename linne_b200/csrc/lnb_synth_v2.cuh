/* lnb_synth_v2.cuh -- cooperative decoder back end: one warp per (block, channel).
 *
 * Replaces the flat synthesis (D2) and de-emphasis (D3) kernels.  Covers reference rows d4, d5
 * (SURVEY section 8a): libs/linne_decoder/src/linne_lpc_synthesize.c:8-83 (order-P integer
 * recursion per unit, layers L-1 .. 0, linne_decoder.c:503-509) and
 * libs/linne_internal/src/linne_utility.c:215-241 (two cascaded first-order de-emphasis filters).
 *
 * The synthesis recursion  y[i] = d[i] - ((half + sum_k c[k] y[i-p+k]) >> s)  is run in TRANSPOSED
 * (systolic) form: instead of re-reading the last p outputs for every sample, each new output is
 * broadcast once and every tap adds its product to the partial sum of the future output it belongs
 * to.  A unit's p taps are spread over G lanes x TT taps per lane; partial sums march from tap to
 * tap (register renaming inside a lane, one shuffle between lanes), and the sum leaving the last tap
 * completes the next output.  Per sample a warp issues TT multiply-adds + 2 shuffles for up to
 * 32*TT taps, the history never leaves registers, and the samples are read and written exactly once
 * per layer with coalesced accesses.  Units of one layer run side by side in lane groups (or one
 * unit per lane when there are 32 or more).  int32 arithmetic wraps exactly like the reference.
 */
#pragma once
#include "lnb_common.cuh"
#include "lnb_decode_core.cuh"

#define LNB_SY_WARPS   4
#define LNB_SY_THREADS (32 * LNB_SY_WARPS)
#define LNB_SY_TILE    256                 /* de-emphasis staging tile per warp (samples) */

/* One round of one layer: up to 32/G units side by side, G lanes per unit, TT taps per lane. */
template <int TT>
__device__ __forceinline__ void lnb_sy_round(int32_t *x, uint32_t m, uint32_t p, uint32_t G, uint32_t unit0,
                                             uint32_t units_in_round, const int8_t *coef, uint32_t rs)
{
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t gl = lane & (G - 1u), ug = lane / G;
    const bool active = ug < units_in_round;
    const uint32_t u = unit0 + (active ? ug : 0u);
    const bool is_last = gl == G - 1u, is_first = gl == 0u;
    const int32_t half = rs ? (int32_t)(1u << (rs - 1u)) : 0;
    int32_t *xu = x + (size_t)u * m;

    int32_t c[TT], acc[TT];
#pragma unroll
    for (int s = 0; s < TT; s++) {
        c[s] = active ? (int32_t)coef[u * p + gl * TT + s] : 0;
        acc[s] = half;
    }
    int32_t y = (active && is_last) ? xu[0] : 0;          /* y[0] = d[0] */
    int32_t keep = 0;

    for (uint32_t j = 0; j + 1u < m; j += TT) {
#pragma unroll
        for (int jj = 0; jj < TT; jj++) {
            const uint32_t js = j + (uint32_t)jj;           /* sample being consumed */
            if (js + 1u < m) {
                const int32_t yb = (G > 1u) ? __shfl_sync(0xffffffffu, y, (int)(G - 1u), (int)G) : y;
                if (G > 1u) {
                    if (gl == (js & (G - 1u))) keep = yb;
                    if ((js & (G - 1u)) == G - 1u && active) xu[js - (G - 1u) + gl] = keep;
                }
                /* physical register r sits at tap (r + jj) % TT of this lane in this step */
#pragma unroll
                for (int r = 0; r < TT; r++)
                    acc[r] = (int32_t)((uint32_t)acc[r] + (uint32_t)c[(r + jj) % TT] * (uint32_t)yb);
                const int rc = (TT - 1 - jj + TT) % TT;     /* register that just passed its last tap */
                const int32_t done = acc[rc];
                const int32_t incoming = (G > 1u) ? __shfl_up_sync(0xffffffffu, done, 1, (int)G) : done;
                acc[rc] = (is_first || G == 1u) ? half : incoming;
                if (G == 1u) acc[rc] = half;
                const uint32_t jn = js + 1u;
                const int32_t dn = (active && is_last) ? xu[jn] : 0;
                y = (jn >= p) ? (int32_t)((uint32_t)dn - (uint32_t)(done >> rs)) : dn;
                if (G == 1u && active) xu[jn] = y;
            }
        }
    }
    if (G > 1u) {                                           /* flush the last, possibly partial, group of outputs */
        const uint32_t js = m - 1u;
        const int32_t yb = __shfl_sync(0xffffffffu, y, (int)(G - 1u), (int)G);
        if (gl == (js & (G - 1u))) keep = yb;
        const uint32_t idx = (js & ~(G - 1u)) + gl;
        if (active && idx <= js) xu[idx] = keep;
    }
}

/* one layer of one block-channel: pick the lane mapping from (P, U) */
__device__ __forceinline__ void lnb_sy_layer(int32_t *x, uint32_t n, uint32_t P, uint32_t U, const int8_t *coef, uint32_t rs)
{
    const uint32_t p = P / U, m = n / U;
    if (m <= p) return;                                       /* nothing is predicted (reference would underflow) */
    uint32_t G, TT;
    if (U >= 32u) { G = 1u; TT = p; }
    else { G = 32u / U; if (G > p) G = p; TT = p / G; }
    const uint32_t per_round = 32u / G;
    for (uint32_t u0 = 0; u0 < U; u0 += per_round) {
        const uint32_t cnt = (U - u0 < per_round) ? U - u0 : per_round;
        if (TT == 1u) lnb_sy_round<1>(x, m, p, G, u0, cnt, coef, rs);
        else if (TT == 2u) lnb_sy_round<2>(x, m, p, G, u0, cnt, coef, rs);
        else lnb_sy_round<4>(x, m, p, G, u0, cnt, coef, rs);
        __syncwarp();
    }
}

__global__ void __launch_bounds__(LNB_SY_THREADS) lnb_synth_v2_kernel(LnbDecodeBatch b)
{
    __shared__ int32_t tile[LNB_SY_WARPS][LNB_SY_TILE];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const uint32_t bc = blockIdx.x * LNB_SY_WARPS + warp;
    if (bc >= b.num_blocks * b.cfg.num_channels) return;
    const uint32_t blk_i = bc / b.cfg.num_channels, ch = bc % b.cfg.num_channels;
    const LnbBlockDesc blk = b.blocks[blk_i];
    if (blk.type != LNB_BLOCK_COMPRESSED || blk.status) return;
    const LnbChanParams &prm = b.params[bc];
    const uint32_t n = blk.nsmp;
    int32_t *x = b.pcm + (size_t)ch * b.cfg.pcm_stride + blk.smp_off;

    for (int l = (int)b.cfg.num_layers - 1; l >= 0; l--) {
        const uint32_t P = b.cfg.layer_params[l];
        uint32_t U = 1u << prm.log2_units[l];
        if (U > P) continue;                                  /* not a stream an encoder writes */
        lnb_sy_layer(x, n, P, U, prm.coef + l * LNB_MAX_PARAMS, prm.rshift[l]);
        __syncwarp();
    }

    /* de-emphasis: z[i] = x[i] + ((z[i-1]*c1) >> 5), y[i] = z[i] + ((y[i-1]*c0) >> 5); serial by nature,
     * so one lane runs the two recurrences over tiles the whole warp stages through shared memory */
    {
        const int32_t c0 = prm.preem_coef[0], c1 = prm.preem_coef[1];
        int32_t zp = prm.preem_prev[1], yp = prm.preem_prev[0];
        int32_t *t = tile[warp];
        for (uint32_t base = 0; base < n; base += LNB_SY_TILE) {
            const uint32_t cnt = (n - base < LNB_SY_TILE) ? n - base : LNB_SY_TILE;
            for (uint32_t i = lane; i < cnt; i += 32u) t[i] = x[base + i];
            __syncwarp();
            if (lane == 0) {
#pragma unroll 8
                for (uint32_t i = 0; i < cnt; i++) {
                    const int32_t z = t[i] + ((zp * c1) >> LNB_PREEM_SHIFT);
                    const int32_t yv = z + ((yp * c0) >> LNB_PREEM_SHIFT);
                    t[i] = yv; zp = z; yp = yv;
                }
            }
            __syncwarp();
            for (uint32_t i = lane; i < cnt; i += 32u) x[base + i] = t[i];
            __syncwarp();
        }
    }
}
