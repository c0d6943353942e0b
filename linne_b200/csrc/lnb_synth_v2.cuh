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
 * 32*TT taps and the history never leaves registers.  The block-channel is staged in shared memory
 * (40 KB per warp at the default block size), read from and written to HBM exactly once.  Units of one layer run side by side in lane groups (or one
 * unit per lane when there are 32 or more).  int32 arithmetic wraps exactly like the reference.
 */
#pragma once
#include "lnb_common.cuh"
#include "lnb_decode_core.cuh"

#define LNB_SY_WARPS   4
#define LNB_SY_THREADS (32 * LNB_SY_WARPS)
#define LNB_SY_MAX_N   10240               /* samples per block-channel staged in shared memory per warp */

/* One step of the single-lane form (one lane per unit, TT <= 4 taps, no shuffles). */
template <int TT, int JJ, bool GUARD>
__device__ __forceinline__ void lnb_sy_lane_step(int32_t *xu, uint32_t js, uint32_t m, uint32_t p, uint32_t rs, int32_t half,
                                                 bool active, const int32_t (&c)[TT], int32_t (&acc)[TT],
                                                 int32_t &y, int32_t &d1, int32_t &d2)
{
    if (GUARD && !(js + 1u < m)) return;
    const int32_t d3 = (active && (!GUARD || js + 3u < m)) ? xu[js + 3u] : 0;
#pragma unroll
    for (int r = 0; r < TT; r++)
        acc[r] = (int32_t)((uint32_t)acc[r] + (uint32_t)c[(r + JJ) % TT] * (uint32_t)y);
    constexpr int rc = (TT - 1 - JJ + TT) % TT;          /* register that just passed its last tap */
    const int32_t done = acc[rc];
    acc[rc] = half;
    const uint32_t jn = js + 1u;
    y = (jn >= p) ? (int32_t)((uint32_t)d1 - (uint32_t)(done >> rs)) : d1;
    if (active) xu[jn] = y;
    d1 = d2; d2 = d3;
}

template <int TT>
__device__ __forceinline__ void lnb_sy_round_lane(int32_t *x, uint32_t m, uint32_t p, uint32_t unit0,
                                                  uint32_t units_in_round, const int8_t *coef, uint32_t rs)
{
    const uint32_t lane = threadIdx.x & 31u;
    const bool active = lane < units_in_round;
    const uint32_t u = unit0 + (active ? lane : 0u);
    const int32_t half = rs ? (int32_t)(1u << (rs - 1u)) : 0;
    int32_t *xu = x + (size_t)u * m;
    int32_t c[TT], acc[TT];
#pragma unroll
    for (int s = 0; s < TT; s++) { c[s] = active ? (int32_t)coef[u * p + s] : 0; acc[s] = half; }
    int32_t y = active ? xu[0] : 0;
    int32_t d1 = (active && 1u < m) ? xu[1] : 0, d2 = (active && 2u < m) ? xu[2] : 0;
    uint32_t j = 0;
    for (; j + TT + 3u < m; j += TT) {                       /* whole groups, no bounds checks */
        lnb_sy_lane_step<TT, 0, false>(xu, j, m, p, rs, half, active, c, acc, y, d1, d2);
        if (TT > 1) lnb_sy_lane_step<TT, 1 % TT, false>(xu, j + 1u, m, p, rs, half, active, c, acc, y, d1, d2);
        if (TT > 2) lnb_sy_lane_step<TT, 2 % TT, false>(xu, j + 2u, m, p, rs, half, active, c, acc, y, d1, d2);
        if (TT > 2) lnb_sy_lane_step<TT, 3 % TT, false>(xu, j + 3u, m, p, rs, half, active, c, acc, y, d1, d2);
    }
    for (; j + 1u < m; j += TT) {                            /* ragged end */
        lnb_sy_lane_step<TT, 0, true>(xu, j, m, p, rs, half, active, c, acc, y, d1, d2);
        if (TT > 1) lnb_sy_lane_step<TT, 1 % TT, true>(xu, j + 1u, m, p, rs, half, active, c, acc, y, d1, d2);
        if (TT > 2) lnb_sy_lane_step<TT, 2 % TT, true>(xu, j + 2u, m, p, rs, half, active, c, acc, y, d1, d2);
        if (TT > 2) lnb_sy_lane_step<TT, 3 % TT, true>(xu, j + 3u, m, p, rs, half, active, c, acc, y, d1, d2);
    }
}

/* Systolic form for long filters: G >= 2 lanes per unit, 4 taps per lane, up to 32/G units side by side.
 * Software-pipelined so that neither shuffle sits on the sample-to-sample dependency chain:
 *   - the lane that owns the unit's stream (last lane of the group) finishes output j+1 from its own
 *     registers:  done = acc[last tap] + c_last * y_j   -- no broadcast needed for that;
 *   - y_j is broadcast right away; the other taps consume the broadcast value a little later;
 *   - a partial sum handed over from the previous lane (shuffle-up) is merged only three steps after
 *     it was sent, when its register reaches the third tap -- until then the register accumulates from
 *     zero (addition commutes), so the shuffle latency is hidden. */
struct LnbSyGroupState {
    int32_t c[4], acc[4], late[4];
    int32_t y, yb, d1, d2;
};

template <int JJ, bool GUARD>
__device__ __forceinline__ void lnb_sy_group_step(int32_t *xu, uint32_t js, uint32_t m, uint32_t p, uint32_t rs, int32_t half,
                                                  uint32_t G, bool io, bool is_first, LnbSyGroupState &st)
{
    constexpr int TT = 4;
    if (GUARD && !(js + 1u < m)) return;
    constexpr int rc = (TT - 1 - JJ + TT) % TT;          /* register at this lane's last tap in this step */
    const int32_t c_last = st.c[TT - 1];
    /* fast chain (stream-owning lane): next output from local state only */
    const int32_t d3 = (io && (!GUARD || js + 3u < m)) ? xu[js + 3u] : 0;
    const int32_t done = (int32_t)((uint32_t)st.acc[rc] + (uint32_t)c_last * (uint32_t)st.y);
    const uint32_t jn = js + 1u;
    const int32_t y_next = (jn >= p) ? (int32_t)((uint32_t)st.d1 - (uint32_t)(done >> rs)) : st.d1;
    if (io) xu[jn] = y_next;
    const int32_t yb_next = __shfl_sync(0xffffffffu, y_next, (int)(G - 1u), (int)G);
    /* slow part: every tap consumes the broadcast sample */
    const int32_t upd = (int32_t)((uint32_t)st.acc[rc] + (uint32_t)c_last * (uint32_t)st.yb);
#pragma unroll
    for (int r = 0; r < TT; r++) {
        if (r != rc) {
            constexpr int dummy = 0; (void)dummy;
            const int tap = (r + JJ) % TT;
            if (tap == TT - 2)                               /* merge the handed-over sum (first lane of a group: none) */
                st.acc[r] = (int32_t)((uint32_t)st.acc[r] + (uint32_t)(is_first ? 0 : st.late[r]));
            st.acc[r] = (int32_t)((uint32_t)st.acc[r] + (uint32_t)st.c[tap] * (uint32_t)st.yb);
        }
    }
    st.late[rc] = __shfl_up_sync(0xffffffffu, upd, 1, (int)G);   /* consumed three steps from now */
    st.acc[rc] = is_first ? half : 0;                             /* fresh partial sum enters at tap 0 next step */
    st.y = y_next; st.yb = yb_next;
    st.d1 = st.d2; st.d2 = d3;
}

__device__ __forceinline__ void lnb_sy_round_group(int32_t *x, uint32_t m, uint32_t p, uint32_t G, uint32_t unit0,
                                                   uint32_t units_in_round, const int8_t *coef, uint32_t rs)
{
    constexpr int TT = 4;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t gl = lane & (G - 1u), ug = lane / G;
    const bool active = ug < units_in_round;
    const uint32_t u = unit0 + (active ? ug : 0u);
    const bool is_first = gl == 0u;
    const bool io = active && (gl == G - 1u);
    const int32_t half = rs ? (int32_t)(1u << (rs - 1u)) : 0;
    int32_t *xu = x + (size_t)u * m;

    LnbSyGroupState st;
#pragma unroll
    for (int s = 0; s < TT; s++) {
        st.c[s] = active ? (int32_t)coef[u * p + gl * TT + s] : 0;
        st.acc[s] = half; st.late[s] = 0;
    }
    st.y = io ? xu[0] : 0;                                   /* y[0] = d[0] */
    st.d1 = (io && 1u < m) ? xu[1] : 0; st.d2 = (io && 2u < m) ? xu[2] : 0;
    st.yb = __shfl_sync(0xffffffffu, st.y, (int)(G - 1u), (int)G);

    uint32_t j = 0;
    for (; j + TT + 3u < m; j += TT) {                       /* whole groups of four steps, no bounds checks */
        lnb_sy_group_step<0, false>(xu, j, m, p, rs, half, G, io, is_first, st);
        lnb_sy_group_step<1, false>(xu, j + 1u, m, p, rs, half, G, io, is_first, st);
        lnb_sy_group_step<2, false>(xu, j + 2u, m, p, rs, half, G, io, is_first, st);
        lnb_sy_group_step<3, false>(xu, j + 3u, m, p, rs, half, G, io, is_first, st);
    }
    for (; j + 1u < m; j += TT) {                            /* ragged end */
        lnb_sy_group_step<0, true>(xu, j, m, p, rs, half, G, io, is_first, st);
        lnb_sy_group_step<1, true>(xu, j + 1u, m, p, rs, half, G, io, is_first, st);
        lnb_sy_group_step<2, true>(xu, j + 2u, m, p, rs, half, G, io, is_first, st);
        lnb_sy_group_step<3, true>(xu, j + 3u, m, p, rs, half, G, io, is_first, st);
    }
}

/* one layer of one block-channel: pick the lane mapping from (P, U) */
__device__ __forceinline__ void lnb_sy_layer(int32_t *x, uint32_t n, uint32_t P, uint32_t U, const int8_t *coef, uint32_t rs)
{
    const uint32_t p = P / U, m = n / U;
    if (m <= p) return;                                       /* nothing is predicted (reference would underflow) */
    uint32_t G, TT;
    if (U >= 32u || p <= 4u) { G = 1u; TT = p; }              /* short filters: one lane per unit, no shuffles */
    else { G = 32u / U; if (G > p / 4u) G = p / 4u; TT = p / G; }
    const uint32_t per_round = 32u / G;
    for (uint32_t u0 = 0; u0 < U; u0 += per_round) {
        const uint32_t cnt = (U - u0 < per_round) ? U - u0 : per_round;
        if (G > 1u) lnb_sy_round_group(x, m, p, G, u0, cnt, coef, rs);
        else if (TT == 1u) lnb_sy_round_lane<1>(x, m, p, u0, cnt, coef, rs);
        else if (TT == 2u) lnb_sy_round_lane<2>(x, m, p, u0, cnt, coef, rs);
        else lnb_sy_round_lane<4>(x, m, p, u0, cnt, coef, rs);
        __syncwarp();
    }
}

/* n_max = samples of shared memory per warp; block-channels longer than that are left to the flat kernels */
__global__ void __launch_bounds__(LNB_SY_THREADS) lnb_synth_v2_kernel(LnbDecodeBatch b, uint32_t n_max)
{
    extern __shared__ __align__(16) int32_t lnb_sy_smem[];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const uint32_t bc = blockIdx.x * LNB_SY_WARPS + warp;
    if (bc >= b.num_blocks * b.cfg.num_channels) return;
    const uint32_t blk_i = bc / b.cfg.num_channels, ch = bc % b.cfg.num_channels;
    const LnbBlockDesc blk = b.blocks[blk_i];
    if (blk.type != LNB_BLOCK_COMPRESSED || blk.status) return;
    const uint32_t n = blk.nsmp;
    if (n > n_max) return;
    if (b.fused_max_n && n <= b.fused_max_n) return;         /* done inside the fused streaming kernel */
    const LnbChanParams &prm = b.params[bc];
    int32_t *gx = b.pcm + (size_t)ch * b.cfg.pcm_stride + blk.smp_off;
    int32_t *x = lnb_sy_smem + (size_t)warp * n_max;

    for (uint32_t i = lane; i < n; i += 32u) x[i] = gx[i];
    __syncwarp();

    for (int l = (int)b.cfg.num_layers - 1; l >= 0; l--) {
        const uint32_t P = b.cfg.layer_params[l];
        uint32_t U = 1u << prm.log2_units[l];
        if (U > P) continue;                                  /* not a stream an encoder writes */
        lnb_sy_layer(x, n, P, U, prm.coef + l * LNB_MAX_PARAMS, prm.rshift[l]);
        __syncwarp();
    }

    /* de-emphasis: z[i] = x[i] + ((z[i-1]*c1) >> 5), y[i] = z[i] + ((y[i-1]*c0) >> 5): serial by nature */
    if (lane == 0) {
        const int32_t c0 = prm.preem_coef[0], c1 = prm.preem_coef[1];
        int32_t zp = prm.preem_prev[1], yp = prm.preem_prev[0];
#pragma unroll 8
        for (uint32_t i = 0; i < n; i++) {
            const int32_t z = x[i] + ((zp * c1) >> LNB_PREEM_SHIFT);
            const int32_t yv = z + ((yp * c0) >> LNB_PREEM_SHIFT);
            x[i] = yv; zp = z; yp = yv;
        }
    }
    __syncwarp();
    for (uint32_t i = lane; i < n; i += 32u) gx[i] = x[i];
}
