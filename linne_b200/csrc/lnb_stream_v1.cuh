/* lnb_stream_v1.cuh -- fused streaming decoder: one CTA per block, the entropy decoder feeding the
 * synthesis cascade sample by sample.
 *
 * Covers reference rows d2-d5 in ONE kernel: entropy decode
 * (libs/linne_decoder/src/linne_decoder.c:457-497, libs/linne_coder/src/linne_coder.c:306-327), synthesis of
 * every layer (linne_lpc_synthesize.c:8-83 in the order of linne_decoder.c:503-509), the two de-emphasis
 * filters (linne_utility.c:215-241) and mid/side -> left/right (:135-147).
 *
 * Why.  With separate kernels a block-channel waits for its whole entropy decode, then runs its layers one
 * after the other: latency = entropy + sum of layers.  Every one of these stages is a sequential recursion
 * over the samples, and each only needs the PREVIOUS stage's output at the same position -- so they form a
 * pipeline.  Here each stage is a warp:
 *     warp 0          entropy decoder (lnb_entropy_v3.cuh), writes residuals into a shared-memory line
 *     warp 1..L       synthesis layer L-1 .. 0, in place on that line, trailing the previous stage
 *     warp L+1        de-emphasis (+ M/S inverse on the second channel), then coalesced stores to the PCM plane
 * and the latency of a block becomes entropy + a pipeline lag of ~100 samples.  A layer split into U units is
 * walked unit after unit (its frontier has to follow the stage in front of it), with the systolic 4-taps-per-lane
 * form of lnb_synth_v2.cuh for units of 8 or more taps and one lane for shorter ones.
 * Stages hand over through progress counters in shared memory (sample counts over the block's channels,
 * published every 32 steps, polled with a short sleep).  A channel's line is reused by the next channel once
 * the last stage has drained it.
 *
 * Blocks longer than LNB_DS_MAX_N samples, and blocks whose CRC check failed, keep the split kernels.
 */
#pragma once
#include "lnb_common.cuh"
#include "lnb_decode_core.cuh"
#include "lnb_entropy_v3.cuh"
#include "lnb_synth_v2.cuh"
#include "lnb_tput_v1.cuh"

#define LNB_DS_MAX_N    10240u
#define LNB_DS_WARPS    (2u + LNB_MAX_LAYERS)          /* entropy + layers + de-emphasis */
#define LNB_DS_THREADS  (32u * (LNB_DS_WARPS + 1u))    /* warp 4 stays empty: it would share a scheduler with the entropy warp */
#define LNB_DS_BATCH    32u                            /* steps between progress updates */
#define LNB_DS_ABORT    0xFFFFFFFFu

struct LnbDsShared {
    volatile uint32_t prog[LNB_DS_WARPS];              /* samples (over all channels of the block) each stage has finished */
    volatile uint32_t abort;
    uint32_t win[LNB_E3_WIN + 2];
    uint16_t huff1[1u << LNB_E3_HUFF1_BITS];
    LnbChanParams params[LNB_MAX_CHANNELS];
};

__device__ __forceinline__ void lnb_ds_publish(LnbDsShared &sm, uint32_t stage, uint32_t value, uint32_t lane)
{
    __syncwarp();
    if (lane == 0) { __threadfence_block(); sm.prog[stage] = value; }
}
/* wait until stage `stage` has finished at least `need` samples; false when the block was aborted */
__device__ __forceinline__ bool lnb_ds_wait(LnbDsShared &sm, uint32_t stage, uint32_t need)
{
    for (;;) {
        const uint32_t have = sm.prog[stage];
        if (have >= need) break;
        if (sm.abort) return false;
        /* the pace maker (entropy warp) needs ~50 ns per sample: sleep roughly until the missing samples can exist,
         * so waiting stages leave the issue slots to the warps that have work */
        const uint32_t ns = (need - have) * 32u;
        __nanosleep(ns < 64u ? 64u : (ns > 4000u ? 4000u : ns));
    }
    __threadfence_block();
    return sm.abort == 0u;
}

/* residual sink of the entropy warp: the shared line of the current channel */
struct LnbDsSink {
    static constexpr bool kPublish = true;
    LnbChanParams *params;
    LnbDsShared *sm;
    int32_t *line;
    uint32_t last_stage;
    __device__ __forceinline__ int32_t *channel(uint32_t) const { return line; }
    __device__ __forceinline__ void begin_channel(uint32_t c, uint32_t n, uint32_t) const
    {   /* the line is free once the last stage has drained the previous channel */
        lnb_ds_wait(*sm, last_stage, c * n);
    }
    __device__ __forceinline__ void publish(uint32_t c, uint32_t n, uint32_t done, uint32_t lane) const
    {
        lnb_ds_publish(*sm, 0u, c * n + done, lane);
    }
    __device__ __forceinline__ void abort(uint32_t lane) const
    {
        __syncwarp();
        if (lane == 0) { sm->abort = 1u; __threadfence_block(); }
    }
};

/* ---- one unit [xu, xu + m) of a layer, p >= 8 taps: G = p/4 lanes, systolic, trailing `up` ---- */
__device__ bool lnb_ds_unit_group(LnbDsShared &sm, uint32_t up, uint32_t self, uint32_t g0 /* global index of xu[0] */, uint32_t g_end,
                                  int32_t *xu, uint32_t m, uint32_t p, const int8_t *coef, uint32_t rs)
{
    constexpr int TT = 4;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t G = p / 4u;
    const bool active = lane < G;
    const uint32_t gl = active ? lane : 0u;
    const bool is_first = gl == 0u;
    const bool io = active && (gl == G - 1u);
    const int32_t half = rs ? (int32_t)(1u << (rs - 1u)) : 0;
    const int width = (int)((G < 32u) ? G : 32u);

    if (!lnb_ds_wait(sm, up, (g0 + 3u < g_end) ? g0 + 3u : g_end)) return false;
    LnbSyGroupState st;
#pragma unroll
    for (int s = 0; s < TT; s++) {
        st.c[s] = active ? (int32_t)coef[gl * TT + s] : 0;
        st.acc[s] = half; st.late[s] = 0;
    }
    st.y = io ? xu[0] : 0;
    st.d1 = (io && 1u < m) ? xu[1] : 0; st.d2 = (io && 2u < m) ? xu[2] : 0;
    st.yb = __shfl_sync(0xffffffffu, st.y, width - 1, width);

    uint32_t j = 0;
    while (j + 1u < m) {
        /* next batch: steps j .. j+B-1 read up to x[j+B+2] and finish outputs up to x[j+B] */
        const uint32_t left = m - 1u - j;
        const uint32_t B = left < LNB_DS_BATCH ? ((left + 3u) & ~3u) : LNB_DS_BATCH;
        const uint32_t need = g0 + j + B + 3u;
        if (!lnb_ds_wait(sm, up, need < g_end ? need : g_end)) return false;
        if (j + B + 3u < m) {
            for (uint32_t t = 0; t < B; t += 4u) {
                lnb_sy_group_step<0, false>(xu, j + t, m, p, rs, half, (uint32_t)width, io, is_first, st);
                lnb_sy_group_step<1, false>(xu, j + t + 1u, m, p, rs, half, (uint32_t)width, io, is_first, st);
                lnb_sy_group_step<2, false>(xu, j + t + 2u, m, p, rs, half, (uint32_t)width, io, is_first, st);
                lnb_sy_group_step<3, false>(xu, j + t + 3u, m, p, rs, half, (uint32_t)width, io, is_first, st);
            }
        } else {
            for (uint32_t t = 0; t < B; t += 4u) {
                lnb_sy_group_step<0, true>(xu, j + t, m, p, rs, half, (uint32_t)width, io, is_first, st);
                lnb_sy_group_step<1, true>(xu, j + t + 1u, m, p, rs, half, (uint32_t)width, io, is_first, st);
                lnb_sy_group_step<2, true>(xu, j + t + 2u, m, p, rs, half, (uint32_t)width, io, is_first, st);
                lnb_sy_group_step<3, true>(xu, j + t + 3u, m, p, rs, half, (uint32_t)width, io, is_first, st);
            }
        }
        j += B;
        const uint32_t fin = (j + 1u < m) ? j + 1u : m;
        lnb_ds_publish(sm, self, g0 + fin, lane);
    }
    lnb_ds_publish(sm, self, g0 + m, lane);
    return true;
}

/* ---- one unit, p = TT <= 4 taps: one lane, history in registers ---- */
template <int TT>
__device__ bool lnb_ds_unit_lane(LnbDsShared &sm, uint32_t up, uint32_t self, uint32_t g0, uint32_t g_end,
                                 int32_t *xu, uint32_t m, const int8_t *coef, uint32_t rs)
{
    const uint32_t lane = threadIdx.x & 31u;
    const bool active = lane == 0u;
    const uint32_t p = TT;
    const int32_t half = rs ? (int32_t)(1u << (rs - 1u)) : 0;
    if (!lnb_ds_wait(sm, up, (g0 + 3u < g_end) ? g0 + 3u : g_end)) return false;
    int32_t c[TT], acc[TT];
#pragma unroll
    for (int s = 0; s < TT; s++) { c[s] = (int32_t)coef[s]; acc[s] = half; }
    int32_t y = active ? xu[0] : 0;
    int32_t d1 = (active && 1u < m) ? xu[1] : 0, d2 = (active && 2u < m) ? xu[2] : 0;
    uint32_t j = 0;
    while (j + 1u < m) {
        const uint32_t left = m - 1u - j;
        const uint32_t B = left < LNB_DS_BATCH ? ((left + 3u) & ~3u) : LNB_DS_BATCH;      /* multiple of 4, hence of TT */
        const uint32_t need = g0 + j + B + 3u;
        if (!lnb_ds_wait(sm, up, need < g_end ? need : g_end)) return false;
        if (j + B + 3u < m) {
#pragma unroll 2
            for (uint32_t t = 0; t < B; t += TT) {
                lnb_sy_lane_step<TT, 0, false>(xu, j + t, m, p, rs, half, active, c, acc, y, d1, d2);
                if (TT > 1) lnb_sy_lane_step<TT, 1 % TT, false>(xu, j + t + 1u, m, p, rs, half, active, c, acc, y, d1, d2);
                if (TT > 2) lnb_sy_lane_step<TT, 2 % TT, false>(xu, j + t + 2u, m, p, rs, half, active, c, acc, y, d1, d2);
                if (TT > 2) lnb_sy_lane_step<TT, 3 % TT, false>(xu, j + t + 3u, m, p, rs, half, active, c, acc, y, d1, d2);
            }
        } else {
            for (uint32_t t = 0; t < B; t += TT) {
                lnb_sy_lane_step<TT, 0, true>(xu, j + t, m, p, rs, half, active, c, acc, y, d1, d2);
                if (TT > 1) lnb_sy_lane_step<TT, 1 % TT, true>(xu, j + t + 1u, m, p, rs, half, active, c, acc, y, d1, d2);
                if (TT > 2) lnb_sy_lane_step<TT, 2 % TT, true>(xu, j + t + 2u, m, p, rs, half, active, c, acc, y, d1, d2);
                if (TT > 2) lnb_sy_lane_step<TT, 3 % TT, true>(xu, j + t + 3u, m, p, rs, half, active, c, acc, y, d1, d2);
            }
        }
        j += B;
        const uint32_t fin = (j + 1u < m) ? j + 1u : m;
        lnb_ds_publish(sm, self, g0 + fin, lane);
    }
    lnb_ds_publish(sm, self, g0 + m, lane);
    return true;
}

/* ---- a synthesis layer of one channel, unit after unit ---- */
__device__ bool lnb_ds_layer(LnbDsShared &sm, uint32_t up, uint32_t self, uint32_t gbase, int32_t *x, uint32_t n,
                             uint32_t P, uint32_t U, const int8_t *coef, uint32_t rs)
{
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t g_end = gbase + n;
    const uint32_t p = (U <= P) ? P / U : 0u, m = n / U;
    if (U > P || m <= p) {                                    /* nothing is predicted: the stage only passes the samples on */
        if (!lnb_ds_wait(sm, up, g_end)) return false;
        lnb_ds_publish(sm, self, g_end, lane);
        return true;
    }
    for (uint32_t u = 0; u < U; u++) {
        int32_t *xu = x + (size_t)u * m;
        const int8_t *cu = coef + u * p;
        const uint32_t g0 = gbase + u * m;
        bool ok;
        if (p >= 8u) ok = lnb_ds_unit_group(sm, up, self, g0, g_end, xu, m, p, cu, rs);
        else if (p == 4u) ok = lnb_ds_unit_lane<4>(sm, up, self, g0, g_end, xu, m, cu, rs);
        else if (p == 2u) ok = lnb_ds_unit_lane<2>(sm, up, self, g0, g_end, xu, m, cu, rs);
        else ok = lnb_ds_unit_lane<1>(sm, up, self, g0, g_end, xu, m, cu, rs);
        if (!ok) return false;
    }
    if (!lnb_ds_wait(sm, up, g_end)) return false;            /* samples behind the last unit are copied (n mod U) */
    lnb_ds_publish(sm, self, g_end, lane);
    return true;
}

/* ---- last stage: de-emphasis of one channel, M/S inverse, coalesced stores to the PCM planes ---- */
__device__ bool lnb_ds_finish(LnbDsShared &sm, const LnbDecodeBatch &b, const LnbBlockDesc &blk, uint32_t up, uint32_t self,
                              uint32_t ch, int32_t *x, uint32_t n)
{
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t gbase = ch * n, g_end = gbase + n;
    const LnbChanParams &prm = sm.params[ch];
    const int32_t c0 = prm.preem_coef[0], c1 = prm.preem_coef[1];
    int32_t zp = prm.preem_prev[1], yp = prm.preem_prev[0];
    int32_t *gx = b.pcm + (size_t)ch * b.cfg.pcm_stride + blk.smp_off;
    const bool ms = b.cfg.ms && b.cfg.num_channels >= 2u && ch == 1u;
    int32_t *g0 = b.pcm + blk.smp_off;                        /* channel 0 (mid), already stored by this warp */
    for (uint32_t i0 = 0; i0 < n; i0 += LNB_DS_BATCH) {
        const uint32_t k = (n - i0 < LNB_DS_BATCH) ? n - i0 : LNB_DS_BATCH;
        if (!lnb_ds_wait(sm, up, gbase + i0 + k)) return false;
        if (lane == 0) {
#pragma unroll 8
            for (uint32_t i = i0; i < i0 + k; i++) {
                const int32_t z = x[i] + ((zp * c1) >> LNB_PREEM_SHIFT);
                const int32_t yv = z + ((yp * c0) >> LNB_PREEM_SHIFT);
                x[i] = yv; zp = z; yp = yv;
            }
        }
        __syncwarp();
        if (lane < k) {
            int32_t v = x[i0 + lane];
            if (ms) {                                          /* linne_utility.c:143-146 */
                int32_t mid = g0[i0 + lane];
                lnb_ms_to_lr(mid, v);
                g0[i0 + lane] = mid;
            }
            gx[i0 + lane] = v;
        }
        lnb_ds_publish(sm, self, gbase + i0 + k, lane);
    }
    (void)g_end;
    return true;
}

/* One CTA per block.  Dynamic shared memory: the channel line, n_max int32. */
__global__ void __launch_bounds__(LNB_DS_THREADS) lnb_stream_v1_kernel(LnbDecodeBatch b, uint32_t n_max)
{
    extern __shared__ __align__(16) int32_t lnb_ds_line[];
    __shared__ LnbDsShared sm;
    /* Warps map to the SM's four schedulers by index mod 4.  The entropy warp (role 0) is the pipeline's pace maker:
     * it gets scheduler 0 to itself (hardware warp 4 exits), the stages take warps 1, 2, 3 and 5. */
    const uint32_t hw_warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const uint32_t warp = (hw_warp < 4u) ? hw_warp : (hw_warp == 4u ? 0xFFu : 4u);        /* role: 0 entropy, 1.. stages */
    const uint32_t blk_i = blockIdx.x;
    LnbBlockDesc &gblk = b.blocks[blk_i];
    const LnbBlockDesc blk = gblk;
    const LnbStreamCfg &cfg = b.cfg;
    const uint32_t C = cfg.num_channels, n = blk.nsmp, L = cfg.num_layers;
    /* blocks this kernel does not take are left to the split kernels (same rule on both sides) */
    if (blk.type != LNB_BLOCK_COMPRESSED || blk.status || n > n_max || n == 0u || lnb_tp_takes(b, blk)) return;

    for (uint32_t i = threadIdx.x; i < (1u << LNB_E3_HUFF1_BITS); i += LNB_DS_THREADS) {
        const uint16_t e = b.tab.huff_lut[i << (LNB_HUFF_LUT_BITS - LNB_E3_HUFF1_BITS)];
        sm.huff1[i] = ((e & 15u) <= LNB_E3_HUFF1_BITS) ? e : (uint16_t)0;
    }
    if (threadIdx.x < LNB_DS_WARPS) sm.prog[threadIdx.x] = 0u;
    if (threadIdx.x == 0) sm.abort = 0u;
    __syncthreads();

    const uint32_t last = L + 1u;                              /* stage index of the de-emphasis warp */
    if (warp == 0xFFu) return;
    if (warp == 0u) {
        LnbE3Win win;
        win.buf = sm.win;
        LnbDsSink sink;
        sink.params = sm.params; sink.sm = &sm; sink.line = lnb_ds_line; sink.last_stage = last;
        lnb_e3_compressed_block(b, gblk, blk, win, sm.huff1, sink, lane);
        if (sm.abort) {                                        /* broken payload: the block reads as silence */
            for (uint32_t c = 0; c < C; c++) {
                int32_t *gout = b.pcm + (size_t)c * cfg.pcm_stride + blk.smp_off;
                for (uint32_t i = lane; i < n; i += 32u) gout[i] = 0;
            }
        }
        return;
    }
    if (warp > last) return;
    /* side information is complete once the entropy warp has published anything at all; stage warps of the
     * first channel wait for samples anyway, the parameters are read after that wait */
    for (uint32_t ch = 0; ch < C; ch++) {
        const uint32_t gbase = ch * n;
        if (!lnb_ds_wait(sm, warp - 1u, gbase + 1u)) return;
        if (warp <= L) {
            const uint32_t l = L - warp;                       /* layers run L-1 .. 0 (linne_decoder.c:503-509) */
            const LnbChanParams &prm = sm.params[ch];
            if (!lnb_ds_layer(sm, warp - 1u, warp, gbase, lnb_ds_line, n, cfg.layer_params[l], 1u << prm.log2_units[l],
                              prm.coef + l * LNB_MAX_PARAMS, prm.rshift[l])) return;
        } else {
            if (!lnb_ds_finish(sm, b, blk, warp - 1u, warp, ch, lnb_ds_line, n)) return;
        }
    }
}
