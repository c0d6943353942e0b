/* lnb_tput_v1.cuh -- throughput decoder for large batches: one THREAD per block (entropy) and one thread
 * per (block, channel) (synthesis cascade + de-emphasis + M/S), 32 independent sequences per warp.
 *
 * Covers the same reference rows as the fused streaming kernel (lnb_stream_v1.cuh): d2/d3 entropy decode
 * (libs/linne_decoder/src/linne_decoder.c:457-497, libs/linne_coder/src/linne_coder.c:306-327), d4 synthesis
 * (libs/linne_decoder/src/linne_lpc_synthesize.c:8-83, layer order of linne_decoder.c:503-509), d5 de-emphasis
 * and mid/side inverse (libs/linne_internal/src/linne_utility.c:215-241, :135-147).
 *
 * Why a second decoder.  The format is serial inside a block (one entry point, channels concatenated, Rice
 * parameters delta-coded inline; synthesis is a recursion over the samples), so lnb_stream_v1 spends a CTA of
 * five warps per block to shorten the LATENCY of that chain -- right for a 10-second clip (44 blocks), but at
 * ~145 warp instructions per sample it is issue-bound on an hour of audio (15 504 blocks).  With tens of
 * thousands of independent sequences in a batch the serial chain can simply stay serial: every lane walks its
 * own sequence, a warp retires 32 samples per pass of the loop, and the cost drops to the ~1 (entropy) and
 * ~P/4 (synthesis) warp instructions per sample the arithmetic needs.
 *
 *   lnb_tp_entropy_kernel   lane = block.  Side information and the recursive-Rice residuals of all channels, read
 *                           through the prefetching bit reader (every lane streams its own payload through L1);
 *                           residuals are transposed through a padded shared-memory tile so that every global
 *                           store is a full 128-byte row of one sequence.
 *   lnb_tp_synth_kernel     lane = (block, channel).  Groups of 8 samples run through the whole cascade in
 *                           registers: for layers of 16..128 taps the history lives in a lane-interleaved
 *                           shared-memory ring (conflict-free; the ring position is warp-uniform because all
 *                           lanes are at the same sample index) and each history sample is loaded once for 8
 *                           outputs (8 IMADs per LDS pair); 2..8-tap layers keep history and taps in registers.
 *                           Unit counts differ between lanes: every lane runs the layer's full tap count with the
 *                           taps of its current unit right-aligned and zero-padded, rebuilt when it enters a unit.
 *                           De-emphasis in registers, M/S with the neighbouring lane by shuffle, 2 x 16-byte
 *                           loads and stores per group and lane.
 *
 * Takes full blocks only (nsmp == block size, a multiple of 1024: every unit length is a multiple of 8 and all
 * lanes of a warp run the same trip counts); tail, raw, silent and over-long blocks stay with the other kernels,
 * which skip what is taken here (lnb_tp_takes).
 */
#pragma once
#include "lnb_common.cuh"
#include "lnb_decode_core.cuh"

#define LNB_TP_TILE_STRIDE 33u

/* same rule on every side: the host only sets `tput`, the kernels decide per block */
LNB_HD bool lnb_tp_takes(const LnbDecodeBatch &b, const LnbBlockDesc &blk)
{
    return b.tput && blk.type == LNB_BLOCK_COMPRESSED && !(blk.status & (LNB_ST_CRC_MISMATCH | LNB_ST_BAD_TYPE))
        && blk.nsmp == b.cfg.block_size && blk.nsmp != 0u && (blk.nsmp & 1023u) == 0u && (blk.smp_off & 3u) == 0u;
}

#if defined(__CUDACC__)

/* ------------------------------------------------------------------------------------------------------
 * entropy: lane = block
 * ------------------------------------------------------------------------------------------------------ */
__global__ void __launch_bounds__(32) lnb_tp_entropy_kernel(LnbDecodeBatch b)
{
    __shared__ int32_t tile[32u * LNB_TP_TILE_STRIDE];
    const uint32_t lane = threadIdx.x;
    const uint32_t blk_i = blockIdx.x * 32u + lane;
    const LnbStreamCfg &cfg = b.cfg;
    const uint32_t C = cfg.num_channels, n = cfg.block_size;
    const bool have = blk_i < b.num_blocks;
    LnbBlockDesc blk;
    if (have) blk = b.blocks[blk_i];
    const bool mine = have && lnb_tp_takes(b, blk);
    if (__ballot_sync(0xffffffffu, mine) == 0u) return;

    LnbFastReader fr;
    uint32_t payload_bit = 0, rel_end_byte = 0, overrun = mine ? 0u : 1u;
    if (mine) {
        const uint32_t payload_off = blk.byte_off + LNB_BLOCK_HEADER_SIZE;
        uint32_t end_byte = blk.byte_off + blk.byte_size;
        if (end_byte > b.stream_size) end_byte = b.stream_size;
        /* positions relative to the aligned word holding the block's first byte (32-bit bit positions) */
        const uint32_t word0 = blk.byte_off >> 2;
        payload_bit = (payload_off - word0 * 4u) * 8u;
        rel_end_byte = end_byte - word0 * 4u;
        lnb_fr_open(fr, (const uint32_t *)b.stream + word0, payload_bit, (rel_end_byte + 3u) >> 2);
    } else {
        fr.words = (const uint32_t *)b.stream; fr.end_word = 0; fr.overrun = 0;
        fr.hi = fr.lo = fr.pre = 0; fr.nbits = 64u; fr.next = 2u;
    }

    /* ---- side information (linne_decoder.c:457-486) ---- */
    if (mine) {
        LnbChanParams *params = b.params + (size_t)blk_i * C;
        for (uint32_t c = 0; c < C; c++)
            for (int f = 0; f < LNB_NUM_PREEM; f++) {
                params[c].preem_prev[f] = lnb_zz_dec(lnb_fr_get(fr, cfg.bits_per_sample + 1u));
                params[c].preem_coef[f] = (uint8_t)lnb_fr_get(fr, LNB_PREEM_SHIFT - 1);
            }
        for (uint32_t c = 0; c < C; c++)
            for (uint32_t l = 0; l < cfg.num_layers; l++) {
                params[c].log2_units[l] = (uint8_t)lnb_fr_get(fr, 3);
                params[c].rshift[l] = (uint8_t)lnb_fr_get(fr, 4);
                int8_t *q = params[c].coef + l * LNB_MAX_PARAMS;
                const uint32_t P = cfg.layer_params[l];
                for (uint32_t i = 0; i < P; i += 4u) {               /* P is a multiple of 4 or equals 2 */
                    uint32_t packed = 0;
                    const uint32_t lim = (P - i < 4u) ? P - i : 4u;
                    for (uint32_t t = 0; t < lim; t++) {
                        const uint32_t e = b.tab.huff_lut[fr.hi >> (32 - LNB_HUFF_LUT_BITS)];
                        lnb_fr_skip(fr, e & 15u);
                        packed |= ((uint32_t)lnb_zz_dec(e >> 4) & 0xFFu) << (8u * t);
                    }
                    if (lim == 4u) *(uint32_t *)(q + i) = packed;
                    else for (uint32_t t = 0; t < lim; t++) q[i + t] = (int8_t)(packed >> (8u * t));
                }
            }
    }

    /* ---- residuals, channel after channel; all lanes walk the same sample index ---- */
    for (uint32_t c = 0; c < C; c++) {
        int32_t *dst = b.pcm + (size_t)c * cfg.pcm_stride + (mine ? blk.smp_off : 0u);
        uint32_t porder = overrun ? 0u : lnb_fr_get(fr, 10);
        if (porder > LNB_MAX_PORDER) { overrun = 1u; porder = 0u; }
        const uint32_t len = n >> porder;
        uint32_t k2 = 0, left = 0, first = 1u;
        for (uint32_t i0 = 0; i0 < n; i0 += 32u) {
#pragma unroll 2
            for (uint32_t r = 0; r < 32u; r++) {
                uint32_t u = 0;
                if (!overrun) {
                    if (left == 0u) {                              /* partition header (linne_coder.c:311-318) */
                        if (first) { k2 = lnb_fr_get(fr, 5); first = 0u; }
                        else k2 = (uint32_t)((int32_t)k2 + lnb_zz_dec(lnb_get_gamma(fr)));
                        if (k2 > 30u) { overrun = 1u; k2 = 30u; }
                        left = len;
                    }
                    u = lnb_get_rice(fr, k2 + 1u, k2);
                    left--;
                    overrun |= fr.overrun;
                }
                tile[r * LNB_TP_TILE_STRIDE + lane] = overrun ? 0 : lnb_zz_dec(u);
            }
            __syncwarp();
            /* row l of the transposed tile = 32 consecutive samples of lane l's sequence: one 128-byte store each */
#pragma unroll 4
            for (uint32_t l = 0; l < 32u; l++) {
                const unsigned long long p = __shfl_sync(0xffffffffu, (unsigned long long)(uintptr_t)dst, (int)l);
                const bool on = (__shfl_sync(0xffffffffu, mine ? 1u : 0u, (int)l)) != 0u;
                if (on) ((int32_t *)(uintptr_t)p)[i0 + lane] = tile[lane * LNB_TP_TILE_STRIDE + l];
            }
            __syncwarp();
        }
        overrun |= fr.overrun;
    }
    if (mine) {
        const uint32_t used = (uint32_t)((lnb_fr_position(fr) - payload_bit + 7u) >> 3);
        b.blocks[blk_i].na = used;                                  /* payload bytes consumed (reference Flush + Tell) */
        if (overrun || (payload_bit >> 3) + used > rel_end_byte) b.blocks[blk_i].status = blk.status | LNB_ST_OVERRUN;
    }
}

/* ------------------------------------------------------------------------------------------------------
 * synthesis cascade: lane = (block, channel)
 * ------------------------------------------------------------------------------------------------------ */
struct LnbTpLayerState {
    const int8_t *coef;         /* the layer's taps in the parameter record (global) */
    uint32_t p, m;              /* taps and samples per unit of THIS lane */
    uint32_t jj, unit;          /* position inside the current unit, index of the current unit */
    uint32_t rs;
    int32_t half;
    bool enabled;               /* this lane predicts at all (reference: units with m <= p are copied) */
};

__device__ __forceinline__ void lnb_tp_layer_init(LnbTpLayerState &s, const LnbChanParams *prm, uint32_t layer, uint32_t Q, uint32_t n, bool active)
{
    const uint32_t U = active ? (1u << prm->log2_units[layer]) : 1u;
    s.coef = active ? prm->coef + layer * LNB_MAX_PARAMS : nullptr;
    s.p = (U <= Q) ? Q / U : 0u;
    s.m = n / U;
    s.enabled = active && U <= Q && s.m > s.p;
    s.rs = active ? prm->rshift[layer] : 0u;
    s.half = s.rs ? (int32_t)(1u << (s.rs - 1u)) : 0;
    s.jj = 0; s.unit = 0;
    if (!s.enabled) { s.m = n; s.p = 0u; }
}

/* layer of Q >= 16 taps: history ring [Q + 8][32 lanes] and the current unit's taps [Q][32 lanes] in shared memory */
template <int Q>
__device__ __forceinline__ void lnb_tp_layer_ring(LnbTpLayerState &s, int32_t *ring, int16_t *cur, uint32_t &base, int32_t (&x)[8], uint32_t lane)
{
    constexpr uint32_t R = Q + 8;
    if (s.jj == 0u) {                                            /* entering a unit: right-aligned, zero-padded taps */
        const uint32_t pad = Q - s.p;
        const int8_t *cu = s.coef + s.unit * s.p;
        for (uint32_t a = 0; a < (uint32_t)Q; a++)
            cur[a * 32u + lane] = (s.enabled && a >= pad) ? (int16_t)cu[a - pad] : (int16_t)0;
    }
    int32_t acc[8], w[8];
#pragma unroll
    for (int j = 0; j < 8; j++) { acc[j] = s.half; w[j] = 0; }
    /* history x[i0 - Q + a], a = 0..Q-1, sits at ring slot (base + 8 + a) mod R; chunks of 8 never straddle the wrap */
    uint32_t slot = base + 8u;
#pragma unroll 2
    for (uint32_t a0 = 0; a0 < (uint32_t)Q; a0 += 8u) {
        if (slot >= R) slot -= R;
        const int32_t *xr = ring + slot * 32u + lane;
        const int16_t *cr = cur + a0 * 32u + lane;
#pragma unroll
        for (int t = 0; t < 8; t++) {
            const int32_t xv = xr[t * 32];
            w[t] = (int32_t)cr[t * 32];                           /* w[a & 7] = tap a; output j uses tap a - j */
#pragma unroll
            for (int j = 0; j < 8; j++) acc[j] += w[(t - j) & 7] * xv;   /* taps before the window: w starts cleared */
        }
        slot += 8u;
    }
    /* the 8 new samples: triangular part, w[t] = tap Q - 8 + t */
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const bool pred = s.enabled && (s.jj + (uint32_t)j >= s.p);
        const int32_t y = pred ? (int32_t)((uint32_t)x[j] - (uint32_t)(acc[j] >> s.rs)) : x[j];
        x[j] = y;
#pragma unroll
        for (int j2 = j + 1; j2 < 8; j2++) acc[j2] += w[8 - (j2 - j)] * y;
    }
    int32_t *xw = ring + base * 32u + lane;
#pragma unroll
    for (int j = 0; j < 8; j++) xw[j * 32] = x[j];
    base += 8u; if (base >= R) base -= R;
    s.jj += 8u;
    if (s.jj >= s.m) { s.jj = 0u; s.unit++; }
}

/* layer of Q <= 8 taps: taps and history in registers */
template <int Q>
struct LnbTpRegLayer { int32_t tap[Q]; int32_t hist[Q]; };

template <int Q>
__device__ __forceinline__ void lnb_tp_layer_reg(LnbTpLayerState &s, LnbTpRegLayer<Q> &rg, int32_t (&x)[8])
{
    if (s.jj == 0u) {
        const uint32_t pad = Q - s.p;
        const int8_t *cu = s.coef + s.unit * s.p;
#pragma unroll
        for (int a = 0; a < Q; a++) rg.tap[a] = (s.enabled && (uint32_t)a >= pad) ? (int32_t)cu[(uint32_t)a - pad] : 0;
    }
    int32_t v[Q + 8];
#pragma unroll
    for (int a = 0; a < Q; a++) v[a] = rg.hist[a];
#pragma unroll
    for (int j = 0; j < 8; j++) {
        int32_t acc = s.half;
#pragma unroll
        for (int a = 0; a < Q; a++) acc += rg.tap[a] * v[j + a];
        const bool pred = s.enabled && (s.jj + (uint32_t)j >= s.p);
        v[Q + j] = pred ? (int32_t)((uint32_t)x[j] - (uint32_t)(acc >> s.rs)) : x[j];
        x[j] = v[Q + j];
    }
#pragma unroll
    for (int a = 0; a < Q; a++) rg.hist[a] = v[8 + a];
    s.jj += 8u;
    if (s.jj >= s.m) { s.jj = 0u; s.unit++; }
}

template <int Q> struct LnbTpLayer {
    LnbTpLayerState st;
    LnbTpRegLayer<(Q >= 1 && Q <= 8) ? Q : 1> rg;
    int32_t *ring; int16_t *cur; uint32_t base;
    static constexpr size_t smem_bytes = (Q >= 16) ? (size_t)(Q + 8) * 32u * 4u + (size_t)Q * 32u * 2u : 0u;
    __device__ __forceinline__ void init(uint8_t *&smem, const LnbChanParams *prm, uint32_t layer, uint32_t n, bool active)
    {
        if (Q == 0) return;
        lnb_tp_layer_init(st, prm, layer, (uint32_t)Q, n, active);
        base = 0;
        if (Q >= 16) {
            ring = (int32_t *)smem; smem += (size_t)(Q + 8) * 32u * 4u;
            cur = (int16_t *)smem; smem += (size_t)Q * 32u * 2u;
            for (uint32_t a = 0; a < (uint32_t)(Q + 8); a++) ring[a * 32u + (threadIdx.x & 31u)] = 0;
        } else {
#pragma unroll
            for (int a = 0; a < ((Q >= 1 && Q <= 8) ? Q : 1); a++) { rg.hist[a] = 0; rg.tap[a] = 0; }
        }
    }
    __device__ __forceinline__ void step(int32_t (&x)[8], uint32_t lane)
    {
        if (Q == 0) return;
        if (Q >= 16) lnb_tp_layer_ring<(Q >= 16) ? Q : 16>(st, ring, cur, base, x, lane);
        else lnb_tp_layer_reg<(Q >= 1 && Q <= 8) ? Q : 1>(st, rg, x);
    }
};

/* Q0, Q1, Q2 = taps of the layers in synthesis order (layer L-1 first); 0 = no such layer */
template <int Q0, int Q1, int Q2>
__global__ void __launch_bounds__(32) lnb_tp_synth_kernel(LnbDecodeBatch b)
{
    extern __shared__ __align__(16) uint8_t lnb_tp_smem[];
    const uint32_t lane = threadIdx.x;
    const LnbStreamCfg &cfg = b.cfg;
    const uint32_t C = cfg.num_channels, n = cfg.block_size, L = cfg.num_layers;
    const uint32_t seq = blockIdx.x * 32u + lane;
    const uint32_t blk_i = seq / C, ch = seq % C;
    const bool have = blk_i < b.num_blocks;
    LnbBlockDesc blk;
    if (have) blk = b.blocks[blk_i];
    const bool taken = have && lnb_tp_takes(b, blk);
    if (__ballot_sync(0xffffffffu, taken) == 0u) return;
    const bool active = taken && !(blk.status & LNB_ST_OVERRUN);
    int32_t *gx = b.pcm + (size_t)ch * cfg.pcm_stride + (taken ? blk.smp_off : 0u);
    if (taken && !active) {                                      /* broken payload: the block reads as silence */
        for (uint32_t i = 0; i < n; i += 4u) *(int4 *)(gx + i) = make_int4(0, 0, 0, 0);
    }
    const LnbChanParams *prm = b.params + (size_t)blk_i * C + ch;

    uint8_t *smem = lnb_tp_smem;
    LnbTpLayer<Q0> l0; LnbTpLayer<Q1> l1; LnbTpLayer<Q2> l2;
    l0.init(smem, prm, L - 1u, n, active);
    l1.init(smem, prm, L - 2u, n, active);
    if (Q2) l2.init(smem, prm, L - 3u, n, active);
    const int32_t c0 = active ? prm->preem_coef[0] : 0, c1 = active ? prm->preem_coef[1] : 0;
    int32_t zp = active ? prm->preem_prev[1] : 0, yp = active ? prm->preem_prev[0] : 0;
    const bool ms = cfg.ms && C >= 2u && ch < 2u;
    __syncwarp();

    int4 nxt0 = make_int4(0, 0, 0, 0), nxt1 = nxt0;
    if (active) { nxt0 = *(const int4 *)gx; nxt1 = *(const int4 *)(gx + 4); }
    for (uint32_t i0 = 0; i0 < n; i0 += 8u) {
        int32_t x[8] = {nxt0.x, nxt0.y, nxt0.z, nxt0.w, nxt1.x, nxt1.y, nxt1.z, nxt1.w};
        if (active && i0 + 8u < n) { nxt0 = *(const int4 *)(gx + i0 + 8u); nxt1 = *(const int4 *)(gx + i0 + 12u); }
        l0.step(x, lane);
        l1.step(x, lane);
        if (Q2) l2.step(x, lane);
#pragma unroll
        for (int j = 0; j < 8; j++) {                            /* linne_utility.c:215-241 */
            const int32_t z = x[j] + ((zp * c1) >> LNB_PREEM_SHIFT);
            const int32_t y = z + ((yp * c0) >> LNB_PREEM_SHIFT);
            x[j] = y; zp = z; yp = y;
        }
        if (cfg.ms && C >= 2u) {                                 /* lanes 2k / 2k+1 hold mid / side of one block (C even) */
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const int32_t other = __shfl_xor_sync(0xffffffffu, x[j], 1);
                if (ms) {
                    int32_t mid = (ch == 0u) ? x[j] : other, side = (ch == 0u) ? other : x[j];
                    lnb_ms_to_lr(mid, side);
                    x[j] = (ch == 0u) ? mid : side;
                }
            }
        }
        if (active) {
            *(int4 *)(gx + i0) = make_int4(x[0], x[1], x[2], x[3]);
            *(int4 *)(gx + i0 + 4u) = make_int4(x[4], x[5], x[6], x[7]);
        }
    }
}

#endif /* __CUDACC__ */
