/* lnb_tput_v2.cuh -- throughput decoder for large batches: one THREAD per block (entropy) and one thread
 * per (block, channel) (synthesis cascade + de-emphasis + M/S), 32 independent sequences per warp.
 *
 * Covers the same reference rows as the fused streaming kernel (lnb_stream_v2.cuh): d2/d3 entropy decode
 * (libs/linne_decoder/src/linne_decoder.c:457-497, libs/linne_coder/src/linne_coder.c:306-327), d4 synthesis
 * (libs/linne_decoder/src/linne_lpc_synthesize.c:8-83, layer order of linne_decoder.c:503-509), d5 de-emphasis
 * and mid/side inverse (libs/linne_internal/src/linne_utility.c:215-241, :135-147).
 *
 * Why a second decoder.  The format is serial inside a block (one entry point, channels concatenated, Rice
 * parameters delta-coded inline; synthesis is a recursion over the samples), so lnb_stream_v2 spends a CTA of
 * six warps per block to shorten the LATENCY of that chain -- right for a 10-second clip (44 blocks), but at
 * ~145 warp instructions per sample it is issue-bound on an hour of audio (15 504 blocks).  With tens of
 * thousands of independent sequences in a batch the serial chain can simply stay serial: every lane walks its
 * own sequence, a warp retires 32 samples per pass of the loop, and the cost drops to the ~1 (entropy) and
 * ~P/4 (synthesis) warp instructions per sample the arithmetic needs.
 *
 *   lnb_tp_entropy_kernel   eight lanes per block, four blocks per warp: speculative code-word starts, one 32-byte
 *                           sector of residuals per round; with ITER (the default) a round goes on past a long code
 *                           word -- the lanes behind it shift inside their 64-bit windows by its extra bits and look
 *                           again -- so all eight code words retire (see below).  How many lanes a block should get
 *                           was measured at every width in round 2 (profiles/r2_tp_entropy_forms.md; all forms stay
 *                           selectable through LINNE_B200_TP_ENTROPY and are tested): a warp per block
 *                           (lnb_tp_entropy_w_kernel, 11 warp instructions per code word, issue-bound, 3.9 ms on a
 *                           1-hour stream), eight lanes with fix-up rounds (6.6, 3.65 ms: the default), eight lanes
 *                           without (26, 4.3 ms: round 1), one lane per block with everything staged through shared
 *                           memory (lnb_tp_entropy_l_kernel / _s_kernel, 2.8 instructions per code word but one warp
 *                           per scheduler: a lone warp's ~5.4 cycles per dependent instruction, 6 ms).
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
#include "lnb_rice_warp.cuh"

/* same rule on every side: the host only sets `tput`, the kernels decide per block */
LNB_HD bool lnb_tp_shape_ok(const LnbDecodeBatch &b, const LnbBlockDesc &blk)
{
    return b.tput && blk.type == LNB_BLOCK_COMPRESSED
        && blk.nsmp == b.cfg.block_size && blk.nsmp != 0u && (blk.nsmp & 1023u) == 0u && (blk.smp_off & 3u) == 0u;
}
/* ... and the block's CRC check did not fail (the entropy stage runs beside the CRC pass and does not look) */
LNB_HD bool lnb_tp_takes(const LnbDecodeBatch &b, const LnbBlockDesc &blk)
{
    return lnb_tp_shape_ok(b, blk) && !(blk.status & (LNB_ST_CRC_MISMATCH | LNB_ST_BAD_TYPE));
}

#if defined(__CUDACC__)

/* ------------------------------------------------------------------------------------------------------
 * entropy: LNB_TG lanes per block, 32 / LNB_TG blocks per warp.
 *
 * One lane per block leaves the machine empty (a 1-hour stream has 15 504 blocks = 485 warps) and a lane's
 * symbol chain of ~90 dependent instructions runs at one instruction per ~7 cycles when nothing else shares
 * its scheduler; a whole warp per block (lnb_entropy_v3.cuh) retires ~3 code words per round on 32 lanes.  Eight
 * lanes per block sit between: with k1 = k2 + 1 the code words of residuals below 3 * 2^k2 -- the large
 * majority -- are all k2 + 2 bits long, so lane j of a group reads the field at  pos + j * (k2 + 2);  the first
 * lane that sees a longer code word ('00' prefix) ends the round, every lane before it holds a valid residual
 * (~4 per round), and the group's eight stores form one 32-byte sector.  Groups of a warp run the same flat
 * loop (one round per pass, headers and window refills as short predicated detours), so the warp does not
 * serialise them, and 3 876 warps keep every scheduler busy.
 * The payload of a group is staged in a shared-memory window of LNB_TG_WIN words filled by the group itself.
 * ------------------------------------------------------------------------------------------------------ */
#define LNB_TG            8u
#define LNB_TG_PER_WARP   (32u / LNB_TG)
#define LNB_TG_WIN        128u
#define LNB_TG_ROUND_BITS (LNB_TG * 33u + 64u + 32u) /* furthest bit a partition header plus a round can look at, relative to their start */

struct LnbTgWin {
    uint32_t *buf;              /* [LNB_TG_WIN + 2] */
    const uint32_t *words;      /* global, word 0 = aligned word holding the block's first byte */
    uint32_t wb;                /* index of the word held in buf[0] */
    uint32_t end_word;          /* words at or past this index read as zero */
    uint32_t limit;             /* a round may start at bit positions up to this one without a refill */
};
__device__ __forceinline__ void lnb_tg_fill(LnbTgWin &w, uint32_t word_idx, uint32_t lg, uint32_t gmask)
{
    __syncwarp(gmask);
    w.wb = word_idx;
    w.limit = (word_idx + LNB_TG_WIN) * 32u - LNB_TG_ROUND_BITS;
#pragma unroll 4
    for (uint32_t i = lg; i < LNB_TG_WIN + 2u; i += LNB_TG) {
        const uint32_t idx = word_idx + i;
        w.buf[i] = (idx < w.end_word) ? lnb_bswap32(w.words[idx]) : 0u;
    }
    __syncwarp(gmask);
}
__device__ __forceinline__ void lnb_tg_ensure(LnbTgWin &w, uint32_t pos, uint32_t span, uint32_t lg, uint32_t gmask)
{
    if (((pos + span) >> 5) + 2u > w.wb + LNB_TG_WIN + 2u || (pos >> 5) < w.wb) lnb_tg_fill(w, pos >> 5, lg, gmask);
}
__device__ __forceinline__ uint32_t lnb_tg_peek(const LnbTgWin &w, uint32_t pos)
{
    const uint32_t i = (pos >> 5) - w.wb;
    return __funnelshift_l(w.buf[i + 1u], w.buf[i], pos & 31u);
}
__device__ __forceinline__ uint32_t lnb_tg_get(const LnbTgWin &w, uint32_t &pos, uint32_t n)   /* 1 <= n <= 32 */
{
    const uint32_t v = lnb_tg_peek(w, pos) >> (32u - n);
    pos += n;
    return v;
}

/* ITER: a round does not end at the first long code word -- the lanes behind it shift inside their 64-bit windows by its
 * extra bits and look again (the scheme of lnb_rice_warp.cuh at a width of eight; the four groups of a warp iterate
 * together, so one pass of the fix-up loop serves four blocks). */
template <bool ITER>
__global__ void __launch_bounds__(32) lnb_tp_entropy_kernel(LnbDecodeBatch b)
{
    __shared__ uint32_t s_win[LNB_TG_PER_WARP][LNB_TG_WIN + 2u];
    const uint32_t lane = threadIdx.x, g = lane / LNB_TG, lg = lane % LNB_TG;
    const uint32_t gmask = ((1u << LNB_TG) - 1u) << (g * LNB_TG);
    const uint32_t blk_i = blockIdx.x * LNB_TG_PER_WARP + g;
    const LnbStreamCfg &cfg = b.cfg;
    const uint32_t C = cfg.num_channels, n = cfg.block_size;
    LnbBlockDesc blk;
    bool active = blk_i < b.num_blocks;
    if (active) { blk = b.blocks[blk_i]; active = lnb_tp_shape_ok(b, blk); }
    if (__ballot_sync(0xffffffffu, active) == 0u) return;

    LnbTgWin win;
    win.buf = s_win[g];
    win.words = (const uint32_t *)b.stream; win.wb = 0; win.end_word = 0; win.limit = 0;
    uint32_t pos = 0, rel_payload = 0, rel_end = 0, overrun = 0;

    /* ---- side information (linne_decoder.c:457-486): every lane of the group reads the same fields ---- */
    if (active) {
        const uint32_t payload_off = blk.byte_off + LNB_BLOCK_HEADER_SIZE;
        uint32_t end_byte = blk.byte_off + blk.byte_size;
        if (end_byte > b.stream_size) end_byte = b.stream_size;
        const uint32_t word0 = blk.byte_off >> 2;               /* bit positions relative to this word never overflow */
        win.words = (const uint32_t *)b.stream + word0;
        rel_payload = payload_off - word0 * 4u; rel_end = end_byte - word0 * 4u;
        win.end_word = (rel_end + 3u) >> 2;
        pos = rel_payload * 8u;
        lnb_tg_fill(win, pos >> 5, lg, gmask);
        LnbChanParams *params = b.params + (size_t)blk_i * C;
        for (uint32_t c = 0; c < C; c++)
            for (int f = 0; f < LNB_NUM_PREEM; f++) {
                lnb_tg_ensure(win, pos, 64u, lg, gmask);
                const int32_t prev = lnb_zz_dec(lnb_tg_get(win, pos, cfg.bits_per_sample + 1u));
                const uint32_t coef = lnb_tg_get(win, pos, LNB_PREEM_SHIFT - 1);
                if (lg == 0u) { params[c].preem_prev[f] = prev; params[c].preem_coef[f] = (uint8_t)coef; }
            }
        for (uint32_t c = 0; c < C; c++)
            for (uint32_t l = 0; l < cfg.num_layers; l++) {
                const uint32_t P = cfg.layer_params[l];
                lnb_tg_ensure(win, pos, 7u + P * 14u + 32u, lg, gmask);
                const uint32_t lu = lnb_tg_get(win, pos, 3), rs = lnb_tg_get(win, pos, 4);
                if (lg == 0u) { params[c].log2_units[l] = (uint8_t)lu; params[c].rshift[l] = (uint8_t)rs; }
                int8_t *q = params[c].coef + l * LNB_MAX_PARAMS;
                for (uint32_t i0 = 0; i0 < P; i0 += LNB_TG) {    /* lane j keeps coefficients j, j + 8, ... */
                    int32_t keep = 0;
                    const uint32_t lim = (P - i0 < LNB_TG) ? P - i0 : LNB_TG;
                    for (uint32_t i = 0; i < lim; i++) {
                        const uint32_t e = b.tab.huff_lut[lnb_tg_peek(win, pos) >> (32 - LNB_HUFF_LUT_BITS)];
                        pos += e & 15u;
                        if (i == lg) keep = lnb_zz_dec(e >> 4);
                    }
                    if (lg < lim) q[i0 + lg] = (int8_t)keep;
                }
            }
    }

    /* ---- residuals.  Outer pass: channel starts and the end of the block (twice per channel).  Inner loop: one
     *      round per group and pass, with the window refill and the partition header (a gamma-coded delta of k2) of
     *      the groups that need one as short predicated detours in front of it. ---- */
    uint32_t chan = 0, left = 0, parts_left = 0, len = 0, k2 = 0;
    uint32_t my_rel = 0, my_end = 0, k2mask = 0, sh = 31u;       /* per-partition constants of this lane */
    bool running = active;
    int32_t *wp = b.pcm;                                         /* where this lane's residual of the next round goes */
    while (__any_sync(0xffffffffu, running)) {
        if (running && left == 0u && parts_left == 0u) {                         /* next channel, or the end of the block */
            if (chan == C || overrun) {
                running = false;
            } else {
                lnb_tg_ensure(win, pos, 64u, lg, gmask);
                uint32_t porder = lnb_tg_get(win, pos, 10);
                if (porder > LNB_MAX_PORDER) { overrun = 1u; porder = 0u; }
                len = n >> porder; parts_left = (1u << porder) - 1u;
                k2 = lnb_tg_get(win, pos, 5);                                    /* first partition: k2 itself (linne_coder.c:313) */
                if (k2 > 30u) { overrun = 1u; k2 = 30u; }
                left = overrun ? 0u : len;
                if (overrun) parts_left = 0u;
                my_rel = lg * (k2 + 2u); my_end = my_rel + k2 + 1u; k2mask = (1u << k2) - 1u; sh = 31u - k2;
                wp = b.pcm + (size_t)chan * cfg.pcm_stride + blk.smp_off + lg;
                chan++;
            }
        }
        if (!__any_sync(0xffffffffu, running)) break;
        /* `running` does not change inside the round loop and a running group runs out of code words eventually */
        while (__all_sync(0xffffffffu, !running || left != 0u || parts_left != 0u)) {
            if (running && pos > win.limit) lnb_tg_fill(win, pos >> 5, lg, gmask);
            if (running && left == 0u) {                                         /* partition header: gamma code of zigzag(k2 - previous k2) */
                const uint32_t h = lnb_tg_peek(win, pos);
                const uint32_t lz = lnb_clz32(h);
                const uint32_t v = ((h << (lz & 15u)) >> (31u - (lz & 15u))) - 1u;
                pos += 2u * lz + 1u;
                k2 = (uint32_t)((int32_t)k2 + lnb_zz_dec(v));
                left = len; parts_left--;
                if (lz > 15u || k2 > 30u) { overrun = 1u; k2 = 30u; left = 0u; parts_left = 0u; }
                my_rel = lg * (k2 + 2u); my_end = my_rel + k2 + 1u; k2mask = (1u << k2) - 1u; sh = 31u - k2;
            }
            if constexpr (ITER) {
                /* the round: lane j keeps 64 bits at its guess pos + j * (k2 + 2) and an offset e inside them */
                const bool go = running && left != 0u;
                const uint32_t R = go ? (left < LNB_TG ? left : LNB_TG) : 0u;
                const uint32_t L = k2 + 2u;
                uint32_t hi = 0xFFFFFFFFu, lo = 0xFFFFFFFFu;
                if (go) {
                    const uint32_t p = pos + my_rel, i = (p >> 5) - win.wb;
                    const uint32_t b0 = win.buf[i], b1 = win.buf[i + 1u], b2 = win.buf[i + 2u];
                    hi = __funnelshift_l(b1, b0, p); lo = __funnelshift_l(b2, b1, p);
                }
                const uint32_t lgkey = lg << 8;
                uint32_t e = 0u, klive = (lg < R) ? lgkey : 0xFFFFFFFFu, v;
                int32_t t;
                for (;;) {
                    v = __funnelshift_lc(lo, hi, e);
                    t = (int32_t)__clz((int)v) - 2;                              /* extra - 1; -1: a short code word */
                    t = t < -1 ? -1 : t;
                    uint32_t key = klive | (uint32_t)t;
                    key = min(key, __shfl_xor_sync(0xffffffffu, key, 1));
                    key = min(key, __shfl_xor_sync(0xffffffffu, key, 2));
                    key = min(key, __shfl_xor_sync(0xffffffffu, key, 4));        /* the group's first live lane with a long code word */
                    if (__all_sync(0xffffffffu, key == 0xFFFFFFFFu)) break;
                    if (key != 0xFFFFFFFFu) { if (lgkey > key) e += (key & 255u) + 1u; else klive = 0xFFFFFFFFu; }
                }
                /* the first lane that cannot vouch for its code word ends the round (lnb_rw_round) */
                const bool invalid = e > 32u || t > 29 - (int32_t)k2 || lg >= R;
                const uint32_t m8 = (__ballot_sync(0xffffffffu, invalid) >> (g * LNB_TG)) & ((1u << LNB_TG) - 1u);
                const uint32_t n_ok = m8 ? (uint32_t)__ffs((int)m8) - 1u : LNB_TG;
                const uint32_t end = my_rel + e + L + (uint32_t)(t + 1);
                const uint32_t bits = __shfl_sync(0xffffffffu, end, (int)(g * LNB_TG + (n_ok ? n_ok - 1u : 0u)));
                if (lg < n_ok) {
                    const uint32_t lz = lnb_clz32(v), ml = (lz > 1u) ? lz : 1u;
                    const uint32_t low = (v >> ((sh - ml) & 31u)) & k2mask;
                    const uint32_t mult = lz ? lz + 1u : ((v >> 30) & 1u);
                    *wp = lnb_zz_dec((mult << k2) + low);
                }
                if (n_ok) { pos += bits; left -= n_ok; wp += n_ok; }
                if (go && n_ok < R) {                                            /* the code word behind them is read serially (rare) */
                    LnbFastReader fr;
                    lnb_fr_open(fr, win.words, pos, win.end_word);
                    const uint32_t u = lnb_get_rice(fr, k2 + 1u, k2);
                    if (lg == 0u) *wp = lnb_zz_dec(u);
                    pos = (uint32_t)lnb_fr_position(fr);
                    wp++; left--;
                    if (fr.overrun) { overrun = 1u; left = 0u; parts_left = 0u; }
                }
            } else {
                /* the round: lane j guesses that code word j starts at pos + j * (k2 + 2) */
                const bool go = running && left != 0u;
                const uint32_t last = (left < LNB_TG ? left : LNB_TG) - 1u;
                const uint32_t hi = go ? lnb_tg_peek(win, pos + my_rel) : 0xFFFFFFFFu;
                const uint32_t lz = lnb_clz32(hi);
                const uint32_t ml = (lz > 1u) ? lz : 1u;
                const bool resolves = go && (hi < 0x40000000u || lg == last) && lg <= last;
                const uint32_t is_short = (lz <= sh) ? 0x8000u : 0u;                 /* whole code word inside the 32-bit peek */
                uint32_t r = resolves ? ((lg << 16) | is_short | (my_end + ml)) : 0xFFFFFFFFu;
                r = __reduce_min_sync(gmask, r);                                     /* one redux per group (tiled-partition style mask) */
                if (go) {
                    const uint32_t first = r >> 16;
                    uint32_t n_ok = first + ((r >> 15) & 1u);
                    {
                        const uint32_t low = (hi >> (sh - ml)) & k2mask;
                        const uint32_t mult = lz ? lz + 1u : ((hi >> 30) & 1u);
                        if (lg < n_ok) *wp = lnb_zz_dec((mult << k2) + low);
                    }
                    if (__builtin_expect((r & 0x8000u) != 0u, 1)) {
                        pos += r & 0x7FFFu;
                    } else {                                                         /* code word longer than 32 bits: its lane finishes it serially */
                        uint32_t endl = 0, bad = 0;
                        if (lg == first) {
                            LnbFastReader fr;
                            lnb_fr_open(fr, win.words, pos + my_rel, win.end_word);
                            const uint32_t q = lnb_fr_zero_run(fr);
                            const uint32_t u = (q == 0u) ? lnb_fr_get(fr, k2 + 1u) : lnb_fr_get(fr, k2) + (2u << k2) + ((q - 1u) << k2);
                            *wp = lnb_zz_dec(u);
                            endl = (uint32_t)lnb_fr_position(fr);
                            bad = fr.overrun;
                        }
                        pos = __shfl_sync(gmask, endl, (int)(g * LNB_TG + first));
                        n_ok = first + 1u;
                        if (__shfl_sync(gmask, bad, (int)(g * LNB_TG + first))) { overrun = 1u; left = n_ok; parts_left = 0u; }
                    }
                    left -= n_ok; wp += n_ok;
                }
            }
        }
    }
    if (active && lg == 0u) {
        const uint32_t used = (pos - rel_payload * 8u + 7u) >> 3;
        b.blocks[blk_i].na = used;                                               /* payload bytes consumed (reference Flush + Tell) */
        if (overrun || rel_payload + used > rel_end) atomicOr(&b.blocks[blk_i].status, (uint32_t)LNB_ST_OVERRUN);
    }
}

/* ------------------------------------------------------------------------------------------------------
 * entropy, second form: ONE WARP per block, rounds of 32 code words (lnb_rice_warp.cuh).
 *
 * The eight-lane rounds above end at the first long code word (~3.1 of 8 lanes retire per ~80 instructions: ~26 warp
 * instructions per code word, the kernel is issue-bound).  Here a round goes on past a long code word -- the lanes behind
 * it shift inside their 64-bit windows by its extra bits and look again, one warp-wide minimum per long code word -- so
 * the per-round work (window loads, residual arithmetic, one coalesced 128-byte store, bookkeeping) is paid once per 32
 * code words.  The payload comes through a per-warp ring filled one 512-byte chunk ahead with 16-byte loads.
 * ------------------------------------------------------------------------------------------------------ */
#define LNB_TW_WARPS 4u

__global__ void __launch_bounds__(32 * LNB_TW_WARPS) lnb_tp_entropy_w_kernel(LnbDecodeBatch b)
{
    __shared__ __align__(16) uint32_t s_ring[LNB_TW_WARPS][LNB_RW_WORDS];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const uint32_t blk_i = blockIdx.x * LNB_TW_WARPS + warp;
    const LnbStreamCfg &cfg = b.cfg;
    const uint32_t C = cfg.num_channels, n = cfg.block_size;
    if (blk_i >= b.num_blocks) return;
    const LnbBlockDesc blk = b.blocks[blk_i];
    if (!lnb_tp_shape_ok(b, blk)) return;

    /* geometry of the block's payload: bit positions count from g0, the 16-byte aligned address at or below its first byte */
    const uintptr_t addr = (uintptr_t)(b.stream + blk.byte_off);
    const uint8_t *g0 = (const uint8_t *)(addr & ~(uintptr_t)15);
    const uint32_t rel0 = (uint32_t)(addr & 15u);
    const uint32_t rel_payload = rel0 + LNB_BLOCK_HEADER_SIZE;
    uint32_t end_byte = blk.byte_off + blk.byte_size;
    if (end_byte > b.stream_size || end_byte < blk.byte_off) end_byte = b.stream_size;
    const uint32_t rel_end = rel0 + (end_byte - blk.byte_off);
    const uint32_t end_word = (rel_end + 3u) >> 2;
    /* the image is followed by >= 16 readable bytes (lnb_types.h): whole 16-byte lines up to there */
    const uintptr_t readable = ((uintptr_t)(b.stream + b.stream_size) + 16u) & ~(uintptr_t)15;
    uint64_t want = ((uint64_t)rel_end + 15u) & ~(uint64_t)15;
    if (want > (uint64_t)(readable - (uintptr_t)g0)) want = (uint64_t)(readable - (uintptr_t)g0);
    LnbRwRing ring;
    lnb_rw_open(ring, s_ring[warp], g0, (uint32_t)(want >> 4), end_word, lane);
    const uint32_t pos_limit = (end_word + 4u) * 32u;            /* nothing sane reads past this */
    uint32_t lane_key;
    asm volatile("shl.b32 %0, %1, 8;" : "=r"(lane_key) : "r"(lane));

    uint32_t pos = rel_payload * 8u, overrun = 0u;
    /* ---- side information (linne_decoder.c:457-486): every lane reads the same fields ---- */
    {
        LnbChanParams *params = b.params + (size_t)blk_i * C;
        for (uint32_t c = 0; c < C; c++)
            for (int f = 0; f < LNB_NUM_PREEM; f++) {
                lnb_rw_ensure(ring, (pos >> 5) + 4u, lane);
                const int32_t prev = lnb_zz_dec(lnb_rw_get(ring, pos, cfg.bits_per_sample + 1u));
                const uint32_t coef = lnb_rw_get(ring, pos, LNB_PREEM_SHIFT - 1);
                if (lane == 0u) { params[c].preem_prev[f] = prev; params[c].preem_coef[f] = (uint8_t)coef; }
            }
        for (uint32_t c = 0; c < C; c++)
            for (uint32_t l = 0; l < cfg.num_layers; l++) {
                const uint32_t P = cfg.layer_params[l];
                lnb_rw_ensure(ring, ((pos + 7u + P * 16u) >> 5) + 2u, lane);     /* a coefficient code is at most 15 bits long */
                const uint32_t lu = lnb_rw_get(ring, pos, 3), rs = lnb_rw_get(ring, pos, 4);
                if (lane == 0u) { params[c].log2_units[l] = (uint8_t)lu; params[c].rshift[l] = (uint8_t)rs; }
                int8_t *q = params[c].coef + l * LNB_MAX_PARAMS;
                for (uint32_t i0 = 0; i0 < P; i0 += 32u) {       /* lane j keeps coefficients j, j + 32, ... */
                    int32_t keep = 0;
                    const uint32_t lim = (P - i0 < 32u) ? P - i0 : 32u;
                    for (uint32_t i = 0; i < lim; i++) {
                        const uint32_t e = b.tab.huff_lut[lnb_rw_peek(ring, pos) >> (32 - LNB_HUFF_LUT_BITS)];
                        pos += e & 15u;
                        if (i == lane) keep = lnb_zz_dec(e >> 4);
                    }
                    if (lane < lim) q[i0 + lane] = (int8_t)keep;
                }
            }
    }

    /* ---- residuals, channel after channel (linne_coder.c:306-327) ---- */
    for (uint32_t c = 0; c < C && !overrun; c++) {
        lnb_rw_ensure(ring, (pos >> 5) + 4u, lane);
        const uint32_t porder = lnb_rw_get(ring, pos, 10);
        uint32_t k2 = lnb_rw_get(ring, pos, 5);                   /* first partition: k2 itself (linne_coder.c:313) */
        if (porder > LNB_MAX_PORDER || k2 > 30u) { overrun = 1u; break; }
        const uint32_t len = n >> porder, parts = 1u << porder;
        int32_t *wp = b.pcm + (size_t)c * cfg.pcm_stride + blk.smp_off;
        for (uint32_t part = 0; part < parts; part++) {
            if (part) {                                          /* gamma code of zigzag(k2 - previous k2) */
                lnb_rw_ensure(ring, (pos >> 5) + 4u, lane);
                const uint32_t h = lnb_rw_peek(ring, pos);
                const uint32_t lz = lnb_clz32(h);
                if (lz > 15u) { overrun = 1u; break; }
                const uint32_t gv = ((h << lz) >> (31u - lz)) - 1u;
                pos += 2u * lz + 1u;
                k2 = (uint32_t)((int32_t)k2 + lnb_zz_dec(gv));
                if (k2 > 30u) { overrun = 1u; break; }
            }
            uint32_t left = len;
            while (left) {
                if (pos > pos_limit) { overrun = 1u; break; }
                const uint32_t R = left < 32u ? left : 32u;
                lnb_rw_ensure(ring, ((pos + 32u * (k2 + 2u) + 128u) >> 5) + 1u, lane);
                uint32_t v, bits;
                const uint32_t n_ok = lnb_rw_round(ring.saddr, LNB_RW_RING, pos, R, k2, lane_key, v, bits);
                if (lane < n_ok) wp[lane] = lnb_rw_value(v, k2);
                pos += bits; wp += n_ok; left -= n_ok;
                if (n_ok < R) {                                  /* rare: the next code word is read serially from global memory */
                    LnbFastReader fr;
                    lnb_fr_open(fr, (const uint32_t *)g0, pos, end_word);
                    const uint32_t u = lnb_get_rice(fr, k2 + 1u, k2);
                    if (lane == 0u) *wp = lnb_zz_dec(u);
                    pos = (uint32_t)lnb_fr_position(fr);
                    wp++; left--;
                    if (fr.overrun) { overrun = 1u; break; }
                }
            }
            if (overrun) break;
        }
    }
    if (lane == 0u) {
        const uint32_t used = (pos - rel_payload * 8u + 7u) >> 3;
        b.blocks[blk_i].na = used;                               /* payload bytes consumed (reference Flush + Tell) */
        if (overrun || rel_payload + used > rel_end) atomicOr(&b.blocks[blk_i].status, (uint32_t)LNB_ST_OVERRUN);
    }
}

/* ------------------------------------------------------------------------------------------------------
 * entropy, third form: ONE LANE per block, everything a lane touches staged through shared memory.
 *
 * A lane that walks its own block needs ~1 warp instruction per code word (32 code words per pass of a ~30-instruction
 * loop) -- an order of magnitude less than any form that spends several lanes on one block's serial chain.  The two
 * lane-per-block kernels measured earlier in round 2 lost (7.1 / 8.8 ms on the 1-hour stream) because every pass paid
 * global-memory round trips: 32 lanes reading 32 different streams and writing 32 different output lines.  Here
 *   - the payload of the 32 blocks comes in through cp.async (LDGSTS): per lane a ring of 128 words, interleaved
 *     [word][lane] so that 32 lanes at 32 different positions never meet in a bank; a block's next 64 words are
 *     requested by the whole warp (two coalesced 128-byte copies) a thousand code words before its lane gets there,
 *     and nobody waits for them: the copies of one check have landed by the next;
 *   - residuals go to a [32 samples][32 lanes] tile and leave it transposed, 128 contiguous bytes per block;
 *   - all lanes decode exactly one code word per pass, so channel ends and tile flushes are warp-uniform; only the
 *     partition headers (a gamma-coded delta of k2) and code words longer than 32 bits are divergent detours.
 * What is left on a pass is the lane's own chain (window from shared memory, leading zeros, length): ~100 cycles,
 * latency-bound with one warp per scheduler -- 20 480 passes per stereo block.
 * ------------------------------------------------------------------------------------------------------ */
#define LNB_TL_RING   128u                                     /* words per lane */
#define LNB_TL_HALF   64u                                      /* words per request */
#define LNB_TL_CHECK  16u                                      /* passes between refill checks */

struct LnbTlLane {
    const uint32_t *gw;         /* the 32-bit word holding the block's first byte */
    uint32_t end_word;          /* words at or past this index read as zero */
    uint32_t issued;            /* words [0, issued) of this block have been requested into the ring */
    uint32_t saddr;             /* shared address of ring word 0 of this lane */
};
__device__ __forceinline__ void lnb_tl_cp4(uint32_t dst, const void *src, uint32_t src_bytes)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" :: "r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
/* the warp requests the next LNB_TL_HALF words of every lane whose bit in `m` is set */
__device__ __forceinline__ void lnb_tl_request(LnbTlLane &st, uint32_t ring_saddr, uint32_t m, uint32_t lane)
{
    while (m) {
        const int l = __ffs((int)m) - 1;
        m &= m - 1u;
        const unsigned long long gp = __shfl_sync(0xffffffffu, (unsigned long long)(uintptr_t)st.gw, l);
        const uint32_t first = __shfl_sync(0xffffffffu, st.issued, l), endw = __shfl_sync(0xffffffffu, st.end_word, l);
        const uint32_t *g = (const uint32_t *)(uintptr_t)gp;
#pragma unroll
        for (uint32_t h = 0; h < LNB_TL_HALF; h += 32u) {
            const uint32_t w = first + h + lane;
            const bool in = w < endw;
            lnb_tl_cp4(ring_saddr + ((w % LNB_TL_RING) * 32u + (uint32_t)l) * 4u, in ? g + w : g, in ? 4u : 0u);
        }
        if ((int)lane == l) st.issued += LNB_TL_HALF;
    }
}
/* keep every lane at least LNB_TL_HALF words ahead of its reader (bit position pos); with `all` the caller wants
 * the data now (side information, after a jump), otherwise the copies of the previous check are waited for */
/* LAG: copy groups (one per check) that may still be in flight when a lazy check returns.  A lane reaches the words it
 * asks for no earlier than LNB_TL_HALF - (words read between two checks) words later, i.e. several checks on: waiting for
 * the previous check's copies (LAG 1) stalls every check for a DRAM round trip. */
template <int LAG>
__device__ __forceinline__ void lnb_tl_check(LnbTlLane &st, uint32_t ring_saddr, uint32_t pos, bool live, bool all, uint32_t lane)
{
    if (live && (pos >> 5) >= st.issued) st.issued = (pos >> 5) & ~(LNB_TL_HALF - 1u);      /* jumped past the ring (long code word) */
    for (;;) {
        const uint32_t m = __ballot_sync(0xffffffffu, live && (pos >> 5) + LNB_TL_HALF >= st.issued);
        if (m == 0u) break;
        lnb_tl_request(st, ring_saddr, m, lane);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    if (all) asm volatile("cp.async.wait_group 0;" ::: "memory");
    else asm volatile("cp.async.wait_group %0;" :: "n"(LAG) : "memory");
    __syncwarp();
}
__device__ __forceinline__ uint32_t lnb_tl_peek(const LnbTlLane &st, uint32_t pos)
{
    const uint32_t i = pos >> 5;
    const uint32_t w0 = lnb_lds32(st.saddr + (i % LNB_TL_RING) * 128u), w1 = lnb_lds32(st.saddr + ((i + 1u) % LNB_TL_RING) * 128u);
    return __funnelshift_l(lnb_bswap32(w1), lnb_bswap32(w0), pos);
}
__device__ __forceinline__ uint32_t lnb_tl_get(const LnbTlLane &st, uint32_t &pos, uint32_t n)   /* 1 <= n <= 32 */
{
    const uint32_t v = lnb_tl_peek(st, pos) >> (32u - n);
    pos += n;
    return v;
}

__global__ void __launch_bounds__(32) lnb_tp_entropy_l_kernel(LnbDecodeBatch b)
{
    __shared__ __align__(16) uint32_t s_ring[LNB_TL_RING * 32u];
    __shared__ int32_t s_tile[32u * 33u];
    const uint32_t lane = threadIdx.x;
    const uint32_t blk_i = blockIdx.x * 32u + lane;
    const LnbStreamCfg &cfg = b.cfg;
    const uint32_t C = cfg.num_channels, n = cfg.block_size;
    LnbBlockDesc blk;
    bool active = blk_i < b.num_blocks;
    if (active) { blk = b.blocks[blk_i]; active = lnb_tp_shape_ok(b, blk); }
    const uint32_t act_mask = __ballot_sync(0xffffffffu, active);
    if (act_mask == 0u) return;

    const uint32_t ring_saddr = lnb_smem_addr(s_ring);
    LnbTlLane st;
    st.gw = (const uint32_t *)b.stream; st.end_word = 0u; st.issued = 0u; st.saddr = ring_saddr + lane * 4u;
    uint32_t pos = 0u, rel_payload = 0u, rel_end = 0u, overrun = 0u, smp_off = 0u;
    if (active) {
        const uint32_t word0 = blk.byte_off >> 2;                /* bit positions relative to this word never overflow */
        uint32_t end_byte = blk.byte_off + blk.byte_size;
        if (end_byte > b.stream_size || end_byte < blk.byte_off) end_byte = b.stream_size;
        st.gw = (const uint32_t *)b.stream + word0;
        rel_payload = blk.byte_off + LNB_BLOCK_HEADER_SIZE - word0 * 4u; rel_end = end_byte - word0 * 4u;
        st.end_word = (rel_end + 3u) >> 2;
        pos = rel_payload * 8u;
        smp_off = blk.smp_off;
    }
    const uint32_t pos_limit = (st.end_word + 4u) * 32u;          /* nothing sane reads past this */
    lnb_tl_check<2>(st, ring_saddr, pos, active, true, lane);        /* the first 128 words of every block */
    lnb_tl_check<2>(st, ring_saddr, pos, active, true, lane);

    /* ---- side information (linne_decoder.c:457-486): each lane parses its own block ---- */
    {
        LnbChanParams *params = b.params + (size_t)(active ? blk_i : 0u) * C;
        for (uint32_t c = 0; c < C; c++) {
            for (int f = 0; f < LNB_NUM_PREEM; f++) {
                const int32_t prev = lnb_zz_dec(lnb_tl_get(st, pos, cfg.bits_per_sample + 1u));
                const uint32_t coef = lnb_tl_get(st, pos, LNB_PREEM_SHIFT - 1);
                if (active) { params[c].preem_prev[f] = prev; params[c].preem_coef[f] = (uint8_t)coef; }
            }
        }
        lnb_tl_check<2>(st, ring_saddr, pos, active, true, lane);
        for (uint32_t c = 0; c < C; c++)
            for (uint32_t l = 0; l < cfg.num_layers; l++) {
                const uint32_t P = cfg.layer_params[l];
                const uint32_t lu = lnb_tl_get(st, pos, 3), rs = lnb_tl_get(st, pos, 4);
                if (active) { params[c].log2_units[l] = (uint8_t)lu; params[c].rshift[l] = (uint8_t)rs; }
                int8_t *q = params[c].coef + l * LNB_MAX_PARAMS;
                for (uint32_t i0 = 0; i0 < P; i0 += 32u) {       /* at most 32 x 15 bits between two checks */
                    const uint32_t lim = (P - i0 < 32u) ? P - i0 : 32u;
                    for (uint32_t i = 0; i < lim; i++) {
                        const uint32_t e = b.tab.huff_lut[lnb_tl_peek(st, pos) >> (32 - LNB_HUFF_LUT_BITS)];
                        pos += e & 15u;
                        if (active) q[i0 + i] = (int8_t)lnb_zz_dec(e >> 4);
                    }
                    lnb_tl_check<2>(st, ring_saddr, pos, active, true, lane);
                }
            }
    }

    /* ---- residuals (linne_coder.c:306-327): one code word per lane and pass ---- */
    for (uint32_t c = 0; c < C; c++) {
        const uint32_t porder = lnb_tl_get(st, pos, 10);
        uint32_t k2 = lnb_tl_get(st, pos, 5);                     /* first partition: k2 itself (linne_coder.c:313) */
        if (porder > LNB_MAX_PORDER || k2 > 30u) { overrun = 1u; k2 = 0u; }
        const uint32_t len = n >> (porder > LNB_MAX_PORDER ? 0u : porder);
        uint32_t left = len;
        const size_t plane = (size_t)c * cfg.pcm_stride;
        for (uint32_t i0 = 0; i0 < n; i0 += 32u) {
#pragma unroll 1
            for (uint32_t s = 0; s < 32u; s++) {
                if ((s % LNB_TL_CHECK) == 0u) lnb_tl_check<2>(st, ring_saddr, pos, active && !overrun, false, lane);
                int32_t val = 0;
                if (active && !overrun) {
                    if (left == 0u) {                            /* partition header: gamma code of zigzag(k2 - previous k2) */
                        const uint32_t h = lnb_tl_peek(st, pos);
                        const uint32_t lz = lnb_clz32(h);
                        const uint32_t z = lz & 15u;
                        const uint32_t gv = ((h << z) >> (31u - z)) - 1u;
                        pos += 2u * lz + 1u;
                        k2 = (uint32_t)((int32_t)k2 + lnb_zz_dec(gv));
                        left = len;
                        if (lz > 15u || k2 > 30u) { overrun = 1u; k2 = 0u; }
                    }
                    const uint32_t v = lnb_tl_peek(st, pos);
                    const uint32_t lz = lnb_clz32(v);
                    if (__builtin_expect(lz + k2 <= 31u, 1)) {
                        const uint32_t ml = (lz > 1u) ? lz : 1u;
                        const uint32_t low = (v >> ((31u - k2 - ml) & 31u)) & ((1u << k2) - 1u);
                        const uint32_t mult = lz ? lz + 1u : ((v >> 30) & 1u);
                        val = lnb_zz_dec((mult << k2) + low);
                        pos += k2 + 1u + ml;
                    } else {                                     /* longer than 32 bits: read from global memory, then re-prime the ring */
                        LnbFastReader fr;
                        lnb_fr_open(fr, st.gw, pos, st.end_word);
                        val = lnb_zz_dec(lnb_get_rice(fr, k2 + 1u, k2));
                        pos = (uint32_t)lnb_fr_position(fr);
                        if (fr.overrun || pos > pos_limit) overrun = 1u;
                    }
                    left--;
                    if (overrun) val = 0;
                }
                const bool jumped = active && !overrun && (pos >> 5) + 2u >= st.issued;
                if (__any_sync(0xffffffffu, jumped)) lnb_tl_check<2>(st, ring_saddr, pos, active && !overrun, true, lane);
                s_tile[s * 33u + lane] = val;
            }
            __syncwarp();
            /* the tile leaves transposed: 32 samples of one block per store instruction */
#pragma unroll 4
            for (uint32_t r = 0; r < 32u; r++) {
                const uint32_t off = __shfl_sync(0xffffffffu, smp_off, (int)r);
                if ((act_mask >> r) & 1u) b.pcm[plane + off + i0 + lane] = s_tile[lane * 33u + r];
            }
            __syncwarp();
        }
    }
    if (active) {
        const uint32_t used = (pos - rel_payload * 8u + 7u) >> 3;
        b.blocks[blk_i].na = used;                               /* payload bytes consumed (reference Flush + Tell) */
        if (overrun || rel_payload + used > rel_end) atomicOr(&b.blocks[blk_i].status, (uint32_t)LNB_ST_OVERRUN);
    }
}

/* ------------------------------------------------------------------------------------------------------
 * entropy, fourth form: ONE LANE per block walking GROUPS of eight code words without a branch.
 *
 * The one-code-word-per-pass kernel above needs few instructions but every pass is one long dependent chain (~100
 * instructions at the ~5 cycles a lone warp gets per dependent instruction), and a batch has only one warp per scheduler.
 * Here a lane keeps a 128-bit window of its payload in registers and walks up to eight code words per step the way the
 * fused decoder's walk does (lnb_stream_v2.cuh: length = (k2 + 32) - leading-one index + top bit, four funnel shifts), with
 * the residual of every slot computed beside the chain.  A slot counts while the window still covers it and no code word
 * before it was longer than 32 bits; what does not count is simply walked again by the next step, so nothing needs a
 * second reader except a code word longer than 32 bits at the head of a step (read from global memory, rare).  Lanes run
 * out of step with each other (partition ends differ), so residuals leave through a per-lane staging ring, 32 at a time,
 * written by the whole warp (128 contiguous bytes per store).
 * ------------------------------------------------------------------------------------------------------ */
#define LNB_TS_OROWS 64u                                       /* staging rows per lane (a lane holds at most 31 + 8) */

__device__ __forceinline__ void lnb_tl_window128(const LnbTlLane &st, uint32_t pos, uint32_t &w0, uint32_t &w1, uint32_t &w2, uint32_t &w3)
{
    const uint32_t i = pos >> 5;
    const uint32_t v0 = lnb_bswap32(lnb_lds32(st.saddr + (i % LNB_TL_RING) * 128u)), v1 = lnb_bswap32(lnb_lds32(st.saddr + ((i + 1u) % LNB_TL_RING) * 128u));
    const uint32_t v2 = lnb_bswap32(lnb_lds32(st.saddr + ((i + 2u) % LNB_TL_RING) * 128u)), v3 = lnb_bswap32(lnb_lds32(st.saddr + ((i + 3u) % LNB_TL_RING) * 128u));
    const uint32_t v4 = lnb_bswap32(lnb_lds32(st.saddr + ((i + 4u) % LNB_TL_RING) * 128u));
    w0 = __funnelshift_l(v1, v0, pos); w1 = __funnelshift_l(v2, v1, pos);
    w2 = __funnelshift_l(v3, v2, pos); w3 = __funnelshift_l(v4, v3, pos);
}

__global__ void __launch_bounds__(32) lnb_tp_entropy_s_kernel(LnbDecodeBatch b)
{
    __shared__ __align__(16) uint32_t s_ring[LNB_TL_RING * 32u];
    __shared__ int32_t s_out[LNB_TS_OROWS * 33u];
    const uint32_t lane = threadIdx.x;
    const uint32_t blk_i = blockIdx.x * 32u + lane;
    const LnbStreamCfg &cfg = b.cfg;
    const uint32_t C = cfg.num_channels, n = cfg.block_size;
    LnbBlockDesc blk;
    bool active = blk_i < b.num_blocks;
    if (active) { blk = b.blocks[blk_i]; active = lnb_tp_shape_ok(b, blk); }
    if (__ballot_sync(0xffffffffu, active) == 0u) return;

    const uint32_t ring_saddr = lnb_smem_addr(s_ring);
    LnbTlLane st;
    st.gw = (const uint32_t *)b.stream; st.end_word = 0u; st.issued = 0u; st.saddr = ring_saddr + lane * 4u;
    uint32_t pos = 0u, rel_payload = 0u, rel_end = 0u, overrun = 0u;
    int32_t *obase = b.pcm;                                       /* plane 0 of this lane's block */
    if (active) {
        const uint32_t word0 = blk.byte_off >> 2;                /* bit positions relative to this word never overflow */
        uint32_t end_byte = blk.byte_off + blk.byte_size;
        if (end_byte > b.stream_size || end_byte < blk.byte_off) end_byte = b.stream_size;
        st.gw = (const uint32_t *)b.stream + word0;
        rel_payload = blk.byte_off + LNB_BLOCK_HEADER_SIZE - word0 * 4u; rel_end = end_byte - word0 * 4u;
        st.end_word = (rel_end + 3u) >> 2;
        pos = rel_payload * 8u;
        obase = b.pcm + blk.smp_off;
    }
    const uint32_t pos_limit = (st.end_word + 4u) * 32u;          /* nothing sane reads past this */
    lnb_tl_check<2>(st, ring_saddr, pos, active, true, lane);        /* the first 128 words of every block */

    /* ---- side information (linne_decoder.c:457-486): each lane parses its own block ---- */
    {
        LnbChanParams *params = b.params + (size_t)(active ? blk_i : 0u) * C;
        for (uint32_t c = 0; c < C; c++) {
            for (int f = 0; f < LNB_NUM_PREEM; f++) {
                const int32_t prev = lnb_zz_dec(lnb_tl_get(st, pos, cfg.bits_per_sample + 1u));
                const uint32_t coef = lnb_tl_get(st, pos, LNB_PREEM_SHIFT - 1);
                if (active) { params[c].preem_prev[f] = prev; params[c].preem_coef[f] = (uint8_t)coef; }
            }
        }
        lnb_tl_check<2>(st, ring_saddr, pos, active, true, lane);
        for (uint32_t c = 0; c < C; c++)
            for (uint32_t l = 0; l < cfg.num_layers; l++) {
                const uint32_t P = cfg.layer_params[l];
                const uint32_t lu = lnb_tl_get(st, pos, 3), rs = lnb_tl_get(st, pos, 4);
                if (active) { params[c].log2_units[l] = (uint8_t)lu; params[c].rshift[l] = (uint8_t)rs; }
                int8_t *q = params[c].coef + l * LNB_MAX_PARAMS;
                for (uint32_t i0 = 0; i0 < P; i0 += 32u) {       /* at most 32 x 15 bits between two checks */
                    const uint32_t lim = (P - i0 < 32u) ? P - i0 : 32u;
                    for (uint32_t i = 0; i < lim; i++) {
                        const uint32_t e = b.tab.huff_lut[lnb_tl_peek(st, pos) >> (32 - LNB_HUFF_LUT_BITS)];
                        pos += e & 15u;
                        if (active) q[i0 + i] = (int8_t)lnb_zz_dec(e >> 4);
                    }
                    lnb_tl_check<2>(st, ring_saddr, pos, active, true, lane);
                }
            }
    }

    /* ---- residuals (linne_coder.c:306-327) ---- */
    const uint32_t out_saddr = lnb_smem_addr(s_out) + lane * 4u;  /* row r of this lane: out_saddr + r * 33 * 4 */
    uint32_t chan = 0u, left = 0u, parts_left = 0u, len = 0u, k2 = 0u;
    uint32_t produced = 0u;                                      /* residuals of the current channel walked so far */
    uint32_t wr = 0u, flushed = 0u;                              /* staging ring: rows written / rows sent, over the whole block */
    bool running = active;
    for (uint32_t step = 0;; step++) {
        if (!__any_sync(0xffffffffu, running)) break;
        if ((step & 3u) == 0u) lnb_tl_check<2>(st, ring_saddr, pos, running, false, lane);
        /* channel and partition headers of the lanes that stand in front of one */
        if (running && left == 0u) {
            if (parts_left == 0u) {
                if (chan == C) {
                    running = false;
                } else {
                    const uint32_t porder = lnb_tl_get(st, pos, 10);
                    k2 = lnb_tl_get(st, pos, 5);                  /* first partition: k2 itself (linne_coder.c:313) */
                    if (porder > LNB_MAX_PORDER || k2 > 30u) { overrun = 1u; running = false; }
                    else { len = n >> porder; parts_left = (1u << porder) - 1u; left = len; produced = 0u; chan++; }
                }
            } else {                                             /* gamma code of zigzag(k2 - previous k2) */
                const uint32_t h = lnb_tl_peek(st, pos);
                const uint32_t lz = lnb_clz32(h);
                const uint32_t z = lz & 15u;
                const uint32_t gv = ((h << z) >> (31u - z)) - 1u;
                pos += 2u * lz + 1u;
                k2 = (uint32_t)((int32_t)k2 + lnb_zz_dec(gv));
                parts_left--; left = len;
                if (lz > 15u || k2 > 30u) { overrun = 1u; running = false; }
            }
        }
        /* the group: up to eight code words from a 128-bit window, no branch */
        uint32_t w0, w1, w2, w3;
        lnb_tl_window128(st, pos, w0, w1, w2, w3);
        const uint32_t gmax = (k2 <= 8u) ? 8u : ((k2 <= 19u) ? 4u : 2u);
        const uint32_t G = running ? (left < gmax ? left : gmax) : 0u;
        const uint32_t K = k2 + 32u, sh = 31u - k2, k2mask = (1u << k2) - 1u;
        uint32_t T = 0u, taken = 0u, adv = 0u, good = 1u;
#pragma unroll
        for (uint32_t s = 0; s < 8u; s++) {
            const uint32_t f = lnb_bfind(w0);                    /* 31 - leading zeros; 0xFFFFFFFF for an all-zero window */
            const uint32_t lz = 31u - f, ml = (lz > 1u) ? lz : 1u;
            const uint32_t low = (w0 >> ((sh - ml) & 31u)) & k2mask;
            const uint32_t mult = lz ? lz + 1u : ((w0 >> 30) & 1u);
            const int32_t val = lnb_zz_dec((mult << k2) + low);
            /* the slot counts: asked for, window still covers its 32 bits, inside them, and so did every slot before */
            good &= (s < G && T <= 96u && (int32_t)f >= (int32_t)k2) ? 1u : 0u;
            const uint32_t L = K - f + (w0 >> 31);               /* k2 + 1 + max(lz, 1) */
            lnb_sts32(out_saddr + ((wr + s) % LNB_TS_OROWS) * 132u, (uint32_t)val);
            w0 = __funnelshift_lc(w1, w0, L); w1 = __funnelshift_lc(w2, w1, L);
            w2 = __funnelshift_lc(w3, w2, L); w3 = __funnelshift_lc(0u, w3, L);
            T += L;
            taken += good;
            adv = good ? T : adv;
        }
        if (running) {
            if (__builtin_expect(taken == 0u, 0)) {              /* a code word longer than 32 bits heads the step: global memory */
                LnbFastReader fr;
                lnb_fr_open(fr, st.gw, pos, st.end_word);
                const int32_t val = lnb_zz_dec(lnb_get_rice(fr, k2 + 1u, k2));
                lnb_sts32(out_saddr + (wr % LNB_TS_OROWS) * 132u, (uint32_t)val);
                pos = (uint32_t)lnb_fr_position(fr);
                taken = 1u;
                if (fr.overrun || pos > pos_limit) { overrun = 1u; running = false; }
            } else {
                pos += adv;
            }
            wr += taken; left -= taken; produced += taken;
        }
        /* a lane that left its ring behind (long code word) gets it re-primed now */
        if (__any_sync(0xffffffffu, running && (pos >> 5) + 6u >= st.issued)) lnb_tl_check<2>(st, ring_saddr, pos, running, true, lane);
        /* residuals leave 32 at a time, one lane's rows per store instruction */
        for (;;) {
            const uint32_t m = __ballot_sync(0xffffffffu, active && wr - flushed >= 32u);
            if (m == 0u) break;
            const int l = __ffs((int)m) - 1;
            const unsigned long long op = __shfl_sync(0xffffffffu, (unsigned long long)(uintptr_t)obase, l);
            const uint32_t fl = __shfl_sync(0xffffffffu, flushed, l);
            /* row fl + j of lane l is sample (fl + j) of the block's concatenated channels: channel (fl / n), index (fl % n) */
            const uint32_t ch = fl / n, idx = fl - ch * n;
            const int32_t v = (int32_t)lnb_lds32(lnb_smem_addr(s_out) + (((fl + lane) % LNB_TS_OROWS) * 33u + (uint32_t)l) * 4u);
            ((int32_t *)(uintptr_t)op)[(size_t)ch * cfg.pcm_stride + idx + lane] = v;
            if ((int)lane == l) flushed += 32u;
        }
    }
    if (active) {
        const uint32_t used = (pos - rel_payload * 8u + 7u) >> 3;
        b.blocks[blk_i].na = used;                               /* payload bytes consumed (reference Flush + Tell) */
        if (overrun || rel_payload + used > rel_end) atomicOr(&b.blocks[blk_i].status, (uint32_t)LNB_ST_OVERRUN);
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
