/* lnb_entropy_v3.cuh -- warp-parallel entropy decode: one warp per block, 32 code words per round.
 *
 * Covers reference rows d2, d3, d6: libs/linne_decoder/src/linne_decoder.c:457-497
 * (side information + residuals), :387-421 (raw), :554-557 (silent), libs/linne_coder/src/linne_coder.c
 * :306-327 (partitioned recursive Rice), :106-127 (gamma), :150-169 (Rice symbol),
 * libs/static_huffman/src/static_huffman.c:145-165 (coefficient symbols).
 *
 * The format gives a block one entry point: channels are concatenated and the Rice parameter of each
 * partition is delta-coded inline, so where a code word starts is known only once the previous one has
 * been measured.  What breaks the chain is a property of the code itself.  With k1 = k2 + 1 a recursive
 * Rice code word is  lz zeros, a one, then k1 bits (lz = 0) or k2 bits (lz >= 1): its length is
 * k2 + 1 + max(lz, 1).  Both lz = 0 and lz = 1 -- all residuals below 3 * 2^k2, the large majority when
 * k2 tracks the partition mean -- give the SAME length k2 + 2.  So each round
 *   - lane j guesses that code word j starts at  pos + j * (k2 + 2),  reads its own window from the
 *     shared-memory copy of the payload and counts the leading zeros;
 *   - a ballot finds the first lane with lz >= 2 (a longer code word): every lane up to and including it
 *     guessed right, extracts its value and stores it (consecutive samples: one coalesced store);
 *   - the position after that lane's code word starts the next round.
 * A round costs one shared-memory read, one clz, one ballot and one shuffle on the dependency chain and
 * retires 1..32 code words.  Code words longer than 32 bits (rare) are finished by their lane on the
 * general serial reader.  The partition's gamma-coded parameter delta and the side information are read
 * redundantly by all lanes (uniform control flow, no broadcast needed).
 *
 * The payload is staged through a per-warp shared-memory window (LNB_E3_WIN big-endian-corrected words,
 * refilled with coalesced loads when the read position nears its end), so HBM is read once, coalesced.
 */
#pragma once
#include "lnb_common.cuh"
#include "lnb_decode_core.cuh"

#define LNB_E3_WARPS   4
#define LNB_E3_THREADS (32 * LNB_E3_WARPS)
#define LNB_E3_WIN     512u                 /* words per warp window (2 KB) */
#define LNB_E3_ROUND_BITS (32u * 33u + 64u) /* furthest bit a round can look at, relative to its start */
#define LNB_E3_HUFF1_BITS 9                 /* first-level coefficient LUT in shared memory */

struct LnbE3Win {
    uint32_t *buf;              /* [LNB_E3_WIN + 2] */
    const uint32_t *words;      /* global, word 0 = aligned word holding the block's first byte */
    uint32_t wb;                /* index of the word held in buf[0] */
    uint32_t end_word;          /* words at or past this index read as zero */
    uint32_t limit;             /* a residual round may start at bit positions up to this one without a refill */
};

__device__ __forceinline__ void lnb_e3_fill(LnbE3Win &w, uint32_t word_idx, uint32_t lane)
{
    __syncwarp();
    w.wb = word_idx;
    w.limit = (word_idx + LNB_E3_WIN) * 32u - LNB_E3_ROUND_BITS;
#pragma unroll 8
    for (uint32_t i = lane; i < LNB_E3_WIN + 2u; i += 32u) {
        const uint32_t idx = word_idx + i;
        w.buf[i] = (idx < w.end_word) ? lnb_bswap32(w.words[idx]) : 0u;
    }
    __syncwarp();
}
/* make bits [pos, pos + span) (+ the 32-bit peek beyond) readable; warp-uniform */
__device__ __forceinline__ void lnb_e3_ensure(LnbE3Win &w, uint32_t pos, uint32_t span, uint32_t lane)
{
    if (((pos + span) >> 5) + 2u > w.wb + LNB_E3_WIN + 2u || (pos >> 5) < w.wb) lnb_e3_fill(w, pos >> 5, lane);
}
/* the 32 bits starting at bit `pos` (MSB first) */
__device__ __forceinline__ uint32_t lnb_e3_peek(const LnbE3Win &w, uint32_t pos)
{
    const uint32_t i = (pos >> 5) - w.wb;
    return __funnelshift_l(w.buf[i + 1u], w.buf[i], pos & 31u);
}
__device__ __forceinline__ uint32_t lnb_e3_get(const LnbE3Win &w, uint32_t &pos, uint32_t n)   /* 1 <= n <= 32 */
{
    const uint32_t v = lnb_e3_peek(w, pos) >> (32u - n);
    pos += n;
    return v;
}

/* Where the residuals of a compressed block go.  The stand-alone kernel writes them to the PCM planes; the
 * fused decoder (lnb_stream_v2.cuh) writes them to shared memory and publishes its progress to the synthesis
 * warps.  `params` receives the side information of the block's channels. */
struct LnbE3PlaneSink {
    static constexpr bool kPublish = false;
    LnbChanParams *params;
    int32_t *pcm; uint32_t stride, smp_off;
    __device__ __forceinline__ int32_t *channel(uint32_t c) const { return pcm + (size_t)c * stride + smp_off; }
    __device__ __forceinline__ void begin_channel(uint32_t, uint32_t, uint32_t) const {}
    __device__ __forceinline__ void publish(uint32_t, uint32_t, uint32_t, uint32_t) const {}
    __device__ __forceinline__ void abort(uint32_t) const {}
};

/* payload of one COMPRESSED block, decoded by one warp */
template <class Sink>
__device__ void lnb_e3_compressed_block(const LnbDecodeBatch &b, LnbBlockDesc &gblk, const LnbBlockDesc &blk, LnbE3Win &win,
                                        const uint16_t *huff1, Sink &sink, uint32_t lane)
{
    const LnbStreamCfg &cfg = b.cfg;
    const uint32_t C = cfg.num_channels, n = blk.nsmp;
    const uint32_t payload_off = blk.byte_off + LNB_BLOCK_HEADER_SIZE;
    uint32_t end_byte = blk.byte_off + blk.byte_size;
    if (end_byte > b.stream_size) end_byte = b.stream_size;
    /* bit positions are relative to the aligned word that holds the block's first byte, so 32-bit
     * positions never overflow however large the stream is */
    const uint32_t word0 = blk.byte_off >> 2;
        win.words = (const uint32_t *)b.stream + word0;
    const uint32_t rel_payload = payload_off - word0 * 4u, rel_end = end_byte - word0 * 4u;
    win.end_word = (rel_end + 3u) >> 2;
    uint32_t pos = rel_payload * 8u;
    uint32_t overrun = 0;
    lnb_e3_fill(win, pos >> 5, lane);

    /* ---- side information (linne_decoder.c:457-486): every lane reads the same fields ---- */
    {
        LnbChanParams *params = sink.params;
        for (uint32_t c = 0; c < C; c++)
            for (int f = 0; f < LNB_NUM_PREEM; f++) {
                lnb_e3_ensure(win, pos, 64u, lane);
                const int32_t prev = lnb_zz_dec(lnb_e3_get(win, pos, cfg.bits_per_sample + 1u));
                const uint32_t coef = lnb_e3_get(win, pos, LNB_PREEM_SHIFT - 1);
                if (lane == 0) { params[c].preem_prev[f] = prev; params[c].preem_coef[f] = (uint8_t)coef; }
            }
        for (uint32_t c = 0; c < C; c++)
            for (uint32_t l = 0; l < cfg.num_layers; l++) {
                const uint32_t P = cfg.layer_params[l];
                lnb_e3_ensure(win, pos, 7u + P * 14u + 32u, lane);
                const uint32_t lu = lnb_e3_get(win, pos, 3), rs = lnb_e3_get(win, pos, 4);
                if (lane == 0) { params[c].log2_units[l] = (uint8_t)lu; params[c].rshift[l] = (uint8_t)rs; }
                int8_t *q = params[c].coef + l * LNB_MAX_PARAMS;
                /* lane j keeps coefficients j, j+32, ...; one 4-byte-aligned store pattern at the end */
                for (uint32_t i0 = 0; i0 < P; i0 += 32u) {
                    int32_t mine = 0;
                    const uint32_t lim = (P - i0 < 32u) ? P - i0 : 32u;
                    for (uint32_t i = 0; i < lim; i++) {
                        const uint32_t top = lnb_e3_peek(win, pos);
                        uint32_t e = huff1[top >> (32 - LNB_E3_HUFF1_BITS)];
                        if (e == 0u) e = b.tab.huff_lut[top >> (32 - LNB_HUFF_LUT_BITS)];
                        pos += e & 15u;
                        if (i == lane) mine = lnb_zz_dec(e >> 4);
                    }
                    if (lane < lim) q[i0 + lane] = (int8_t)mine;
                }
            }
    }

    /* ---- residuals, channel after channel (linne_coder.c:306-327) ---- */
    for (uint32_t c = 0; c < C; c++) {
        if (overrun) {                                           /* broken stream: the remaining channels read as silence */
            int32_t *gout = b.pcm + (size_t)c * cfg.pcm_stride + blk.smp_off;
            for (uint32_t i = lane; i < n; i += 32u) gout[i] = 0;
            continue;
        }
        sink.begin_channel(c, n, lane);
        int32_t *out = sink.channel(c);
        uint32_t published = 0;
        lnb_e3_ensure(win, pos, 64u, lane);
        uint32_t porder = lnb_e3_get(win, pos, 10);
        if (porder > LNB_MAX_PORDER) { overrun = 1; porder = 0; }
        const uint32_t len = n >> porder, parts = 1u << porder;
        uint32_t k2 = 0, done = 0;
        for (uint32_t part = 0; part < parts && !overrun; part++) {
            lnb_e3_ensure(win, pos, 96u, lane);
            if (part == 0) {
                k2 = lnb_e3_get(win, pos, 5);
            } else {                                             /* gamma code of zigzag(k2 - previous k2) */
                const uint32_t h = lnb_e3_peek(win, pos);
                const uint32_t lz = lnb_clz32(h);
                if (lz > 15u) { overrun = 1; break; }
                const uint32_t v = ((h << lz) >> (31u - lz)) - 1u;
                pos += 2u * lz + 1u;
                k2 = (uint32_t)((int32_t)k2 + lnb_zz_dec(v));
            }
            if (k2 > 30u) { overrun = 1; k2 = 30u; }
            const uint32_t step = k2 + 2u;
            const uint32_t my_rel = lane * step;                  /* guessed start of this lane's code word, relative to pos */
            const uint32_t my_end = my_rel + k2 + 1u;             /* + max(lz, 1) = end of the code word */
            uint32_t rem = len;
            const uint32_t k2mask = (1u << k2) - 1u;
            while (rem) {
                if (pos > win.limit) lnb_e3_fill(win, pos >> 5, lane);
                const uint32_t cnt = rem < 32u ? rem : 32u;
                const uint32_t hi = lnb_e3_peek(win, pos + my_rel);
                const uint32_t lz = lnb_clz32(hi);
                const uint32_t ml = (lz > 1u) ? lz : 1u;
                /* The round is resolved by the lowest lane that either holds a longer code word (lz >= 2) or
                 * is the last lane of the run: one min-reduction over  lane | short flag | end position. */
                const bool resolves = (hi < 0x40000000u || lane + 1u == cnt) && lane < cnt;
                const uint32_t is_short = (lz + k2 <= 31u) ? 0x8000u : 0u;   /* whole code word inside the 32-bit peek */
                const uint32_t key = resolves ? ((lane << 16) | is_short | (my_end + ml)) : 0xFFFFFFFFu;
                const uint32_t r = __reduce_min_sync(0xffffffffu, key);
                const uint32_t first = r >> 16;
                uint32_t n_ok = first + ((r >> 15) & 1u);
                {
                    /* the k2 bits after the unary part sit (ml + 1) bits below the top of the peek */
                    const uint32_t low = (hi >> ((31u - k2) - ml)) & k2mask;
                    const uint32_t mult = lz ? lz + 1u : ((hi >> 30) & 1u);
                    const int32_t v = lnb_zz_dec((mult << k2) + low);
                    if (lane < n_ok) out[done + lane] = v;
                }
                if (__builtin_expect((r & 0x8000u) != 0u, 1)) {
                    pos += r & 0x7FFFu;
                } else {
                    /* code word longer than 32 bits: its lane finishes it on the serial reader */
                    uint32_t endl = 0, bad = 0;
                    if (lane == first) {
                        LnbFastReader fr;
                        lnb_fr_open(fr, win.words, pos + my_rel, win.end_word);
                        const uint32_t q = lnb_fr_zero_run(fr);
                        const uint32_t u = (q == 0u) ? lnb_fr_get(fr, k2 + 1u)
                                                     : lnb_fr_get(fr, k2) + (2u << k2) + ((q - 1u) << k2);
                        out[done + lane] = lnb_zz_dec(u);
                        endl = (uint32_t)lnb_fr_position(fr);
                        bad = fr.overrun;
                    }
                    pos = __shfl_sync(0xffffffffu, endl, (int)first);
                    n_ok = first + 1u;
                    if (__shfl_sync(0xffffffffu, bad, (int)first)) { overrun = 1; done += n_ok; break; }
                }
                done += n_ok;
                rem -= n_ok;
                if (Sink::kPublish && done - published >= 128u) { sink.publish(c, n, done, lane); published = done; }
            }
            if (Sink::kPublish) { sink.publish(c, n, done, lane); published = done; }
        }
        /* samples a broken stream leaves uncovered */
        if (overrun || parts * len < n)
            for (uint32_t i = done + lane; i < n; i += 32u) if (overrun || i >= parts * len) out[i] = 0;
        if (overrun) sink.abort(lane);
        else if (Sink::kPublish) sink.publish(c, n, n, lane);
    }
    if (overrun) sink.abort(lane);
    if (lane == 0) {
        const uint32_t used = (pos - rel_payload * 8u + 7u) >> 3;
        gblk.na = used;                                          /* payload bytes consumed (reference Flush + Tell) */
        if (overrun || rel_payload + used > rel_end) gblk.status = blk.status | LNB_ST_OVERRUN;
    }
}

__global__ void __launch_bounds__(LNB_E3_THREADS) lnb_entropy_v3_kernel(LnbDecodeBatch b)
{
    __shared__ uint32_t s_win[LNB_E3_WARPS][LNB_E3_WIN + 2];
    __shared__ uint16_t s_huff1[1u << LNB_E3_HUFF1_BITS];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const uint32_t blk_i = blockIdx.x * LNB_E3_WARPS + warp;

    /* first-level coefficient table: code words of at most LNB_E3_HUFF1_BITS bits resolve in shared memory */
    for (uint32_t i = threadIdx.x; i < (1u << LNB_E3_HUFF1_BITS); i += LNB_E3_THREADS) {
        const uint16_t e = b.tab.huff_lut[i << (LNB_HUFF_LUT_BITS - LNB_E3_HUFF1_BITS)];
        s_huff1[i] = ((e & 15u) <= LNB_E3_HUFF1_BITS) ? e : (uint16_t)0;
    }
    __syncthreads();
    if (blk_i >= b.num_blocks) return;

    LnbBlockDesc &gblk = b.blocks[blk_i];
    const LnbBlockDesc blk = gblk;
    const LnbStreamCfg &cfg = b.cfg;
    const uint32_t C = cfg.num_channels, n = blk.nsmp;
    const uint32_t payload_off = blk.byte_off + LNB_BLOCK_HEADER_SIZE;
    uint32_t end_byte = blk.byte_off + blk.byte_size;
    if (end_byte > b.stream_size) end_byte = b.stream_size;

    if (blk.type == LNB_BLOCK_SILENT) {                          /* linne_decoder.c:554-557 */
        for (uint32_t c = 0; c < C; c++) {
            int32_t *dst = b.pcm + (size_t)c * cfg.pcm_stride + blk.smp_off;
            for (uint32_t i = lane; i < n; i += 32u) dst[i] = 0;
        }
        if (lane == 0) gblk.na = 0;
        return;
    }
    if (blk.type == LNB_BLOCK_RAW) {                             /* linne_decoder.c:387-421 */
        const uint32_t bytes = cfg.bits_per_sample >> 3;
        if ((uint64_t)payload_off + (uint64_t)bytes * n * C > end_byte) { if (lane == 0) gblk.status = blk.status | LNB_ST_OVERRUN; return; }
        const uint8_t *p = b.stream + payload_off;
        for (uint32_t c = 0; c < C; c++) {
            int32_t *dst = b.pcm + (size_t)c * cfg.pcm_stride + blk.smp_off;
            for (uint32_t i = lane; i < n; i += 32u)
                dst[i] = lnb_zz_dec(lnb_get_be(p + ((size_t)i * C + c) * bytes, (int)bytes));
        }
        if (lane == 0) gblk.na = bytes * n * C;
        return;
    }
    if (blk.type != LNB_BLOCK_COMPRESSED) { if (lane == 0) gblk.status = blk.status | LNB_ST_BAD_TYPE; return; }
    if (b.fused_max_n && n <= b.fused_max_n && n > 0u) return;   /* taken (or, after a failed CRC, dropped) by the fused kernel */

    LnbE3Win win;
    win.buf = s_win[warp];
    LnbE3PlaneSink sink;
    sink.params = b.params + (size_t)blk_i * C;
    sink.pcm = b.pcm; sink.stride = cfg.pcm_stride; sink.smp_off = blk.smp_off;
    lnb_e3_compressed_block(b, gblk, blk, win, s_huff1, sink, lane);
}
