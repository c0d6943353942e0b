/* lnb_entropy_v2.cuh -- cooperative entropy decode: one warp per block.
 *
 * Replaces the one-thread-per-block payload decode (D1).  Covers reference rows d2, d3, d6 (SURVEY
 * section 8a): libs/linne_decoder/src/linne_decoder.c:457-497 (side information + residuals),
 * :387-421 (raw), :554-557 (silent), libs/linne_coder/src/linne_coder.c:306-327 (partitioned
 * recursive Rice), :106-127 (gamma), :150-169 (Rice symbol).
 *
 * The format leaves one serial dependency per block: where a code word starts is known only after the
 * previous one has been measured (channels are concatenated, Rice parameters are delta-coded inline).
 * The kernel keeps exactly that on the serial chain and nothing else:
 *   lane 0   walks the bitstream and only MEASURES code words: clz of a 64-bit register window gives
 *            the unary part, the length is k2 + 1 + max(lz, 1) (because k1 = k2 + 1), one funnel shift
 *            advances the window.  It records the start bit of up to 32 code words.
 *   lanes    then extract the 32 values in parallel from their start bits (own window, own clz), undo
 *            the zig-zag and store 32 consecutive samples with one coalesced store.
 * Code words longer than 32 bits (rare) are decoded completely by lane 0.  Raw and silent blocks are
 * plain data-parallel copies.
 */
#pragma once
#include "lnb_common.cuh"
#include "lnb_decode_core.cuh"

#define LNB_EN_WARPS   4
#define LNB_EN_THREADS (32 * LNB_EN_WARPS)
#define LNB_EN_DIRECT  0xFFFFFFFFu        /* pos[] marker: value already decoded by lane 0 */

__global__ void __launch_bounds__(LNB_EN_THREADS) lnb_entropy_v2_kernel(LnbDecodeBatch b)
{
    __shared__ uint32_t s_pos[LNB_EN_WARPS][32];
    __shared__ uint32_t s_val[LNB_EN_WARPS][32];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const uint32_t blk_i = blockIdx.x * LNB_EN_WARPS + warp;
    if (blk_i >= b.num_blocks) return;
    LnbBlockDesc &gblk = b.blocks[blk_i];
    const LnbBlockDesc blk = gblk;
    const LnbStreamCfg &cfg = b.cfg;
    const uint32_t C = cfg.num_channels, n = blk.nsmp;
    const uint32_t payload_off = blk.byte_off + LNB_BLOCK_HEADER_SIZE;
    uint32_t end_byte = blk.byte_off + blk.byte_size;
    if (end_byte > b.stream_size) end_byte = b.stream_size;
    uint32_t *pos = s_pos[warp], *val = s_val[warp];

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

    /* bit positions are kept relative to the aligned word that holds the block's first byte, so that
     * 32-bit positions never overflow however large the stream is */
    const uint32_t word0 = blk.byte_off >> 2;
    const uint32_t *words = (const uint32_t *)b.stream + word0;
    const uint32_t rel_payload = payload_off - word0 * 4u, rel_end = end_byte - word0 * 4u;
    const uint32_t end_word = (rel_end + 3u) >> 2;
    LnbFastReader fr;                                            /* meaningful on lane 0 only */
    uint32_t overrun = 0;

    /* ---- side information: a few hundred fields, serial on lane 0 (linne_decoder.c:457-486) ---- */
    if (lane == 0) {
        LnbChanParams *params = b.params + (size_t)blk_i * C;
        LnbBitReader br;
        lnb_br_open(br, words, rel_payload, rel_end);
        for (uint32_t c = 0; c < C; c++)
            for (int f = 0; f < LNB_NUM_PREEM; f++) {
                params[c].preem_prev[f] = lnb_zz_dec(lnb_br_get(br, cfg.bits_per_sample + 1u));
                params[c].preem_coef[f] = (uint8_t)lnb_br_get(br, LNB_PREEM_SHIFT - 1);
            }
        for (uint32_t c = 0; c < C; c++)
            for (uint32_t l = 0; l < cfg.num_layers; l++) {
                params[c].log2_units[l] = (uint8_t)lnb_br_get(br, 3);
                params[c].rshift[l] = (uint8_t)lnb_br_get(br, 4);
                int8_t *q = params[c].coef + l * LNB_MAX_PARAMS;
                for (uint32_t i = 0; i < cfg.layer_params[l]; i++) {
                    const uint32_t e = b.tab.huff_lut[lnb_br_peek(br, LNB_HUFF_LUT_BITS)];
                    lnb_br_skip(br, e & 15u);
                    q[i] = (int8_t)lnb_zz_dec(e >> 4);
                }
            }
        overrun |= br.overrun;
        lnb_fr_open(fr, words, (uint64_t)br.next_word * 32u - br.nbits, end_word);
    }

    /* ---- residuals, channel after channel ---- */
    for (uint32_t c = 0; c < C; c++) {
        int32_t *out = b.pcm + (size_t)c * cfg.pcm_stride + blk.smp_off;
        uint32_t porder = 0;
        if (lane == 0) {
            porder = lnb_fr_get(fr, 10);
            if (porder > LNB_MAX_PORDER) { overrun = 1; porder = 0; }
        }
        porder = __shfl_sync(0xffffffffu, porder, 0);
        const uint32_t len = n >> porder, parts = 1u << porder;
        uint32_t k2 = 0, done = 0;
        for (uint32_t part = 0; part < parts; part++) {
            if (lane == 0) {
                if (part == 0) k2 = lnb_fr_get(fr, 5);
                else k2 = (uint32_t)((int32_t)k2 + lnb_zz_dec(lnb_get_gamma(fr)));
                if (k2 > 30u) { overrun = 1; k2 = 30u; }
            }
            k2 = __shfl_sync(0xffffffffu, k2, 0);
            const uint32_t k1 = k2 + 1u;
            for (uint32_t s0 = 0; s0 < len; s0 += 32u) {
                const uint32_t cnt = (len - s0 < 32u) ? len - s0 : 32u;
                if (lane == 0) {
                    /* measure: start bit of each of the next `cnt` code words.  Straight-line, predicated
                     * body (taken branches cost a lone warp ~15 cycles each); a code longer than 32 bits only
                     * raises a flag, and the run is redone from that symbol on the general path. */
                    uint32_t i = 0;
                    while (i < cnt) {
                        uint32_t hi = fr.hi, lo = fr.lo, nbits = fr.nbits, nxt = fr.next, pre = fr.pre;
                        uint32_t position = nxt * 32u - nbits;
                        uint32_t first_long = cnt;
                        const uint32_t kc = k2 + 1u;
#pragma unroll 4
                        for (uint32_t j = i; j < cnt; j++) {
                            const uint32_t lz = lnb_clz32(hi);
                            const uint32_t ml = (lz > 1u) ? lz : 1u;
                            const uint32_t L = kc + ml;              /* k1 = k2 + 1: both code forms have this length */
                            first_long = (lz + k2 > 31u && j < first_long) ? j : first_long;
                            pos[j] = position;
                            position += L;
                            hi = lnb_shl64_hi(hi, lo, L);
                            lo = lnb_shl32_clamped(lo, L);
                            nbits -= L;
                            const bool need = nbits < 32u;           /* refill from the prefetched word */
                            hi |= need ? lnb_shr32_clamped(pre, nbits) : 0u;
                            lo = need ? lnb_shl32_clamped(pre, 32u - nbits) : lo;
                            nbits += need ? 32u : 0u;
                            nxt += need ? 1u : 0u;
                            if (need) pre = (nxt < end_word) ? lnb_bswap32(words[nxt]) : 0u;
                        }
                        if (first_long >= cnt) {                     /* common case: the whole run was short codes */
                            fr.hi = hi; fr.lo = lo; fr.nbits = nbits; fr.next = nxt; fr.pre = pre;
                            if (nxt > end_word + 2u) fr.overrun = 1;
                            break;
                        }
                        /* re-synchronise at the long code word, decode it on the general path, go on after it */
                        lnb_fr_open(fr, words, pos[first_long], end_word);
                        {
                            const uint32_t q = lnb_fr_zero_run(fr);
                            val[first_long] = (q == 0u) ? lnb_fr_get(fr, k1) : lnb_fr_get(fr, k2) + (1u << k1) + ((q - 1u) << k2);
                            pos[first_long] = LNB_EN_DIRECT;
                        }
                        if (fr.overrun) {                            /* corrupt data: fill the rest of the run and stop measuring */
                            for (uint32_t j = first_long + 1u; j < cnt; j++) { val[j] = 0; pos[j] = LNB_EN_DIRECT; }
                            break;
                        }
                        i = first_long + 1u;
                    }
                }
                __syncwarp();
                if (lane < cnt) {
                    /* extract: every lane decodes its own code word from its start bit */
                    uint32_t u;
                    const uint32_t q = pos[lane];
                    if (q == LNB_EN_DIRECT) {
                        u = val[lane];
                    } else {
                        const uint32_t w = q >> 5, sh = q & 31u;
                        const uint32_t w0 = (w < end_word) ? lnb_bswap32(words[w]) : 0u;
                        const uint32_t w1 = (w + 1u < end_word) ? lnb_bswap32(words[w + 1u]) : 0u;
                        const uint32_t hi = __funnelshift_l(w1, w0, sh);
                        const uint32_t lz = lnb_clz32(hi);
                        const uint32_t ml = (lz > 1u) ? lz : 1u;
                        const uint32_t t = (hi << ml) << 1;                       /* bits after the unary part (k2 of them) */
                        const uint32_t low = (t >> 1) >> (31u - k2);
                        const uint32_t mult = lz ? lz + 1u : ((hi >> 30) & 1u);   /* lz = 0: the bit after the leading one */
                        u = (mult << k2) + low;
                    }
                    out[done + s0 + lane] = lnb_zz_dec(u);
                }
                __syncwarp();
            }
            done += len;
        }
    }
    if (lane == 0) {
        overrun |= fr.overrun;
        const uint32_t used = (uint32_t)((lnb_fr_position(fr) - (uint64_t)rel_payload * 8u + 7u) >> 3);
        gblk.na = used;                                          /* payload bytes consumed (reference Flush + Tell) */
        if (overrun || rel_payload + used > rel_end) gblk.status = blk.status | LNB_ST_OVERRUN;
    }
}
