/* lnb_decode_core.cuh -- per-work-item bodies of the decode kernels.
 *
 * Work items:
 *   entropy decode   one per BLOCK        (serial by construction: channels are concatenated and the
 *                                          Rice parameters are delta-coded inline, reference
 *                                          libs/linne_coder/src/linne_coder.c:306-327)
 *   synthesis        one per (block, channel, unit) and layer
 *   de-emphasis      one per (block, channel)
 *   M/S -> L/R       one per sample pair
 */
#pragma once
#include "lnb_common.cuh"

/* ------------------------------------------------------------------------------------------
 * MSB-first bit reader over the device copy of the stream (32-bit aligned words, big-endian bit
 * order inside the byte stream).  Semantics of reference libs/bit_stream/include/bit_stream.h:305-394;
 * here with a 64-bit left-aligned window, clz-based zero-run scan and an explicit end so corrupt
 * data cannot run away.
 * ------------------------------------------------------------------------------------------ */
struct LnbBitReader {
    const uint32_t *words;
    uint32_t next_word, end_word;
    uint64_t acc;        /* valid bits left-aligned, everything below is zero */
    uint32_t nbits;      /* number of valid bits in acc */
    uint32_t start_bit;  /* absolute bit position (in the stream buffer) where reading began */
    uint32_t overrun;
};

LNB_HD void lnb_br_refill(LnbBitReader &br)
{
    while (br.nbits <= 32u) {
        uint32_t w = 0;
        if (br.next_word < br.end_word) w = lnb_bswap32(br.words[br.next_word]);   /* zeros past the end */
        br.next_word++;
        br.acc |= (uint64_t)w << (32u - br.nbits);
        br.nbits += 32u;
    }
}

LNB_HD void lnb_br_open(LnbBitReader &br, const uint32_t *words, uint32_t byte_off, uint32_t end_byte)
{
    br.words = words;
    br.next_word = byte_off >> 2;
    br.end_word = (end_byte + 3u) >> 2;
    br.acc = 0; br.nbits = 0; br.overrun = 0;
    br.start_bit = byte_off * 8u;
    lnb_br_refill(br);
    const uint32_t skip = (byte_off & 3u) * 8u;
    br.acc <<= skip; br.nbits -= skip;
}

LNB_HD uint32_t lnb_br_get(LnbBitReader &br, uint32_t n)   /* 0 <= n <= 32 */
{
    if (n == 0) return 0;
    if (br.nbits < n) lnb_br_refill(br);
    const uint32_t v = (uint32_t)(br.acc >> (64u - n));
    br.acc <<= n; br.nbits -= n;
    return v;
}

LNB_HD uint32_t lnb_br_peek(LnbBitReader &br, uint32_t n)  /* 1 <= n <= 32, does not consume */
{
    if (br.nbits < n) lnb_br_refill(br);
    return (uint32_t)(br.acc >> (64u - n));
}

LNB_HD void lnb_br_skip(LnbBitReader &br, uint32_t n) { br.acc <<= n; br.nbits -= n; }

/* zeros up to the next one bit; the one is consumed */
LNB_HD uint32_t lnb_br_zero_run(LnbBitReader &br)
{
    uint32_t run = 0;
    for (;;) {
        if (br.nbits == 0 || br.acc == 0) {
            run += br.nbits;
            br.acc = 0; br.nbits = 0;
            if (br.next_word >= br.end_word) { br.overrun = 1; return run; }
            lnb_br_refill(br);
            continue;
        }
        const uint32_t lz = lnb_clz64(br.acc);
        run += lz;
        /* shifting by lz+1 <= 64: split to stay defined when lz == 63 */
        br.acc <<= lz; br.acc <<= 1;
        br.nbits -= lz + 1u;
        return run;
    }
}

/* bytes consumed so far, rounded up (what reference BitStream_Flush + Tell report, bit_stream.h:402-406) */
LNB_HD uint32_t lnb_br_bytes_consumed(const LnbBitReader &br)
{
    const uint64_t loaded_bits = (uint64_t)br.next_word * 32u - br.start_bit;
    const uint64_t used = loaded_bits - br.nbits;
    return (uint32_t)((used + 7u) >> 3);
}

/* ------------------------------------------------------------------------------------------
 * Residual reader: the hot loop of the decoder (one thread walks a block's payload: the format
 * concatenates channels and delta-codes the Rice parameters inline, so there is no second entry
 * point into a block).  What matters is the length of the dependent instruction chain per symbol.
 * State: a 64-bit MSB-aligned window (hi:lo) with `nbits` valid bits (>= 32 between symbols) and one
 * prefetched word.  A symbol costs clz(hi) -> code length -> one funnel shift on the chain; the
 * refill (every ~3 symbols) ORs a prefetched word in.
 * Bit semantics: reference bit_stream.h:305-394; codes: linne_coder.c:106-127 (gamma), :150-169 (Rice).
 * ------------------------------------------------------------------------------------------ */
struct LnbFastReader {
    const uint32_t *words;
    uint32_t next, end_word;     /* index of the word held in `pre`; words at or past end_word read as zero */
    uint32_t hi, lo, pre;
    uint32_t nbits;              /* valid bits in hi:lo, 32..64 between symbols */
    uint32_t overrun;
};

LNB_HD uint32_t lnb_fr_word(const LnbFastReader &r, uint32_t i)
{
    return (i < r.end_word) ? lnb_bswap32(r.words[i]) : 0u;
}
/* (hi:lo) << n for 0 <= n <= 32: new high / low words */
LNB_HD uint32_t lnb_shl64_hi(uint32_t hi, uint32_t lo, uint32_t n)
{
#if defined(__CUDA_ARCH__)
    return __funnelshift_lc(lo, hi, n);
#else
    return (n >= 32u) ? lo : (uint32_t)((((uint64_t)hi << 32) | lo) >> (32u - n));
#endif
}
LNB_HD uint32_t lnb_shl32_clamped(uint32_t x, uint32_t n)      /* x << n, 0 for n >= 32 */
{
#if defined(__CUDA_ARCH__)
    return __funnelshift_lc(0u, x, n);
#else
    return (n >= 32u) ? 0u : x << n;
#endif
}
LNB_HD uint32_t lnb_shr32_clamped(uint32_t x, uint32_t n)      /* x >> n, 0 for n >= 32 */
{
#if defined(__CUDA_ARCH__)
    return __funnelshift_rc(x, 0u, n);
#else
    return (n >= 32u) ? 0u : x >> n;
#endif
}
LNB_HD void lnb_fr_refill(LnbFastReader &r)                     /* call when nbits < 32 */
{
    const uint32_t w = r.pre;
    r.hi |= lnb_shr32_clamped(w, r.nbits);
    r.lo = lnb_shl32_clamped(w, 32u - r.nbits);
    r.nbits += 32u;
    r.next++;
    r.pre = lnb_fr_word(r, r.next);
    if (r.next > r.end_word + 2u) r.overrun = 1;
}
LNB_HD void lnb_fr_open(LnbFastReader &r, const uint32_t *words, uint64_t bit_position, uint32_t end_word)
{
    r.words = words; r.end_word = end_word; r.overrun = 0;
    const uint32_t idx = (uint32_t)(bit_position >> 5), s = (uint32_t)(bit_position & 31u);
    const uint32_t w0 = lnb_fr_word(r, idx), w1 = lnb_fr_word(r, idx + 1u);
    r.hi = lnb_shl64_hi(w0, w1, s);
    r.lo = lnb_shl32_clamped(w1, s);
    r.nbits = 64u - s;
    r.next = idx + 2u;
    r.pre = lnb_fr_word(r, r.next);
}
LNB_HD void lnb_fr_skip(LnbFastReader &r, uint32_t n)           /* n <= 32 */
{
    r.hi = lnb_shl64_hi(r.hi, r.lo, n);
    r.lo = lnb_shl32_clamped(r.lo, n);
    r.nbits -= n;
    if (r.nbits < 32u) lnb_fr_refill(r);
}
LNB_HD uint32_t lnb_fr_get(LnbFastReader &r, uint32_t n)         /* 0 <= n <= 32 */
{
    const uint32_t v = lnb_shr32_clamped(r.hi, 32u - n);
    lnb_fr_skip(r, n);
    return v;
}
/* zeros up to the next one bit; the one is consumed */
LNB_HD uint32_t lnb_fr_zero_run(LnbFastReader &r)
{
    uint32_t run = 0;
    for (;;) {
        if (r.hi != 0u) {
            const uint32_t lz = lnb_clz32(r.hi);
            lnb_fr_skip(r, lz + 1u);
            return run + lz;
        }
        run += 32u;
        lnb_fr_skip(r, 32u);
        if (r.overrun) return run;
    }
}
LNB_HD uint64_t lnb_fr_position(const LnbFastReader &r) { return (uint64_t)r.next * 32u - r.nbits; }

LNB_HD uint32_t lnb_get_gamma(LnbFastReader &r)
{
    const uint32_t nd = lnb_fr_zero_run(r) + 1u;
    if (nd == 1u) return 0;
    if (nd > 32u) { r.overrun = 1; return 0; }
    return (uint32_t)(((uint64_t)1 << (nd - 1u)) + lnb_fr_get(r, nd - 1u) - 1u);
}

/* one recursive-Rice symbol, k1 = k2 + 1, 0 <= k2 <= 30.  Straight-line for codes of <= 32 bits:
 *   lz = leading zeros (0 -> '1' + k1 bits; lz >= 1 -> lz zeros, '1', k2 bits) */
LNB_HD uint32_t lnb_get_rice(LnbFastReader &r, uint32_t k1, uint32_t k2)
{
    const uint32_t hi = r.hi;
    const uint32_t lz = lnb_clz32(hi);
    const uint32_t kk = lz ? k2 : k1;
    const uint32_t total = lz + 1u + kk;
    if (total <= 32u) {                                          /* also implies hi != 0 */
        const uint32_t low = ((hi << lz) << 1 >> 1) >> (31u - kk);
        const uint32_t base = lz ? ((1u << k1) + ((lz - 1u) << k2)) : 0u;
        lnb_fr_skip(r, total);
        return low + base;
    }
    const uint32_t q = lnb_fr_zero_run(r);                       /* long code: zero run, then k2 bits */
    if (q == 0u) return lnb_fr_get(r, k1);
    return lnb_fr_get(r, k2) + (1u << k1) + ((q - 1u) << k2);
}

/* residual of one channel: reference linne_coder.c:306-327 */
LNB_HD void lnb_decode_residual(LnbFastReader &br, int32_t *out, uint32_t n)
{
    const uint32_t porder = lnb_fr_get(br, 10);
    const uint32_t len = (porder < 32u) ? (n >> porder) : 0u;
    const uint32_t parts = (porder <= LNB_MAX_PORDER) ? (1u << porder) : 0u;
    uint32_t k2 = 0;
    int32_t *dst = out;
    if (porder > LNB_MAX_PORDER) { br.overrun = 1; }
    for (uint32_t part = 0; part < parts; part++) {
        if (part == 0) k2 = lnb_fr_get(br, 5);
        else k2 = (uint32_t)((int32_t)k2 + lnb_zz_dec(lnb_get_gamma(br)));
        if (k2 > 30u) { br.overrun = 1; k2 = 30u; }           /* never produced by a valid encoder */
        const uint32_t k1 = k2 + 1u;
        for (uint32_t s = 0; s < len; s++) *dst++ = lnb_zz_dec(lnb_get_rice(br, k1, k2));
        if (br.overrun) break;
    }
    /* samples a broken porder leaves uncovered */
    for (uint32_t s = parts * len; s < n && (parts == 0 || br.overrun); s++) out[s] = 0;
}

/* ------------------------------------------------------------------------------------------
 * One block: framing already validated on the host (sync, size); this parses the payload.
 *   compressed: reference libs/linne_decoder/src/linne_decoder.c:457-497 (side info + residuals)
 *   raw:        :387-421      silent: :554-557
 * Residuals (or final PCM for raw/silent) are written into the PCM planes at the block's sample
 * offset; the synthesis kernels then work in place.
 * ------------------------------------------------------------------------------------------ */
LNB_HD void lnb_decode_block_payload(const LnbStreamCfg &cfg, const LnbDevTables &tab,
                                     const uint8_t *stream, uint32_t stream_size,
                                     LnbBlockDesc &blk, LnbChanParams *params /* [C] */, int32_t *pcm)
{
    const uint32_t C = cfg.num_channels, n = blk.nsmp;
    const uint32_t payload_off = blk.byte_off + LNB_BLOCK_HEADER_SIZE;
    uint32_t end_byte = blk.byte_off + blk.byte_size;
    if (end_byte > stream_size) end_byte = stream_size;

    if (blk.type == LNB_BLOCK_SILENT) {
        for (uint32_t c = 0; c < C; c++) {
            int32_t *dst = pcm + (size_t)c * cfg.pcm_stride + blk.smp_off;
            for (uint32_t i = 0; i < n; i++) dst[i] = 0;
        }
        blk.na = 0;
        return;
    }
    if (blk.type == LNB_BLOCK_RAW) {
        const uint32_t bytes = cfg.bits_per_sample >> 3;
        const uint8_t *p = stream + payload_off;
        if ((uint64_t)payload_off + (uint64_t)bytes * n * C > end_byte) { blk.status |= LNB_ST_OVERRUN; return; }
        for (uint32_t i = 0; i < n; i++)
            for (uint32_t c = 0; c < C; c++) {
                pcm[(size_t)c * cfg.pcm_stride + blk.smp_off + i] = lnb_zz_dec(lnb_get_be(p, (int)bytes));
                p += bytes;
            }
        blk.na = bytes * n * C;
        return;
    }
    if (blk.type != LNB_BLOCK_COMPRESSED) { blk.status |= LNB_ST_BAD_TYPE; return; }

    LnbBitReader br;
    lnb_br_open(br, (const uint32_t *)stream, payload_off, end_byte);
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
                const uint32_t e = tab.huff_lut[lnb_br_peek(br, LNB_HUFF_LUT_BITS)];
                lnb_br_skip(br, e & 15u);
                q[i] = (int8_t)lnb_zz_dec(e >> 4);
            }
        }
    LnbFastReader fr;
    lnb_fr_open(fr, (const uint32_t *)stream, (uint64_t)br.next_word * 32u - br.nbits, (end_byte + 3u) >> 2);
    for (uint32_t c = 0; c < C; c++)
        lnb_decode_residual(fr, pcm + (size_t)c * cfg.pcm_stride + blk.smp_off, n);
    /* payload bytes consumed, rounded up (reference Flush + Tell, bit_stream.h:402-406) */
    blk.na = (uint32_t)((lnb_fr_position(fr) - (uint64_t)payload_off * 8u + 7u) >> 3);
    if (br.overrun || fr.overrun || payload_off + blk.na > end_byte) blk.status |= LNB_ST_OVERRUN;
}

/* ------------------------------------------------------------------------------------------
 * Synthesis of one unit of one layer, in place: exact inverse of the predictor.
 * reference libs/linne_decoder/src/linne_lpc_synthesize.c:8-83.  int32 wraps (uint32 maths).
 * ------------------------------------------------------------------------------------------ */
LNB_HD void lnb_synthesize_unit(int32_t *x, uint32_t m, const int8_t *c, uint32_t p, uint32_t rshift)
{
    if (m <= p) return;
    const uint32_t half = rshift ? (1u << (rshift - 1u)) : 0u;
    for (uint32_t t = 0; t < m - p; t++) {
        uint32_t acc = half;
        for (uint32_t k = 0; k < p; k++) acc += (uint32_t)(int32_t)c[k] * (uint32_t)x[t + k];
        x[t + p] = (int32_t)((uint32_t)x[t + p] - (uint32_t)((int32_t)acc >> rshift));
    }
}

/* two cascaded first-order de-emphasis filters: reference linne_utility.c:215-241 */
LNB_HD void lnb_deemphasis(int32_t *x, uint32_t n, const int32_t prev[2], const uint8_t coef[2])
{
    const int32_t c0 = coef[0], c1 = coef[1];
    if (n == 0) return;
    x[0] += (prev[1] * c1) >> LNB_PREEM_SHIFT;
    if (n >= 2) x[1] += (x[0] * c1) >> LNB_PREEM_SHIFT;
    x[0] += (prev[0] * c0) >> LNB_PREEM_SHIFT;
    for (uint32_t i = 2; i < n; i++) {
        x[i] += (x[i - 1] * c1) >> LNB_PREEM_SHIFT;
        x[i - 1] += (x[i - 2] * c0) >> LNB_PREEM_SHIFT;
    }
    if (n >= 2) x[n - 1] += (x[n - 2] * c0) >> LNB_PREEM_SHIFT;
}

/* mid/side -> left/right for one sample pair: reference linne_utility.c:143-146 */
LNB_HD void lnb_ms_to_lr(int32_t &a, int32_t &b)
{
    a -= b >> 1;
    b += a;
}
