/* lnb_stream_v2.cuh -- fused streaming decoder: one CTA per block, the code-word walk feeding the synthesis
 * cascade sample by sample.
 *
 * Covers reference rows d2-d5 in ONE kernel: entropy decode
 * (libs/linne_decoder/src/linne_decoder.c:457-497, libs/linne_coder/src/linne_coder.c:306-327), synthesis of
 * every layer (linne_lpc_synthesize.c:8-83 in the order of linne_decoder.c:503-509), the two de-emphasis
 * filters (linne_utility.c:215-241) and mid/side -> left/right (:135-147).
 *
 * Every stage of a block's decode is a sequential recursion over the samples, and each only needs the PREVIOUS
 * stage's output at the same position -- so they form a pipeline of warps handing a shared-memory line on through
 * progress counters:
 *     stage 0         the WALK: where does each code word start?  The format gives a block one entry point
 *                     (channels concatenated, Rice parameters delta-coded inline), so this is the one truly serial
 *                     chain of the decoder.  It is kept as short as the code allows: ONE lane holds a 64-bit window
 *                     of the payload in registers; per code word it counts the leading zeros of the window, derives
 *                     the length  k2 + 1 + max(lz, 1)  and shifts -- four dependent instructions -- and stores the
 *                     32-bit window itself for the next stage.  (Round 1 let 32 lanes guess 32 code-word starts per
 *                     round; a round ended at the first long code word, ~3 retired per ~80 dependent instructions:
 *                     ~95 cycles per sample against ~25-30 here.)  The lane keeps its payload window in shared
 *                     memory filled by itself, four 2 KB chunks ahead, with bulk-asynchronous copies (lnb_bulk.cuh).
 *     stage 1         EXTRACT: 32 lanes turn 32 stored windows into residuals (everything about a code word except
 *                     its start is a function of its first 32 bits and the partition's k2).
 *     stage 2..L+1    synthesis layer L-1 .. 0, in place on the line (units walked one after the other; systolic
 *                     4-taps-per-lane form of lnb_synth_v2.cuh for units of 8 or more taps, one lane below)
 *     stage L+2       de-emphasis (+ M/S inverse on the second channel), then coalesced stores to the PCM planes
 * The latency of a block is the walk plus a pipeline lag of a few hundred samples.  A channel's line is reused by
 * the next channel once the last stage has drained it.
 *
 * Blocks longer than LNB_DS_MAX_N samples, and blocks whose CRC check failed, keep the split kernels.
 */
#pragma once
#include "lnb_common.cuh"
#include "lnb_decode_core.cuh"
#include "lnb_entropy_v3.cuh"
#include "lnb_synth_v2.cuh"
#include "lnb_tput_v2.cuh"
#include "lnb_bulk.cuh"

#define LNB_DS_MAX_N       10240u
#define LNB_DS_STAGES      (3u + LNB_MAX_LAYERS)        /* walk, extract, layers, de-emphasis */
#define LNB_DS_HW_WARPS    7u                           /* warp 4 stays empty: it would share a scheduler with the walk */
#define LNB_DS_THREADS     (32u * LNB_DS_HW_WARPS)
#define LNB_DS_BATCH       32u                          /* steps between progress updates */
#define LNB_DS_CHUNK_WORDS 512u                         /* payload window: chunks of 2 KB ... */
#define LNB_DS_CHUNKS      4u                           /* ... four of them in flight */
#define LNB_DS_RING_WORDS  (LNB_DS_CHUNK_WORDS * LNB_DS_CHUNKS)

struct __align__(128) LnbDsShared {
    uint32_t ring[LNB_DS_RING_WORDS];                  /* payload bytes as they are in the stream, word w of the block at w % RING */
    uint64_t full_bar[LNB_DS_CHUNKS];                  /* one transaction barrier per ring slot */
    volatile uint32_t prog[LNB_DS_STAGES];             /* samples (over all channels of the block) each stage has finished */
    volatile uint32_t abort;
    uint32_t porder[LNB_MAX_CHANNELS];                 /* partition order of each channel (walk -> extract) */
    uint32_t done_mask[LNB_DS_MAX_N / 32u];            /* samples the walk already stored as final values (long code words, fill) */
    uint8_t  k2tab[LNB_MAX_PARTITIONS];                /* Rice parameter of every partition of the current channel */
    uint16_t huff1[1u << LNB_E3_HUFF1_BITS];
    LnbChanParams params[LNB_MAX_CHANNELS];
};

__device__ __forceinline__ void lnb_ds_publish(LnbDsShared &sm, uint32_t stage, uint32_t value, uint32_t lane)
{
    __syncwarp();
    if (lane == 0) { __threadfence_block(); sm.prog[stage] = value; }
}
/* the same for a stage that runs on a single lane */
__device__ __forceinline__ void lnb_ds_publish_lane(LnbDsShared &sm, uint32_t stage, uint32_t value)
{
    __threadfence_block();
    sm.prog[stage] = value;
}
/* wait until stage `stage` has finished at least `need` samples; false when the block was aborted */
__device__ __forceinline__ bool lnb_ds_wait(LnbDsShared &sm, uint32_t stage, uint32_t need)
{
    for (;;) {
        const uint32_t have = sm.prog[stage];
        if (have >= need) break;
        if (sm.abort) return false;
        /* the pace maker (the walk) needs ~15 ns per sample: sleep roughly until the missing samples can exist,
         * so waiting stages leave the issue slots to the warps that have work */
        const uint32_t ns = (need - have) * 8u;
        __nanosleep(ns < 32u ? 32u : (ns > 2000u ? 2000u : ns));
    }
    __threadfence_block();
    return sm.abort == 0u;
}

/* ---- stage 0: the payload window and the walk ------------------------------------------------------------- */
/* Geometry of a block's window: plain scalars, so that everything the walking lane touches per code word stays
 * in registers and shared-memory address space. */
struct LnbDsWin {
    const uint8_t *g0;          /* 16-byte aligned global address of ring word 0 (at most 15 bytes before the block) */
    uint32_t end_word;          /* words at or past this index read as zero (end of the block) */
    uint32_t load_bytes;        /* bytes worth loading, a multiple of 16 */
    uint32_t nchunks;
};
__device__ __forceinline__ void lnb_ds_issue(LnbDsShared &sm, const LnbDsWin &w, uint32_t c)
{
    const uint32_t off = c * (LNB_DS_CHUNK_WORDS * 4u);
    const uint32_t left = w.load_bytes - off;
    const uint32_t bytes = left < LNB_DS_CHUNK_WORDS * 4u ? left : LNB_DS_CHUNK_WORDS * 4u;
    uint64_t *bar = &sm.full_bar[c % LNB_DS_CHUNKS];
    lnb_mbar_arrive_expect_tx(bar, bytes);
    lnb_bulk_load(&sm.ring[(c % LNB_DS_CHUNKS) * LNB_DS_CHUNK_WORDS], w.g0 + off, bytes, bar);
}
/* The walk reaches chunk c: keep four chunks in flight (chunk c - 1 has been read completely, its slot is free)
 * and make sure chunk c has landed.  Returns the new count of issued chunks.  Out of line: once per 2 KB. */
__device__ __noinline__ uint32_t lnb_ds_enter_chunk(LnbDsShared &sm, const uint8_t *g0, uint32_t load_bytes, uint32_t nchunks,
                                                    uint32_t issued, uint32_t c)
{
    LnbDsWin w; w.g0 = g0; w.end_word = 0; w.load_bytes = load_bytes; w.nchunks = nchunks;
    const uint32_t want = (c + LNB_DS_CHUNKS < nchunks) ? c + LNB_DS_CHUNKS : nchunks;
    while (issued < want) { lnb_ds_issue(sm, w, issued); issued++; }
    if (c < nchunks) lnb_mbar_wait(&sm.full_bar[c % LNB_DS_CHUNKS], (c / LNB_DS_CHUNKS) & 1u);
    return issued;
}
/* word i of the block's window, big-endian corrected (valid for chunks that have landed) */
__device__ __forceinline__ uint32_t lnb_ds_word(const LnbDsShared &sm, uint32_t end_word, uint32_t i)
{
    return (i < end_word) ? lnb_bswap32(sm.ring[i % LNB_DS_RING_WORDS]) : 0u;
}
__device__ __forceinline__ uint32_t lnb_ds_peek(const LnbDsShared &sm, uint32_t end_word, uint32_t pos)
{
    const uint32_t i = pos >> 5;
    return __funnelshift_l(lnb_ds_word(sm, end_word, i + 1u), lnb_ds_word(sm, end_word, i), pos & 31u);
}
__device__ __forceinline__ uint32_t lnb_ds_get(const LnbDsShared &sm, uint32_t end_word, uint32_t &pos, uint32_t n)   /* 1 <= n <= 32 */
{
    const uint32_t v = lnb_ds_peek(sm, end_word, pos) >> (32u - n);
    pos += n;
    return v;
}

/* The walking lane's reader: hi:lo = the next `cnt` (32..64) payload bits, left-aligned; `nw` = word `wi`, the next
 * one to be merged (fetched one merge ahead, so the shared-memory latency is off the chain).  A macro over plain
 * locals: nothing here may have its address taken. */
#define LNB_DS_TAKE(len_expr)                                                                         \
    do {                                                                                              \
        const uint32_t len_ = (len_expr);                                                             \
        hi = __funnelshift_lc(lo, hi, len_);                                                          \
        lo = __funnelshift_lc(0u, lo, len_);                                                          \
        cnt -= len_;                                                                                  \
        if (cnt < 32u) {                                                                              \
            hi |= nw >> cnt;                                                                          \
            lo = __funnelshift_r(0u, nw, cnt);             /* nw << (32 - cnt), 0 when cnt == 0 */    \
            cnt += 32u;                                                                               \
            wi++;                                                                                     \
            if ((wi % LNB_DS_CHUNK_WORDS) == 0u)                                                      \
                issued = lnb_ds_enter_chunk(sm, win.g0, win.load_bytes, win.nchunks, issued, wi / LNB_DS_CHUNK_WORDS); \
            nw = (wi < win.end_word) ? lnb_bswap32(lnb_lds32(ring_addr + ((wi * 4u) & (LNB_DS_RING_WORDS * 4u - 1u)))) : 0u; \
        }                                                                                             \
    } while (0)
/* One code word of at most 32 bits: store its window, step over it.  f = index of the leading one (31 - lz);
 * length = k2 + 1 + max(lz, 1) = (k2 + 32) - min(f, 30).  A longer code word (f < k2, or an all-zero window:
 * f = -1) leaves through `long_label`. */
#define LNB_DS_STEP(slot, long_label)                                                                 \
    do {                                                                                              \
        const uint32_t f_ = lnb_bfind(hi);                                                            \
        if (__builtin_expect((int32_t)f_ < (int32_t)k2, 0)) { at = (slot); goto long_label; }         \
        lnb_sts32(dst_addr + 4u * (slot), hi);                                                        \
        LNB_DS_TAKE(k2p32 - (f_ < 30u ? f_ : 30u));                                                   \
    } while (0)

/* one recursive-Rice residual from a code word's first 32 bits (whole code word inside them: lz <= 31 - k2) */
__device__ __forceinline__ int32_t lnb_ds_value(uint32_t hi, uint32_t k2)
{
    const uint32_t lz = lnb_clz32(hi);
    const uint32_t ml = (lz > 1u) ? lz : 1u;
    const uint32_t low = (hi >> ((31u - k2 - ml) & 31u)) & ((1u << k2) - 1u);
    const uint32_t mult = lz ? lz + 1u : ((hi >> 30) & 1u);
    return lnb_zz_dec((mult << k2) + low);
}

/* Stage 0 of one COMPRESSED block.  All lanes of the warp parse the side information (the same fields, uniform
 * control flow); lane 0 alone walks the residual code words.  `line` is the CTA's dynamic shared memory. */
__device__ __forceinline__ void lnb_ds_walk(const LnbDecodeBatch &b, LnbBlockDesc &gblk, const LnbBlockDesc &blk, LnbDsShared &sm,
                                            int32_t *line, uint32_t last_stage, uint32_t lane)
{
    const LnbStreamCfg &cfg = b.cfg;
    const uint32_t C = cfg.num_channels, n = blk.nsmp;
    LnbDsWin win;
    const uintptr_t addr = (uintptr_t)(b.stream + blk.byte_off);
    win.g0 = (const uint8_t *)(addr & ~(uintptr_t)15);
    const uint32_t rel0 = (uint32_t)(addr & 15u);
    const uint32_t rel_payload = rel0 + LNB_BLOCK_HEADER_SIZE;
    uint32_t end_byte = blk.byte_off + blk.byte_size;
    if (end_byte > b.stream_size || end_byte < blk.byte_off) end_byte = b.stream_size;
    const uint32_t rel_end = rel0 + (end_byte - blk.byte_off);
    win.end_word = (rel_end + 3u) >> 2;
    {   /* the image is followed by >= 16 readable bytes (lnb_types.h): whole 16-byte lines up to there */
        const uintptr_t readable = ((uintptr_t)(b.stream + b.stream_size) + 16u) & ~(uintptr_t)15;
        const uint64_t room = (uint64_t)(readable - (uintptr_t)win.g0);
        uint64_t want = ((uint64_t)rel_end + 15u) & ~(uint64_t)15;
        if (want > room) want = room;
        win.load_bytes = (uint32_t)want;
    }
    win.nchunks = (win.load_bytes + LNB_DS_CHUNK_WORDS * 4u - 1u) / (LNB_DS_CHUNK_WORDS * 4u);
    uint32_t issued = win.nchunks < LNB_DS_CHUNKS ? win.nchunks : LNB_DS_CHUNKS;
    if (lane == 0) for (uint32_t c = 0; c < issued; c++) lnb_ds_issue(sm, win, c);
    /* the side information lies inside the first two chunks (<= 2.2 KB for 8 channels of 24 bits at -m 7) */
    for (uint32_t c = 0; c < issued; c++) lnb_mbar_wait(&sm.full_bar[c], 0u);
    uint32_t pos = rel_payload * 8u;
    uint32_t overrun = 0;

    /* ---- side information (linne_decoder.c:457-486): every lane reads the same fields ---- */
    {
        LnbChanParams *params = sm.params;
        for (uint32_t c = 0; c < C; c++)
            for (int f = 0; f < LNB_NUM_PREEM; f++) {
                const int32_t prev = lnb_zz_dec(lnb_ds_get(sm, win.end_word, pos, cfg.bits_per_sample + 1u));
                const uint32_t coef = lnb_ds_get(sm, win.end_word, pos, LNB_PREEM_SHIFT - 1);
                if (lane == 0) { params[c].preem_prev[f] = prev; params[c].preem_coef[f] = (uint8_t)coef; }
            }
        for (uint32_t c = 0; c < C; c++)
            for (uint32_t l = 0; l < cfg.num_layers; l++) {
                const uint32_t P = cfg.layer_params[l];
                const uint32_t lu = lnb_ds_get(sm, win.end_word, pos, 3), rs = lnb_ds_get(sm, win.end_word, pos, 4);
                if (lane == 0) { params[c].log2_units[l] = (uint8_t)lu; params[c].rshift[l] = (uint8_t)rs; }
                int8_t *q = params[c].coef + l * LNB_MAX_PARAMS;
                /* lane j keeps coefficients j, j+32, ... */
                for (uint32_t i0 = 0; i0 < P; i0 += 32u) {
                    int32_t mine = 0;
                    const uint32_t lim = (P - i0 < 32u) ? P - i0 : 32u;
                    for (uint32_t i = 0; i < lim; i++) {
                        const uint32_t top = lnb_ds_peek(sm, win.end_word, pos);
                        uint32_t e = sm.huff1[top >> (32 - LNB_E3_HUFF1_BITS)];
                        if (e == 0u) e = b.tab.huff_lut[top >> (32 - LNB_HUFF_LUT_BITS)];
                        pos += e & 15u;
                        if (i == lane) mine = lnb_zz_dec(e >> 4);
                    }
                    if (lane < lim) q[i0 + lane] = (int8_t)mine;
                }
            }
    }
    __syncwarp();

    /* ---- residuals, channel after channel (linne_coder.c:306-327): lane 0 walks ---- */
    if (lane == 0) {
        uint32_t hi, lo, cnt, nw, wi;
        const uint32_t ring_addr = lnb_smem_addr(sm.ring), line_addr = lnb_smem_addr(line);
        {
            const uint32_t s = pos & 31u;
            wi = pos >> 5;
            const uint32_t w0 = lnb_ds_word(sm, win.end_word, wi), w1 = lnb_ds_word(sm, win.end_word, wi + 1u);
            hi = __funnelshift_l(w1, w0, s);
            lo = w1 << s;
            cnt = 64u - s;
            wi += 2u;
            if (wi / LNB_DS_CHUNK_WORDS) issued = lnb_ds_enter_chunk(sm, win.g0, win.load_bytes, win.nchunks, issued, wi / LNB_DS_CHUNK_WORDS);
            nw = lnb_ds_word(sm, win.end_word, wi);
        }
        for (uint32_t c = 0; c < C && !overrun; c++) {
            /* the line (and the k2 table) is free once the last stage has drained the previous channel */
            if (!lnb_ds_wait(sm, last_stage, c * n)) { overrun = 1; break; }
            const uint32_t gbase = c * n;
            const uint32_t porder = hi >> 22;
            LNB_DS_TAKE(10u);
            if (porder > LNB_MAX_PORDER) { overrun = 1; break; }
            sm.porder[c] = porder;
            const uint32_t len = n >> porder, parts = 1u << porder;
            uint32_t k2 = 0, done = 0, published = 0;
            for (uint32_t part = 0; part < parts; part++) {
                if (part == 0) {
                    k2 = hi >> 27;
                    LNB_DS_TAKE(5u);
                } else {                                         /* gamma code of zigzag(k2 - previous k2) */
                    const uint32_t lz = lnb_clz32(hi);
                    if (lz > 15u) { overrun = 1; break; }
                    const uint32_t v = ((hi << lz) >> (31u - lz)) - 1u;
                    LNB_DS_TAKE(2u * lz + 1u);
                    k2 = (uint32_t)((int32_t)k2 + lnb_zz_dec(v));
                }
                if (k2 > 30u) { overrun = 1; break; }
                sm.k2tab[part] = (uint8_t)k2;
                const uint32_t k2p32 = k2 + 32u;
                uint32_t rem = len;
                while (rem) {
                    /* groups of eight keep the loop counter and the store address off the chain; a long code word
                     * leaves the group, is finished below and the walk resumes behind it */
                    uint32_t at = 0;
                    uint32_t dst_addr = line_addr + 4u * done;
                    if (rem >= 8u) {
                        LNB_DS_STEP(0u, long_cw); LNB_DS_STEP(1u, long_cw); LNB_DS_STEP(2u, long_cw); LNB_DS_STEP(3u, long_cw);
                        LNB_DS_STEP(4u, long_cw); LNB_DS_STEP(5u, long_cw); LNB_DS_STEP(6u, long_cw); LNB_DS_STEP(7u, long_cw);
                        rem -= 8u; done += 8u;
                    } else {
                        LNB_DS_STEP(0u, long_cw);
                        rem -= 1u; done += 1u;
                    }
                    if (done - published >= 64u) { lnb_ds_publish_lane(sm, 0u, gbase + done); published = done; }
                    continue;
                long_cw:
                    {   /* code word longer than 32 bits (rare): finished here, marked as final for the next stage */
                        uint32_t q = 0;
                        while (hi == 0u) {
                            q += 32u;
                            LNB_DS_TAKE(32u);
                            if (wi > win.end_word + 2u) { overrun = 1; break; }
                        }
                        if (overrun) break;
                        const uint32_t z = lnb_clz32(hi);
                        q += z;
                        LNB_DS_TAKE(z + 1u);
                        const uint32_t low = k2 ? (hi >> (32u - k2)) : 0u;
                        LNB_DS_TAKE(k2);
                        const uint32_t u = (q == 0u) ? 0u : low + (2u << k2) + ((q - 1u) << k2);   /* q >= 2 here */
                        const uint32_t i = done + at;
                        line[i] = lnb_zz_dec(u);
                        sm.done_mask[i >> 5] |= 1u << (i & 31u);
                        rem -= at + 1u; done += at + 1u;
                    }
                }
                if (done - published >= 32u || part + 1u == parts) { lnb_ds_publish_lane(sm, 0u, gbase + done); published = done; }
                if (overrun) break;
            }
            if (overrun) break;
            /* samples a partition order that does not divide the block leaves uncovered read as zero */
            for (uint32_t i = done; i < n; i++) { line[i] = 0; sm.done_mask[i >> 5] |= 1u << (i & 31u); }
            lnb_ds_publish_lane(sm, 0u, gbase + n);
        }
        if (overrun) { sm.abort = 1u; __threadfence_block(); }
        const uint32_t used = (wi * 32u - cnt - rel_payload * 8u + 7u) >> 3;
        gblk.na = used;                                          /* payload bytes consumed (reference Flush + Tell) */
        if (overrun || rel_payload + used > rel_end) gblk.status = blk.status | LNB_ST_OVERRUN;
    }
    __syncwarp();
}

/* ---- stage 1: stored code-word windows -> residuals, 32 at a time ---- */
__device__ bool lnb_ds_extract(LnbDsShared &sm, int32_t *line, uint32_t ch, uint32_t n)
{
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t gbase = ch * n;
    if (!lnb_ds_wait(sm, 0u, gbase + 1u)) return false;       /* the channel's partition order is known */
    const uint32_t porder = sm.porder[ch];
    const uint32_t len = n >> porder;
    const uint32_t magic = len >= 2u ? 0xFFFFFFFFu / len + 1u : 0u;   /* i / len == umulhi(i, magic) for i, len < 2^16 */
    for (uint32_t i0 = 0; i0 < n; i0 += LNB_DS_BATCH) {
        const uint32_t k = (n - i0 < LNB_DS_BATCH) ? n - i0 : LNB_DS_BATCH;
        if (!lnb_ds_wait(sm, 0u, gbase + i0 + k)) return false;
        const uint32_t mw = sm.done_mask[i0 >> 5];
        if (lane < k && !((mw >> lane) & 1u)) {
            const uint32_t i = i0 + lane;
            const uint32_t part = len >= 2u ? __umulhi(i, magic) : i;
            line[i] = lnb_ds_value((uint32_t)line[i], sm.k2tab[part]);
        }
        __syncwarp();
        if (mw && lane == 0) sm.done_mask[i0 >> 5] = 0u;
        lnb_ds_publish(sm, 1u, gbase + i0 + k, lane);
    }
    return true;
}

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
__global__ void __launch_bounds__(LNB_DS_THREADS) lnb_stream_v2_kernel(LnbDecodeBatch b, uint32_t n_max)
{
    extern __shared__ __align__(16) int32_t lnb_ds_line[];
    __shared__ LnbDsShared sm;
    const uint32_t hw_warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const uint32_t blk_i = blockIdx.x;
    LnbBlockDesc &gblk = b.blocks[blk_i];
    const LnbBlockDesc blk = gblk;
    const LnbStreamCfg &cfg = b.cfg;
    const uint32_t C = cfg.num_channels, n = blk.nsmp, L = cfg.num_layers;
    /* blocks this kernel does not take are left to the split kernels (same rule on both sides: b.fused_max_n) */
    if (blk.type != LNB_BLOCK_COMPRESSED || blk.status || n > n_max || n > b.fused_max_n || n == 0u || lnb_tp_takes(b, blk)) return;

    /* Warps map to the SM's four schedulers by index mod 4.  The walk (stage 0) is the pipeline's pace maker: it
     * gets scheduler 0 to itself (hardware warp 4 exits); extract and de-emphasis (light) share scheduler 1, the
     * first two synthesis layers -- one of them is the long one in every preset -- take schedulers 2 and 3, the
     * short third layer sits beside the first. */
    const uint32_t last = L + 2u;                              /* stage index of the de-emphasis warp */
    uint32_t stage;
    if (hw_warp == 0u) stage = 0u;
    else if (hw_warp == 1u) stage = 1u;
    else if (hw_warp == 2u) stage = 2u;
    else if (hw_warp == 3u) stage = (L >= 2u) ? 3u : 0xFFu;
    else if (hw_warp == 5u) stage = last;
    else if (hw_warp == 6u) stage = (L >= 3u) ? 4u : 0xFFu;
    else stage = 0xFFu;

    for (uint32_t i = threadIdx.x; i < (1u << LNB_E3_HUFF1_BITS); i += LNB_DS_THREADS) {
        const uint16_t e = b.tab.huff_lut[i << (LNB_HUFF_LUT_BITS - LNB_E3_HUFF1_BITS)];
        sm.huff1[i] = ((e & 15u) <= LNB_E3_HUFF1_BITS) ? e : (uint16_t)0;
    }
    for (uint32_t i = threadIdx.x; i < LNB_DS_MAX_N / 32u; i += LNB_DS_THREADS) sm.done_mask[i] = 0u;
    if (threadIdx.x < LNB_DS_STAGES) sm.prog[threadIdx.x] = 0u;
    if (threadIdx.x == 0) {
        sm.abort = 0u;
        for (uint32_t c = 0; c < LNB_DS_CHUNKS; c++) lnb_mbar_init(&sm.full_bar[c], 1u);
        lnb_mbar_init_fence();
    }
    __syncthreads();

    if (stage == 0xFFu) return;
    if (stage == 0u) {
        lnb_ds_walk(b, gblk, blk, sm, lnb_ds_line, last, lane);
        if (sm.abort) {                                        /* broken payload: the block reads as silence */
            for (uint32_t c = 0; c < C; c++) {
                int32_t *gout = b.pcm + (size_t)c * cfg.pcm_stride + blk.smp_off;
                for (uint32_t i = lane; i < n; i += 32u) gout[i] = 0;
            }
        }
        return;
    }
    /* side information is complete once the walk has published anything at all */
    for (uint32_t ch = 0; ch < C; ch++) {
        const uint32_t gbase = ch * n;
        if (stage == 1u) {
            if (!lnb_ds_extract(sm, lnb_ds_line, ch, n)) return;
            continue;
        }
        if (!lnb_ds_wait(sm, stage - 1u, gbase + 1u)) return;
        if (stage < last) {
            const uint32_t l = L - (stage - 1u);               /* layers run L-1 .. 0 (linne_decoder.c:503-509) */
            const LnbChanParams &prm = sm.params[ch];
            if (!lnb_ds_layer(sm, stage - 1u, stage, gbase, lnb_ds_line, n, cfg.layer_params[l], 1u << prm.log2_units[l],
                              prm.coef + l * LNB_MAX_PARAMS, prm.rshift[l])) return;
        } else {
            if (!lnb_ds_finish(sm, b, blk, stage - 1u, stage, ch, lnb_ds_line, n)) return;
        }
    }
}
