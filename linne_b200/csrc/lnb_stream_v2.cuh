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
 *                     ~95 cycles per sample.)  The payload reaches shared memory through bulk-asynchronous copies,
 *                     four 2 KB chunks ahead of the walk, issued by a loader warp (lnb_bulk.cuh).
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
#include "lnb_rice_warp.cuh"

#define LNB_DS_MAX_N       10240u
#define LNB_DS_STAGES      (3u + LNB_MAX_LAYERS)        /* walk, extract, layers, de-emphasis */
#define LNB_DS_HW_WARPS    8u                           /* seven roles and a warp that leaves: the walk has a scheduler to itself */
#define LNB_DS_THREADS     (32u * LNB_DS_HW_WARPS)
#define LNB_DS_BATCH       32u                          /* steps between progress updates */
#define LNB_DS_CHUNK_WORDS 512u                         /* payload window: chunks of 2 KB ... */
#define LNB_DS_CHUNKS      4u                           /* ... four of them in flight */
#define LNB_DS_RING_WORDS  (LNB_DS_CHUNK_WORDS * LNB_DS_CHUNKS)
#define LNB_DS_MIRROR      16u                          /* words of the ring's head repeated behind its end */

struct __align__(128) LnbDsShared {
    uint32_t ring[LNB_DS_RING_WORDS + LNB_DS_MIRROR];  /* big-endian corrected payload words, word w of the block at w % RING */
    uint64_t full_bar[LNB_DS_CHUNKS];                  /* one transaction barrier per ring slot */
    volatile uint32_t loaded_words;                    /* payload words [0, loaded_words) are in the ring (loader -> readers) */
    volatile uint32_t walk_word;                       /* the walk has left everything below this word (walk -> loader) */
    volatile uint32_t walk_done;
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
/* Every polling loop of this kernel gives up after LNB_DS_PATIENCE rounds (seconds of waiting: something is broken,
 * not slow), says where, and aborts the block -- a kernel that never ends would take the whole GPU with it. */
#define LNB_DS_PATIENCE (1u << 20)
__device__ __noinline__ void lnb_ds_give_up(LnbDsShared &sm, const char *where, uint32_t a, uint32_t b)
{
    printf("linne_b200: stream_v2 block %u gave up waiting in %s (%u, %u) loaded=%u walk_word=%u prog=%u %u %u %u %u %u\n",
           blockIdx.x, where, a, b, sm.loaded_words, sm.walk_word, sm.prog[0], sm.prog[1], sm.prog[2], sm.prog[3], sm.prog[4], sm.prog[5]);
    sm.abort = 1u; sm.walk_done = 1u; sm.loaded_words = 0xFFFFFFFFu;
    __threadfence_block();
}
#ifdef LNB_DS_TIMING
__device__ unsigned long long lnb_ds_waited[8];                 /* cycles each hardware warp of block 0 spent in stage waits */
__device__ unsigned int lnb_ds_count[8];                        /* block 0: fast groups of 8, of 4, generic groups, careful groups, partitions */
#define LNB_DS_COUNT(i) do { if (blockIdx.x == 0) lnb_ds_count[i]++; } while (0)
#define LNB_DS_T0() const long long t_wait0_ = clock64()
#define LNB_DS_T1() do { if (blockIdx.x == 0 && (threadIdx.x & 31u) == 0u) lnb_ds_waited[threadIdx.x >> 5] += (unsigned long long)(clock64() - t_wait0_); } while (0)
#else
#define LNB_DS_COUNT(i) do { } while (0)
#define LNB_DS_T0() do { } while (0)
#define LNB_DS_T1() do { } while (0)
#endif
__device__ __forceinline__ bool lnb_ds_wait(LnbDsShared &sm, uint32_t stage, uint32_t need)
{
    if (sm.prog[stage] >= need) { __threadfence_block(); return sm.abort == 0u; }
    LNB_DS_T0();
    for (uint32_t spins = 0;; spins++) {
        const uint32_t have = sm.prog[stage];
        if (have >= need) break;
        if (sm.abort) return false;
        if (spins > LNB_DS_PATIENCE) { lnb_ds_give_up(sm, "stage wait", stage, need); return false; }
        /* the pace maker (the walk) needs ~15 ns per sample: sleep roughly until the missing samples can exist,
         * so waiting stages leave the issue slots to the warps that have work */
        /* ~8 ns per missing sample when a block has an SM to itself; with several CTAs per SM the walk is ~4x slower and
         * every needless wake-up takes issue slots from somebody's pace maker */
        const uint32_t ns = (need - have) * (gridDim.x > 296u ? 40u : 8u);
        __nanosleep(ns < 32u ? 32u : (ns > 4000u ? 4000u : ns));
    }
    LNB_DS_T1();
    __threadfence_block();
    return sm.abort == 0u;
}

/* ---- the payload window: a ring of big-endian corrected words kept filled by the loader warp ------------------
 * Geometry of a block's window: ring word w holds payload word w (counted from g0, the 16-byte aligned address at or
 * below the block's first byte) modulo the ring size; LNB_DS_MIRROR words behind the ring repeat its head, so a
 * reader may run that far past the end without wrapping. */
struct LnbDsWin {
    const uint8_t *g0;
    uint32_t rel_payload, rel_end;   /* first payload byte / end of the block, relative to g0 */
    uint32_t end_word;               /* words at or past this index are not part of the block */
    uint32_t load_bytes;             /* bytes worth loading, a multiple of 16 */
    uint32_t nchunks;
};
__device__ __forceinline__ LnbDsWin lnb_ds_window(const LnbDecodeBatch &b, const LnbBlockDesc &blk)
{
    LnbDsWin w;
    const uintptr_t addr = (uintptr_t)(b.stream + blk.byte_off);
    w.g0 = (const uint8_t *)(addr & ~(uintptr_t)15);
    const uint32_t rel0 = (uint32_t)(addr & 15u);
    w.rel_payload = rel0 + LNB_BLOCK_HEADER_SIZE;
    uint32_t end_byte = blk.byte_off + blk.byte_size;
    if (end_byte > b.stream_size || end_byte < blk.byte_off) end_byte = b.stream_size;
    w.rel_end = rel0 + (end_byte - blk.byte_off);
    w.end_word = (w.rel_end + 3u) >> 2;
    /* the image is followed by >= 16 readable bytes (lnb_types.h): whole 16-byte lines up to there */
    const uintptr_t readable = ((uintptr_t)(b.stream + b.stream_size) + 16u) & ~(uintptr_t)15;
    const uint64_t room = (uint64_t)(readable - (uintptr_t)w.g0);
    uint64_t want = ((uint64_t)w.rel_end + 15u) & ~(uint64_t)15;
    if (want > room) want = room;
    w.load_bytes = (uint32_t)want;
    w.nchunks = (w.load_bytes + LNB_DS_CHUNK_WORDS * 4u - 1u) / (LNB_DS_CHUNK_WORDS * 4u);
    return w;
}

/* The loader (one warp): bulk-asynchronous copies of 2 KB chunks into the ring, four in flight; a chunk that has
 * landed is byte-swapped in place (the bit readers want big-endian words) and announced through `loaded_words`; a
 * slot is refilled once the walk has left the chunk it held (`walk_word`). */
__device__ void lnb_ds_loader(const LnbDsWin &w, LnbDsShared &sm, uint32_t lane)
{
    uint32_t issued = 0, swapped = 0, walk_chunk = 0, idle = 0;
    for (;;) {
        const uint32_t may = (walk_chunk + LNB_DS_CHUNKS < w.nchunks) ? walk_chunk + LNB_DS_CHUNKS : w.nchunks;
        if (lane == 0) {
            for (uint32_t c = issued; c < may; c++) {
                const uint32_t off = c * (LNB_DS_CHUNK_WORDS * 4u);
                const uint32_t left = w.load_bytes - off;
                const uint32_t bytes = left < LNB_DS_CHUNK_WORDS * 4u ? left : LNB_DS_CHUNK_WORDS * 4u;
                uint64_t *bar = &sm.full_bar[c % LNB_DS_CHUNKS];
                lnb_proxy_fence_async();                        /* the slot was read and written through the generic proxy */
                lnb_mbar_arrive_expect_tx(bar, bytes);
                lnb_bulk_load(&sm.ring[(c % LNB_DS_CHUNKS) * LNB_DS_CHUNK_WORDS], w.g0 + off, bytes, bar);
            }
        }
        if (issued < may) issued = may;
        if (swapped < issued) {
            const uint32_t c = swapped, slot = c % LNB_DS_CHUNKS;
            lnb_mbar_wait(&sm.full_bar[slot], (c / LNB_DS_CHUNKS) & 1u);
            uint32_t *dst = &sm.ring[slot * LNB_DS_CHUNK_WORDS];
#pragma unroll 4
            for (uint32_t i = lane; i < LNB_DS_CHUNK_WORDS; i += 32u) dst[i] = lnb_bswap32(dst[i]);
            __syncwarp();
            if (slot == 0u && lane < LNB_DS_MIRROR) sm.ring[LNB_DS_RING_WORDS + lane] = sm.ring[lane];
            __syncwarp();
            swapped++;
            if (lane == 0) { __threadfence_block(); sm.loaded_words = (swapped == w.nchunks) ? 0xFFFFFFFFu : swapped * LNB_DS_CHUNK_WORDS; }
            continue;
        }
        if (swapped == w.nchunks) break;
        /* every issued chunk has landed: wait for the walk to free a slot */
        if (sm.walk_done) break;
        const uint32_t wc = sm.walk_word / LNB_DS_CHUNK_WORDS;
        if (wc == walk_chunk) {
            __nanosleep(1500);                                  /* a chunk lasts the walk >= 50 us: no need to look often */
            if (++idle > LNB_DS_PATIENCE) { if (lane == 0) lnb_ds_give_up(sm, "loader", issued, swapped); break; }
        } else { walk_chunk = wc; idle = 0; }
    }
    /* nothing may still be in flight towards this CTA's shared memory when it exits */
    while (swapped < issued) { lnb_mbar_wait(&sm.full_bar[swapped % LNB_DS_CHUNKS], (swapped / LNB_DS_CHUNKS) & 1u); swapped++; }
    if (lane == 0 && w.nchunks == 0u) sm.loaded_words = 0xFFFFFFFFu;
}

/* wait until payload words [0, need) have landed; returns the count known to have landed */
__device__ __noinline__ uint32_t lnb_ds_await_words(LnbDsShared &sm, uint32_t need)
{
    uint32_t have, spins = 0;
    LNB_DS_T0();
    while ((have = sm.loaded_words) < need) {
        __nanosleep(100);
        if (++spins > LNB_DS_PATIENCE) { lnb_ds_give_up(sm, "payload wait", need, have); break; }
    }
    LNB_DS_T1();
    __threadfence_block();
    return sm.loaded_words;
}
/* payload word i (valid once landed) */
__device__ __forceinline__ uint32_t lnb_ds_word(const LnbDsShared &sm, uint32_t end_word, uint32_t i)
{
    return (i < end_word) ? sm.ring[i % LNB_DS_RING_WORDS] : 0u;
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

/* one recursive-Rice residual from a code word's first 32 bits (whole code word inside them: lz <= 31 - k2) */
__device__ __forceinline__ int32_t lnb_ds_value(uint32_t hi, uint32_t k2)
{
    const uint32_t lz = lnb_clz32(hi);
    const uint32_t ml = (lz > 1u) ? lz : 1u;
    const uint32_t low = (hi >> ((31u - k2 - ml) & 31u)) & ((1u << k2) - 1u);
    const uint32_t mult = lz ? lz + 1u : ((hi >> 30) & 1u);
    return lnb_zz_dec((mult << k2) + low);
}

/* ---- stage 0: the walk -----------------------------------------------------------------------------------------
 * What the hardware dictates (tools/ubench/lat_bench.cu, walk_bench.cu on a B200): a dependent integer instruction
 * issues 4 cycles after its producer, a warp instruction occupies its pipe (ALU or FMA) for 2 cycles, and a branch
 * on a freshly computed predicate costs ~20 cycles whether it is taken or not.  So the walk is written as GROUPS of
 * G code words of straight-line code without a single branch or select:
 *     f = index of the leading one of the window;  length = (k2 + 32) - f + (window >> 31)  [= k2 + 1 + max(lz, 1)]
 *     the 128-bit window (four registers) shifts left by the length -- four funnel shifts, all fed by that length
 * with "a code word was longer than 32 bits" and "the group ran out of window" accumulated arithmetically and looked
 * at ONCE per group, together with the loader's progress and the publishing of the walk's own.  After a group the
 * window is re-read from shared memory at the new bit position (the only place its latency shows).  A group that
 * turns out bad is simply walked again from its start position by the careful one-code-word-at-a-time reader.
 * The walk stores each code word's 32-bit window; stage 1 turns windows into residuals. */
template <int G>
__device__ __forceinline__ void lnb_ds_group(uint32_t &w0, uint32_t &w1, uint32_t &w2, uint32_t &w3, uint32_t dst_addr,
                                             uint32_t k2, uint32_t &T, uint32_t &bad)
{
    const uint32_t K = k2 + 32u;
#pragma unroll
    for (int s = 0; s < G; s++) {
        const uint32_t f = lnb_bfind(w0);
        const uint32_t t = w0 >> 31;
        lnb_sts32(dst_addr + 4u * (uint32_t)s, w0);
        const uint32_t L = K - f + t;
        asm("mad.hi.u32 %0, %1, 2, %0;" : "+r"(bad) : "r"(f - k2));   /* + 1 whenever f < k2 (f = -1 for an all-zero window); FMA pipe */
        w0 = __funnelshift_lc(w1, w0, L); w1 = __funnelshift_lc(w2, w1, L);
        w2 = __funnelshift_lc(w3, w2, L); w3 = __funnelshift_lc(0u, w3, L);
        T += L;
    }
}
/* the window at bit position pos (its words have landed) */
__device__ __forceinline__ void lnb_ds_reload(uint32_t ring_addr, uint32_t pos, uint32_t &w0, uint32_t &w1, uint32_t &w2, uint32_t &w3)
{
    const uint32_t a = ring_addr + ((pos >> 5) % LNB_DS_RING_WORDS) * 4u, sh = pos & 31u;
    const uint32_t v0 = lnb_lds32(a), v1 = lnb_lds32(a + 4u), v2 = lnb_lds32(a + 8u), v3 = lnb_lds32(a + 12u), v4 = lnb_lds32(a + 16u);
    w0 = __funnelshift_l(v1, v0, sh); w1 = __funnelshift_l(v2, v1, sh);
    w2 = __funnelshift_l(v3, v2, sh); w3 = __funnelshift_l(v4, v3, sh);
}

/* Stage 0 of one COMPRESSED block.  All lanes of the warp parse the side information (the same fields, uniform
 * control flow); lane 0 alone walks the residual code words.  `line` is the CTA's dynamic shared memory. */
__device__ __forceinline__ void lnb_ds_walk(const LnbDecodeBatch &b, LnbBlockDesc &gblk, const LnbBlockDesc &blk, const LnbDsWin &win,
                                            LnbDsShared &sm, int32_t *line, uint32_t last_stage, uint32_t lane)
{
    const LnbStreamCfg &cfg = b.cfg;
    const uint32_t C = cfg.num_channels, n = blk.nsmp;
    /* the side information lies inside the first two chunks (<= 2.2 KB for 8 channels of 24 bits at -m 7) */
    uint32_t loaded = lnb_ds_await_words(sm, (win.nchunks < 2u ? win.nchunks : 2u) * LNB_DS_CHUNK_WORDS);
    uint32_t pos = win.rel_payload * 8u;
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
        /* shared-window addresses, computed once: made opaque so that the compiler keeps them in registers instead of
         * re-deriving them (a special-register read of ~40 cycles) in front of every access inside the loops */
        uint32_t ring_addr, line_addr;
        asm volatile("mov.u32 %0, %1;" : "=r"(ring_addr) : "r"(lnb_smem_addr(sm.ring)));
        asm volatile("mov.u32 %0, %1;" : "=r"(line_addr) : "r"(lnb_smem_addr(line)));
        const uint32_t pos_limit = (win.end_word + 4u) * 32u;   /* nothing sane reads past this */
        uint32_t walk_chunk = 0;
        /* make the words a reader at `pos` may touch available (its window and one group of code words), tell the loader */
#define LNB_DS_TELL_LOADER()                                                                                   \
        do {                                                                                                  \
            if (((pos >> 5) / LNB_DS_CHUNK_WORDS) != walk_chunk) { walk_chunk = (pos >> 5) / LNB_DS_CHUNK_WORDS; sm.walk_word = pos >> 5; } \
        } while (0)
#define LNB_DS_SERVICE()                                                                                      \
        do {                                                                                                  \
            const uint32_t need_ = (pos >> 5) + 12u;                                                          \
            LNB_DS_TELL_LOADER();                               /* first: a loader that waits for us must not be waited for */ \
            if (need_ > loaded) loaded = lnb_ds_await_words(sm, need_ < win.nchunks * LNB_DS_CHUNK_WORDS ? need_ : win.nchunks * LNB_DS_CHUNK_WORDS); \
        } while (0)
        for (uint32_t c = 0; c < C && !overrun; c++) {
            /* the line (and the k2 table) is free once the last stage has drained the previous channel */
            if (!lnb_ds_wait(sm, last_stage, c * n)) { overrun = 1; break; }
            const uint32_t gbase = c * n;
            LNB_DS_SERVICE();
            const uint32_t first = lnb_ds_peek(sm, win.end_word, pos);
            const uint32_t porder = first >> 22;
            uint32_t k2 = (first >> 17) & 31u;                       /* first partition: k2 itself (linne_coder.c:313) */
            pos += 15u;
            if (porder > LNB_MAX_PORDER || k2 > 30u) { overrun = 1; break; }
            sm.porder[c] = porder;
            const uint32_t len = n >> porder, parts = 1u << porder;
            uint32_t done = 0, published = 0;
            uint32_t w0, w1, w2, w3;
            lnb_ds_reload(ring_addr, pos, w0, w1, w2, w3);
            for (uint32_t part = 0; part < parts && !overrun; part++) {
                uint32_t T = 0, bad = 0;
                const uint32_t k2_prev = k2;
                bool with_header = part != 0u;
                if (with_header) {                                   /* gamma code of zigzag(k2 - previous k2), branch-free */
                    const uint32_t lz = lnb_clz32(w0);
                    const uint32_t z = lz & 15u;
                    const uint32_t v = ((w0 << z) >> (31u - z)) - 1u;
                    const uint32_t L = 2u * z + 1u;
                    k2 = (uint32_t)((int32_t)k2 + lnb_zz_dec(v));
                    bad = (lz > 15u ? 1u : 0u) | (k2 > 30u ? 1u : 0u);
                    k2 &= 31u;
                    w0 = __funnelshift_lc(w1, w0, L); w1 = __funnelshift_lc(w2, w1, L);
                    w2 = __funnelshift_lc(w3, w2, L); w3 = __funnelshift_lc(0u, w3, L);
                    T = L;
                }
                sm.k2tab[part] = (uint8_t)k2;
                LNB_DS_COUNT(4); if (k2 <= 7u) LNB_DS_COUNT(5); else if (k2 == 8u) LNB_DS_COUNT(6); else LNB_DS_COUNT(7);
                /* code words per group: as many as safely fit the 96 bits a group may use (typical length k2 + 2, a few longer) */
                uint32_t gmax = (k2 <= 8u) ? 8u : ((k2 <= 19u) ? 4u : 2u);
                uint32_t rem = len;
                bool force_careful = false;
                /* the common case as a tight loop: whole groups of gmax code words, ONE data-dependent branch per group */
#define LNB_DS_FAST_LOOP(G)                                                                                   \
                while (rem >= G) {                                                                            \
                    lnb_ds_group<G>(w0, w1, w2, w3, line_addr + 4u * done, k2, T, bad);                       \
                    LNB_DS_COUNT(G == 8u ? 0 : 1);                                                            \
                    const uint32_t npos_ = pos + T, ndone_ = done + G;                                        \
                    if (__builtin_expect((bad | (T > 96u ? 1u : 0u) | ((npos_ >> 5) + 12u > loaded ? 1u : 0u) \
                                          | (ndone_ - published >= 64u ? 1u : 0u)) != 0u, 0)) {               \
                        if (bad | (T > 96u ? 1u : 0u)) { force_careful = true; break; }   /* nothing committed: the careful reader redoes the group */ \
                        pos = npos_; done = ndone_;                                                           \
                        LNB_DS_SERVICE();                                                                     \
                        if (done - published >= 32u) { lnb_ds_publish_lane(sm, 0u, gbase + done); published = done; } \
                    } else { pos = npos_; done = ndone_; }                                                    \
                    rem -= G; T = 0; with_header = false;                                                     \
                    lnb_ds_reload(ring_addr, pos, w0, w1, w2, w3);                                            \
                }
                if (gmax == 8u) { LNB_DS_FAST_LOOP(8u) }
                /* ... and a rest of four to seven code words the same way (short partitions: 20 = 8 + 8 + 4) */
                if (gmax >= 4u && !force_careful) { LNB_DS_FAST_LOOP(4u) }
#undef LNB_DS_FAST_LOOP
                while (rem) {
                    const uint32_t dst_addr = line_addr + 4u * done;
                    const uint32_t room = rem < gmax ? rem : gmax;
                    uint32_t g;
                    LNB_DS_COUNT(2);
                    if (force_careful) { g = room; bad = 1u; force_careful = false; }   /* the group that failed: min(gmax, rem) */
                    else if (room >= 8u) { lnb_ds_group<8>(w0, w1, w2, w3, dst_addr, k2, T, bad); g = 8u; }
                    else if (room >= 4u) { lnb_ds_group<4>(w0, w1, w2, w3, dst_addr, k2, T, bad); g = 4u; }
                    else if (room >= 2u) { lnb_ds_group<2>(w0, w1, w2, w3, dst_addr, k2, T, bad); g = 2u; }
                    else { lnb_ds_group<1>(w0, w1, w2, w3, dst_addr, k2, T, bad); g = 1u; }
                    /* the one look per group: long code word or header trouble, window exhausted, loader behind, time to publish */
                    if (__builtin_expect((bad | (T > 96u ? 1u : 0u)) != 0u, 0)) {
                        /* the careful reader: this group again from its first bit, one field per pass, every check in place */
                        LNB_DS_COUNT(3);
                        LNB_DS_SERVICE();
                        if (with_header) {
                            const uint32_t h = lnb_ds_peek(sm, win.end_word, pos);
                            const uint32_t lz = lnb_clz32(h);
                            if (lz > 15u) { overrun = 1; break; }
                            const uint32_t v = ((h << lz) >> (31u - lz)) - 1u;
                            pos += 2u * lz + 1u;
                            k2 = (uint32_t)((int32_t)k2_prev + lnb_zz_dec(v));
                            if (k2 > 30u) { overrun = 1; break; }
                            sm.k2tab[part] = (uint8_t)k2;
                            gmax = (k2 <= 8u) ? 8u : ((k2 <= 19u) ? 4u : 2u);
                        }
                        for (uint32_t s = 0; s < g; s++) {
                            LNB_DS_SERVICE();
                            const uint32_t h = lnb_ds_peek(sm, win.end_word, pos);
                            const uint32_t lz = lnb_clz32(h);
                            const uint32_t i = done + s;
                            if (lz + k2 <= 31u) {
                                line[i] = (int32_t)h;
                                pos += k2 + 1u + (lz > 1u ? lz : 1u);
                                continue;
                            }
                            /* code word longer than 32 bits: finished here, marked as final for the next stage */
                            uint32_t q = 0, hh = h;
                            while (hh == 0u) {
                                q += 32u; pos += 32u;
                                if (pos > pos_limit) { overrun = 1; break; }
                                LNB_DS_SERVICE();
                                hh = lnb_ds_peek(sm, win.end_word, pos);
                            }
                            if (overrun) break;
                            const uint32_t z = lnb_clz32(hh);
                            q += z; pos += z + 1u;
                            LNB_DS_SERVICE();
                            const uint32_t low = k2 ? (lnb_ds_peek(sm, win.end_word, pos) >> (32u - k2)) : 0u;
                            pos += k2;
                            line[i] = lnb_zz_dec(low + (2u << k2) + ((q - 1u) << k2));         /* q >= 2 here */
                            sm.done_mask[i >> 5] |= 1u << (i & 31u);
                        }
                        if (overrun) break;
                        bad = 0;
                    } else {
                        pos += T;
                    }
                    T = 0; with_header = false; rem -= g; done += g;
                    if (((pos >> 5) + 12u > loaded) | (done - published >= 64u)) {
                        LNB_DS_SERVICE();
                        if (done - published >= 32u) { lnb_ds_publish_lane(sm, 0u, gbase + done); published = done; }
                    }
                    lnb_ds_reload(ring_addr, pos, w0, w1, w2, w3);
                }
                if (overrun) break;
                LNB_DS_TELL_LOADER();
                if (done - published >= 32u || part + 1u == parts) { lnb_ds_publish_lane(sm, 0u, gbase + done); published = done; }
            }
            if (overrun) break;
            /* samples a partition order that does not divide the block leaves uncovered read as zero */
            for (uint32_t i = done; i < n; i++) { line[i] = 0; sm.done_mask[i >> 5] |= 1u << (i & 31u); }
            lnb_ds_publish_lane(sm, 0u, gbase + n);
        }
#undef LNB_DS_SERVICE
#undef LNB_DS_TELL_LOADER
        if (overrun) { sm.abort = 1u; __threadfence_block(); }
        const uint32_t used = (pos - win.rel_payload * 8u + 7u) >> 3;
        gblk.na = used;                                          /* payload bytes consumed (reference Flush + Tell) */
        if (overrun || win.rel_payload + used > win.rel_end) gblk.status = blk.status | LNB_ST_OVERRUN;
        sm.walk_done = 1u;
    }
    __syncwarp();
}

/* The walk as a WARP (kernel template argument WARP_WALK; LINNE_B200_WALK=warp): rounds of 32 code words that continue
 * past long code words (lnb_rice_warp.cuh) and leave RESIDUALS on the line -- the extract stage has nothing left to do,
 * its warp leaves and the walk publishes as stage 1.  Measured inside this kernel on the 10-second clip: 0.84-0.86 ms per
 * launch against 0.79-0.80 ms with the one-lane walk (a fix-up pass costs ~160 cycles -- leading-zero count, warp-wide
 * minimum through the uniform datapath, a branch on its result -- and a quarter of the code words need one), so the
 * one-lane walk stays the default; the warp form is kept as the measured alternative. */
template <int UNUSED = 0>
__device__ __forceinline__ void lnb_ds_walk_warp(const LnbDecodeBatch &b, LnbBlockDesc &gblk, const LnbBlockDesc &blk, const LnbDsWin &win,
                                            LnbDsShared &sm, int32_t *line, uint32_t last_stage, uint32_t lane)
{
    const LnbStreamCfg &cfg = b.cfg;
    const uint32_t C = cfg.num_channels, n = blk.nsmp;
    /* the side information lies inside the first two chunks (<= 2.2 KB for 8 channels of 24 bits at -m 7) */
    uint32_t loaded = lnb_ds_await_words(sm, (win.nchunks < 2u ? win.nchunks : 2u) * LNB_DS_CHUNK_WORDS);
    uint32_t pos = win.rel_payload * 8u;
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

    /* ---- residuals, channel after channel (linne_coder.c:306-327): the warp walks, 32 code words per round ---- */
    {
        uint32_t ring_addr, lane_key;
        asm volatile("mov.u32 %0, %1;" : "=r"(ring_addr) : "r"(lnb_smem_addr(sm.ring)));
        asm volatile("shl.b32 %0, %1, 8;" : "=r"(lane_key) : "r"(lane));
        const uint32_t pos_limit = (win.end_word + 4u) * 32u;   /* nothing sane reads past this */
        const uint32_t all_words = win.nchunks * LNB_DS_CHUNK_WORDS;
        uint32_t walk_chunk = 0;
        /* make the words a reader at `pos` may touch within `span` bits available, tell the loader where the walk is */
#define LNB_DS_SERVICE(span)                                                                                  \
        do {                                                                                                  \
            const uint32_t need_ = ((pos + (span)) >> 5) + 2u;                                                \
            if (((pos >> 5) / LNB_DS_CHUNK_WORDS) != walk_chunk) {  /* first: a loader that waits for us must not be waited for */ \
                walk_chunk = (pos >> 5) / LNB_DS_CHUNK_WORDS;                                                 \
                if (lane == 0) sm.walk_word = pos >> 5;                                                       \
            }                                                                                                 \
            if (need_ > loaded) loaded = lnb_ds_await_words(sm, need_ < all_words ? need_ : all_words);       \
        } while (0)
        for (uint32_t c = 0; c < C && !overrun; c++) {
            /* the line is free once the last stage has drained the previous channel */
            if (!lnb_ds_wait(sm, last_stage, c * n)) { overrun = 1; break; }
            const uint32_t gbase = c * n;
            LNB_DS_SERVICE(64u);
            const uint32_t first = lnb_ds_peek(sm, win.end_word, pos);
            const uint32_t porder = first >> 22;
            uint32_t k2 = (first >> 17) & 31u;                       /* first partition: k2 itself (linne_coder.c:313) */
            pos += 15u;
            if (porder > LNB_MAX_PORDER || k2 > 30u) { overrun = 1; break; }
            const uint32_t len = n >> porder, parts = 1u << porder;
            uint32_t done = 0;
            for (uint32_t part = 0; part < parts && !overrun; part++) {
                if (part) {                                          /* gamma code of zigzag(k2 - previous k2) */
                    LNB_DS_SERVICE(64u);
                    const uint32_t h = lnb_ds_peek(sm, win.end_word, pos);
                    const uint32_t lz = lnb_clz32(h);
                    if (lz > 15u) { overrun = 1; break; }
                    const uint32_t gv = ((h << lz) >> (31u - lz)) - 1u;
                    pos += 2u * lz + 1u;
                    k2 = (uint32_t)((int32_t)k2 + lnb_zz_dec(gv));
                    if (k2 > 30u) { overrun = 1; break; }
                }
                LNB_DS_COUNT(4);
                uint32_t left = len;
                while (left) {
                    if (pos > pos_limit) { overrun = 1; break; }
                    const uint32_t R = left < 32u ? left : 32u;
                    LNB_DS_SERVICE(32u * (k2 + 2u) + 96u);
                    uint32_t v, bits;
                    const uint32_t n_ok = lnb_rw_round(ring_addr, LNB_DS_RING_WORDS, pos, R, k2, lane_key, v, bits);
                    if (lane < n_ok) line[done + lane] = lnb_rw_value(v, k2);
                    pos += bits; done += n_ok; left -= n_ok;
                    LNB_DS_COUNT(0);
                    if (n_ok < R) {                              /* the code word behind them, read with every check in place (rare) */
                        LNB_DS_COUNT(3);
                        LNB_DS_SERVICE(64u);
                        const uint32_t h = lnb_ds_peek(sm, win.end_word, pos);
                        const uint32_t lz = lnb_clz32(h);
                        int32_t val;
                        if (lz + k2 <= 31u) {
                            val = lnb_rw_value(h, k2);
                            pos += k2 + 1u + (lz > 1u ? lz : 1u);
                        } else {                                 /* longer than 32 bits */
                            uint32_t q = 0, hh = h;
                            while (hh == 0u) {
                                q += 32u; pos += 32u;
                                if (pos > pos_limit) { overrun = 1; break; }
                                LNB_DS_SERVICE(64u);
                                hh = lnb_ds_peek(sm, win.end_word, pos);
                            }
                            if (overrun) break;
                            const uint32_t z = lnb_clz32(hh);
                            q += z; pos += z + 1u;
                            LNB_DS_SERVICE(64u);
                            const uint32_t low = k2 ? (lnb_ds_peek(sm, win.end_word, pos) >> (32u - k2)) : 0u;
                            pos += k2;
                            val = lnb_zz_dec(low + (2u << k2) + ((q - 1u) << k2));             /* q >= 2 here */
                        }
                        if (lane == 0) line[done] = val;
                        done++; left--;
                    }
                    lnb_ds_publish(sm, 1u, gbase + done, lane);
                }
            }
            if (overrun) break;
            /* samples a partition order that does not divide the block leaves uncovered read as zero */
            for (uint32_t i = done + lane; i < n; i += 32u) line[i] = 0;
            lnb_ds_publish(sm, 1u, gbase + n, lane);
        }
#undef LNB_DS_SERVICE
        if (lane == 0) {
            if (overrun) { sm.abort = 1u; __threadfence_block(); }
            const uint32_t used = (pos - win.rel_payload * 8u + 7u) >> 3;
            gblk.na = used;                                      /* payload bytes consumed (reference Flush + Tell) */
            if (overrun || win.rel_payload + used > win.rel_end) gblk.status = blk.status | LNB_ST_OVERRUN;
            sm.walk_done = 1u;
        }
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

__device__ unsigned int lnb_ds_sm_ticket[256];                  /* per SM: CTAs of this kernel that started there (role rotation) */

/* One CTA per block.  Dynamic shared memory: the channel line, n_max int32. */
template <bool WARP_WALK>
__global__ void __launch_bounds__(LNB_DS_THREADS, 4) lnb_stream_v2_kernel(LnbDecodeBatch b, uint32_t n_max)
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

    /* Warps map to the SM's four schedulers by index mod 4, and a role keeps its scheduler busy to a very different
     * degree: the walk (stage 0) is the pace maker and issue-bound on its own (it shares its scheduler with the warp
     * that leaves), the long synthesis layer comes next; loader, extract and de-emphasis are light.  Virtual warp v:
     *     v0 walk | v4 (leaves)      v1 extract | v5 loader      v2 first layer | v6 third (short) layer      v3 second
     *     layer -- the long one in every three-layer preset | v7 de-emphasis
     * Several CTAs share an SM (four fit); if all of them put their walk on scheduler 0 it would serve four pace makers
     * while the others idle.  So every CTA draws a ticket from its SM and rotates its roles by it. */
    __shared__ uint32_t s_rot;
    if (threadIdx.x == 0) {
        uint32_t smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        s_rot = atomicAdd(&lnb_ds_sm_ticket[smid & 255u], 1u) & 3u;
    }
    for (uint32_t i = threadIdx.x; i < (1u << LNB_E3_HUFF1_BITS); i += LNB_DS_THREADS) {
        const uint16_t e = b.tab.huff_lut[i << (LNB_HUFF_LUT_BITS - LNB_E3_HUFF1_BITS)];
        sm.huff1[i] = ((e & 15u) <= LNB_E3_HUFF1_BITS) ? e : (uint16_t)0;
    }
    for (uint32_t i = threadIdx.x; i < LNB_DS_MAX_N / 32u; i += LNB_DS_THREADS) sm.done_mask[i] = 0u;
    if (threadIdx.x < LNB_DS_STAGES) sm.prog[threadIdx.x] = 0u;
    if (threadIdx.x == 0) {
        sm.abort = 0u; sm.loaded_words = 0u; sm.walk_word = 0u; sm.walk_done = 0u;
        for (uint32_t c = 0; c < LNB_DS_CHUNKS; c++) lnb_mbar_init(&sm.full_bar[c], 1u);
        lnb_mbar_init_fence();
    }
    __syncthreads();

    const uint32_t vwarp = (hw_warp + 8u - s_rot) & 7u;
    const uint32_t last = L + 2u;                              /* stage index of the de-emphasis warp */
    uint32_t stage;
    if (vwarp == 0u) stage = WARP_WALK ? 1u : 0u;               /* the warp walk delivers residuals: it is stage 1 and v1 leaves */
    else if (vwarp == 1u) stage = WARP_WALK ? 0xFFu : 1u;
    else if (vwarp == 2u) stage = 2u;
    else if (vwarp == 3u) stage = (L >= 2u) ? 3u : 0xFFu;
    else if (vwarp == 6u) stage = (L >= 3u) ? 4u : 0xFFu;
    else if (vwarp == 7u) stage = last;
    else stage = 0xFFu;                                        /* v4 leaves, v5 is the loader (below) */

#ifdef LNB_DS_TIMING
    const long long t_kernel0 = clock64();
    if (blockIdx.x == 0 && lane == 0) { lnb_ds_waited[hw_warp] = 0; if (hw_warp == 0) for (int i = 0; i < 8; i++) lnb_ds_count[i] = 0; }
    struct LnbDsTimingReport { long long t0; uint32_t w; __device__ ~LnbDsTimingReport() {
        if (blockIdx.x == 0 && (threadIdx.x & 31u) == 0u) printf("stream_v2 timing: warp %u ran %lld cycles, waited %llu; counts %u %u %u %u parts %u (k2<=7 %u, 8 %u, more %u)\n", w, clock64() - t0, lnb_ds_waited[w], lnb_ds_count[0], lnb_ds_count[1], lnb_ds_count[2], lnb_ds_count[3], lnb_ds_count[4], lnb_ds_count[5], lnb_ds_count[6], lnb_ds_count[7]); } } report_{t_kernel0, hw_warp};
#endif
#ifdef LNB_DS_TIMING
    if (b.cfg.check_crc & 0x100u) {                            /* debug: the walk alone (results are wrong) */
        if (vwarp != 0u && vwarp != 5u) return;
        if (threadIdx.x == 0) for (uint32_t i = WARP_WALK ? 2u : 1u; i < LNB_DS_STAGES; i++) sm.prog[i] = 0x7FFFFFFFu;
    }
#endif
    if (vwarp == 5u) { lnb_ds_loader(lnb_ds_window(b, blk), sm, lane); return; }
    if (stage == 0xFFu) return;
    if (vwarp == 0u) {
        if (WARP_WALK) lnb_ds_walk_warp(b, gblk, blk, lnb_ds_window(b, blk), sm, lnb_ds_line, last, lane);
        else lnb_ds_walk(b, gblk, blk, lnb_ds_window(b, blk), sm, lnb_ds_line, last, lane);
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
        if (!WARP_WALK && stage == 1u) {
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
