/* lnb_rice_warp.cuh -- a warp decodes the recursive-Rice code words of ONE block, 32 per round.
 *
 * Reference behaviour: libs/linne_coder/src/linne_coder.c:150-169 (RecursiveRice_GetCode) inside the partition loop of
 * :306-327; bit order of libs/bit_stream/include/bit_stream.h:354-394.
 *
 * With k1 = k2 + 1 a code word is  q zeros, a one, then k1 bits (q = 0) or k2 bits (q >= 1):  k2 + 2 bits for q <= 1 and
 * k2 + 2 + x bits with the EXTRA  x = q - 1  for the long ones (q >= 2, about a quarter of all code words by the way the
 * encoder picks k2).  Where a code word starts depends on the extras of all code words in front of it -- one serial
 * chain per block.  A round resolves that chain for 32 code words with one warp-wide minimum per LONG code word:
 *     lane j keeps a 64-bit window of the payload at its guess  pos + j * (k2 + 2)  and an offset e_j (0 at first);
 *     every lane reads the extra of the code word it sees at offset e_j;  the first lane (not yet final) that sees a
 *     long one is right about it, because everything in front of it is final: it and the lanes before it become
 *     final, the lanes behind it add its extra to their offset (a shift inside their window, no memory access) and
 *     look again -- until no lane sees a long code word.
 * The round ends early at a code word that does not fit the 32 bits a lane looks at (finished by the caller's serial
 * reader) or when the extras outgrow the 32 bits of slack in the windows.  Measured alone (tools/ubench/
 * warp_walk_bench.cu, one warp, a quarter of the code words long): 55 cycles per code word latency-bound -- no better
 * than one lane walking alone -- but a fraction of its warp instructions per code word, which is what the issue-bound
 * throughput decoder (lnb_tput_v2.cuh) needs: it runs a round on thousands of warps.
 *
 * The payload of a block reaches the lanes through a per-warp ring in shared memory that the warp fills itself: 16-byte
 * loads one chunk ahead (held in registers while the previous chunk is consumed), byte-swapped on the way in.
 */
#pragma once
#include "lnb_common.cuh"
#include "lnb_bulk.cuh"

#if defined(__CUDACC__)

#define LNB_RW_RING   256u          /* words of a warp's payload ring */
#define LNB_RW_CHUNK  128u          /* words per refill: one 16-byte load per lane */
#define LNB_RW_MIRROR 4u            /* words of the ring's head repeated behind its end (a lane reads three words in a row) */
#define LNB_RW_WORDS  (LNB_RW_RING + LNB_RW_MIRROR)

struct LnbRwRing {
    uint32_t saddr;                 /* shared-memory address of the ring */
    const uint4 *g;                 /* 16-byte aligned address at or below the block's first byte */
    uint32_t lines;                 /* readable 16-byte lines from there */
    uint32_t end_word;              /* words at or past this index are not part of the block: they read as zero */
    uint32_t filled;                /* words [0, filled) have been stored into the ring */
    uint4 pre;                      /* this lane's four words of chunk [filled, filled + CHUNK), in flight */
};
__device__ __forceinline__ uint4 lnb_rw_fetch(const LnbRwRing &r, uint32_t first_word, uint32_t lane)
{
    const uint32_t i16 = first_word / 4u + lane;
    return (i16 < r.lines) ? __ldg(r.g + i16) : make_uint4(0u, 0u, 0u, 0u);
}
__device__ __forceinline__ void lnb_rw_open(LnbRwRing &r, uint32_t *ring, const uint8_t *g0, uint32_t lines, uint32_t end_word, uint32_t lane)
{
    r.saddr = lnb_smem_addr(ring);
    r.g = (const uint4 *)g0; r.lines = lines; r.end_word = end_word; r.filled = 0u;
    r.pre = lnb_rw_fetch(r, 0u, lane);
}
__device__ __forceinline__ void lnb_rw_advance(LnbRwRing &r, uint32_t lane)
{
    const uint32_t w = r.filled + lane * 4u;
    const uint32_t a = r.saddr + (w % LNB_RW_RING) * 4u;
    const uint32_t v0 = (w < r.end_word) ? lnb_bswap32(r.pre.x) : 0u, v1 = (w + 1u < r.end_word) ? lnb_bswap32(r.pre.y) : 0u;
    const uint32_t v2 = (w + 2u < r.end_word) ? lnb_bswap32(r.pre.z) : 0u, v3 = (w + 3u < r.end_word) ? lnb_bswap32(r.pre.w) : 0u;
    __syncwarp();                                                /* everybody is done reading the chunk this one replaces */
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" :: "r"(a), "r"(v0), "r"(v1), "r"(v2), "r"(v3) : "memory");
    if ((w % LNB_RW_RING) == 0u)
        asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" :: "r"(r.saddr + LNB_RW_RING * 4u), "r"(v0), "r"(v1), "r"(v2), "r"(v3) : "memory");
    r.filled += LNB_RW_CHUNK;
    r.pre = lnb_rw_fetch(r, r.filled, lane);
    __syncwarp();
}
/* make every word up to and including `need_word` readable (a reader may not look back further than
 * LNB_RW_RING - LNB_RW_CHUNK words behind it) */
__device__ __forceinline__ void lnb_rw_ensure(LnbRwRing &r, uint32_t need_word, uint32_t lane)
{
    while (need_word >= r.filled) lnb_rw_advance(r, lane);
}
__device__ __forceinline__ uint32_t lnb_rw_peek(const LnbRwRing &r, uint32_t pos)
{
    const uint32_t a = r.saddr + ((pos >> 5) % LNB_RW_RING) * 4u;
    return __funnelshift_l(lnb_lds32(a + 4u), lnb_lds32(a), pos);
}
__device__ __forceinline__ uint32_t lnb_rw_get(const LnbRwRing &r, uint32_t &pos, uint32_t n)   /* 1 <= n <= 32 */
{
    const uint32_t v = lnb_rw_peek(r, pos) >> (32u - n);
    pos += n;
    return v;
}

/* One round: up to R (1..32) code words of parameter k2 (<= 30) from bit position pos; the ring holds every word up to
 * (pos + 32 * (k2 + 2) + 96) / 32.  Returns how many code words are final (lane j < n_ok holds the first 32 bits of
 * its code word in v, all of it inside them) and the bits they take.  n_ok < R: the code word behind them does not
 * fit 32 bits, or the extras outgrew the windows -- the caller reads one code word serially and starts a new round.
 * lane_key = lane << 8 (kept by the caller so that it is not rebuilt per round).
 * The loop itself looks at nothing but "does a live lane see a long code word": a lane whose offset has left its
 * window (e > 32: the clamped shift hands it the low word) or whose code word does not fit computes garbage that only
 * reaches lanes behind it, and those are cut off after the loop by one vote. */
__device__ __forceinline__ uint32_t lnb_rw_round(uint32_t ring_saddr, uint32_t ring_words, uint32_t pos, uint32_t R, uint32_t k2,
                                                 uint32_t lane_key, uint32_t &v, uint32_t &bits)
{
    const uint32_t L = k2 + 2u;
    const uint32_t rel = (lane_key >> 8) * L;
    const uint32_t p = pos + rel;
    const uint32_t a = ring_saddr + ((p >> 5) % ring_words) * 4u;
    const uint32_t w0 = lnb_lds32(a), w1 = lnb_lds32(a + 4u), w2 = lnb_lds32(a + 8u);
    const uint32_t hi = __funnelshift_l(w1, w0, p), lo = __funnelshift_l(w2, w1, p);
    uint32_t e = 0u;
    uint32_t klive = (lane_key < (R << 8)) ? lane_key : 0xFFFFFFFFu;
    int32_t t;
    for (;;) {
        v = __funnelshift_lc(lo, hi, e);
        /* t = extra - 1: -1 for the short code words (at most one leading zero), 30 for an all-zero window */
        t = (int32_t)__clz((int)v) - 2;
        t = t < -1 ? -1 : t;
        const uint32_t r = __reduce_min_sync(0xffffffffu, klive | (uint32_t)t);
        if (r == 0xFFFFFFFFu) break;                             /* nobody (still live) sees a long code word */
        if (lane_key > r) e += (r & 255u) + 1u; else klive = 0xFFFFFFFFu;    /* lanes behind it follow; the others are final */
    }
    /* the first lane that cannot vouch for its code word ends the round: window left (e > 32), code word longer than the
     * 32 bits in v (extra - 1 > 29 - k2; an all-zero window always is), or beyond the R code words asked for */
    const bool invalid = e > 32u || t > 29 - (int32_t)k2 || lane_key >= (R << 8);
    const uint32_t m = __ballot_sync(0xffffffffu, invalid);
    const uint32_t n_ok = m ? (uint32_t)__ffs((int)m) - 1u : 32u;
    const uint32_t end = rel + e + L + (uint32_t)(t + 1);         /* where this lane's code word ends, relative to pos */
    bits = n_ok ? __shfl_sync(0xffffffffu, end, (int)(n_ok - 1u)) : 0u;
    return n_ok;
}

/* residual of a code word that lies inside its first 32 bits (lz <= 31 - k2): linne_coder.c:150-169 + the sign fold */
__device__ __forceinline__ int32_t lnb_rw_value(uint32_t hi, uint32_t k2)
{
    const uint32_t lz = lnb_clz32(hi);
    const uint32_t ml = (lz > 1u) ? lz : 1u;
    const uint32_t low = (hi >> ((31u - k2 - ml) & 31u)) & ((1u << k2) - 1u);
    const uint32_t mult = lz ? lz + 1u : ((hi >> 30) & 1u);
    return lnb_zz_dec((mult << k2) + low);
}

#endif /* __CUDACC__ */
