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
 *   lnb_tp_entropy_kernel   lane = block.  Side information and the recursive-Rice residuals of all channels through a
 *                           64-bit register window per lane (every lane streams its own payload through L1, next
 *                           line prefetched); four residuals per lane leave as one 16-byte store.
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
 * entropy: one LANE per block, 32 blocks per warp, one code word (or one partition / channel header) per lane
 * and pass of a single flat loop.
 *
 * The walk over a block's code words is serial by format; with thousands of blocks in a batch it can stay serial
 * as long as a lane's chain is short and the 32 lanes of a warp never wait for each other.  Round 1 spent eight
 * lanes per block on speculative code-word starts (~80 warp instructions per round of ~3 code words, 3.65 ms for
 * a 1-hour stream).  Here every lane keeps a 64-bit window of ITS payload in registers:
 *     lz = clz(hi);  length = k2 + 1 + max(lz, 1);  window <<= length        (the whole dependent chain)
 * and the value comes out of the same 32 bits.  Partition headers (gamma-coded delta of k2) and channel headers
 * (partition order + first k2) are just other kinds of pass, executed under one vote only when some lane is at
 * one, so lanes with different partition orders do not serialise each other.  The next payload word is fetched
 * one merge ahead through L1 (a lane streams 128-byte lines of its own payload; the line after the current one is
 * prefetched), four residuals are collected per lane and leave as one 16-byte store.
 * ------------------------------------------------------------------------------------------------------ */
struct LnbTeChain {
    const uint32_t *words;      /* 16-byte aligned, at most 15 bytes before the block's first byte */
    uint32_t end_word;          /* words at or past this index read as zero */
    uint32_t hi, lo, cnt, nw, wi;
};
__device__ __forceinline__ uint32_t lnb_te_fetch(const LnbTeChain &r, uint32_t wi)
{
    const uint32_t *p = r.words + wi;
    /* entering a 128-byte line: ask for the next one now, it is needed ~100 code words from here */
    if ((((uintptr_t)p) & 127u) == 0u && wi + 32u < r.end_word)
        asm volatile("prefetch.global.L1 [%0];" :: "l"(p + 32));
    return (wi < r.end_word) ? lnb_bswap32(__ldg(p)) : 0u;
}
__device__ __forceinline__ void lnb_te_take(LnbTeChain &r, uint32_t len)                      /* 0 <= len <= 32 */
{
    r.hi = __funnelshift_lc(r.lo, r.hi, len);
    r.lo = __funnelshift_lc(0u, r.lo, len);
    r.cnt -= len;
    if (r.cnt < 32u) {
        r.hi |= r.nw >> r.cnt;
        r.lo = __funnelshift_r(0u, r.nw, r.cnt);               /* nw << (32 - cnt), 0 when cnt == 0 */
        r.cnt += 32u;
        r.wi++;
        r.nw = lnb_te_fetch(r, r.wi);
    }
}
__device__ __forceinline__ void lnb_te_open(LnbTeChain &r, uint32_t pos)
{
    const uint32_t wi = pos >> 5, s = pos & 31u;
    const uint32_t w0 = (wi < r.end_word) ? lnb_bswap32(__ldg(r.words + wi)) : 0u;
    const uint32_t w1 = (wi + 1u < r.end_word) ? lnb_bswap32(__ldg(r.words + wi + 1u)) : 0u;
    r.hi = __funnelshift_l(w1, w0, s);
    r.lo = w1 << s;
    r.cnt = 64u - s;
    r.wi = wi + 2u;
    r.nw = (r.wi < r.end_word) ? lnb_bswap32(__ldg(r.words + r.wi)) : 0u;
}
__device__ __forceinline__ uint32_t lnb_te_get(LnbTeChain &r, uint32_t n)                     /* 1 <= n <= 32 */
{
    const uint32_t v = r.hi >> (32u - n);
    lnb_te_take(r, n);
    return v;
}

__global__ void __launch_bounds__(32) lnb_tp_entropy_kernel(LnbDecodeBatch b)
{
    const uint32_t lane = threadIdx.x;
    const uint32_t blk_i = blockIdx.x * 32u + lane;
    const LnbStreamCfg &cfg = b.cfg;
    const uint32_t C = cfg.num_channels, n = cfg.block_size;
    LnbBlockDesc blk;
    bool active = blk_i < b.num_blocks;
    if (active) { blk = b.blocks[blk_i]; active = lnb_tp_shape_ok(b, blk); }
    if (__ballot_sync(0xffffffffu, active) == 0u) return;

    LnbTeChain r;
    r.words = (const uint32_t *)b.stream; r.end_word = 0; r.hi = r.lo = r.nw = 0; r.cnt = 64u; r.wi = 0;
    uint32_t rel_payload = 0, rel_end = 0, overrun = 0;

    /* ---- side information (linne_decoder.c:457-486) ---- */
    if (active) {
        const uintptr_t addr = (uintptr_t)(b.stream + blk.byte_off);
        const uint32_t rel0 = (uint32_t)(addr & 15u);
        r.words = (const uint32_t *)(addr & ~(uintptr_t)15);
        uint32_t end_byte = blk.byte_off + blk.byte_size;
        if (end_byte > b.stream_size || end_byte < blk.byte_off) end_byte = b.stream_size;
        rel_payload = rel0 + LNB_BLOCK_HEADER_SIZE; rel_end = rel0 + (end_byte - blk.byte_off);
        r.end_word = (rel_end + 3u) >> 2;
        lnb_te_open(r, rel_payload * 8u);
        LnbChanParams *params = b.params + (size_t)blk_i * C;
        for (uint32_t c = 0; c < C; c++)
            for (int f = 0; f < LNB_NUM_PREEM; f++) {
                params[c].preem_prev[f] = lnb_zz_dec(lnb_te_get(r, cfg.bits_per_sample + 1u));
                params[c].preem_coef[f] = (uint8_t)lnb_te_get(r, LNB_PREEM_SHIFT - 1);
            }
        for (uint32_t c = 0; c < C; c++)
            for (uint32_t l = 0; l < cfg.num_layers; l++) {
                const uint32_t P = cfg.layer_params[l];
                params[c].log2_units[l] = (uint8_t)lnb_te_get(r, 3);
                params[c].rshift[l] = (uint8_t)lnb_te_get(r, 4);
                uint32_t *q = (uint32_t *)(params[c].coef + l * LNB_MAX_PARAMS);     /* 4 taps per store */
                uint32_t four = 0;
                for (uint32_t i = 0; i < P; i++) {
                    const uint32_t e = b.tab.huff_lut[r.hi >> (32 - LNB_HUFF_LUT_BITS)];
                    lnb_te_take(r, e & 15u);
                    four |= ((uint32_t)lnb_zz_dec(e >> 4) & 0xFFu) << (8u * (i & 3u));
                    if ((i & 3u) == 3u || i + 1u == P) { q[i >> 2] = four; four = 0; }
                }
            }
    }

    /* ---- residuals: one flat loop, one code word or header per lane and pass ---- */
    uint32_t chan = 0, left = 0, parts_left = 0, len = 0, k2 = 0, idx = 0;
    uint32_t sh = 31u, k2p1 = 1u, k2mask = 0u, fin = 0u;
    int32_t a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    bool running = active;
    int32_t *wp = b.pcm;
    while (__any_sync(0xffffffffu, running)) {
        const uint32_t h = r.hi;
        const uint32_t lz = lnb_clz32(h);
        const bool is_rice = left != 0u;
        /* the Rice view of the window (every lane computes it; lanes at a header ignore it) */
        const uint32_t ml = (lz > 1u) ? lz : 1u;
        uint32_t take = k2p1 + ml;
        if (__any_sync(0xffffffffu, running && !is_rice)) {
            if (running && !is_rice) {
                if (parts_left) {                                /* partition header: gamma code of zigzag(k2 - previous k2) */
                    const uint32_t z = lz & 15u;
                    const uint32_t v = ((h << z) >> (31u - z)) - 1u;
                    take = 2u * lz + 1u;
                    k2 = (uint32_t)((int32_t)k2 + lnb_zz_dec(v));
                    left = len; parts_left--;
                    if (lz > 15u || k2 > 30u) { overrun = 1u; k2 = 30u; left = 0u; parts_left = 0u; take = 0u; }
                } else if (chan < C && !overrun) {               /* channel header: partition order, first k2 (linne_coder.c:310-313) */
                    uint32_t porder = h >> 22;
                    k2 = (h >> 17) & 31u;
                    take = 15u;
                    if (porder > LNB_MAX_PORDER) { overrun = 1u; porder = 0u; }
                    if (k2 > 30u) { overrun = 1u; k2 = 30u; }
                    len = n >> porder; parts_left = (1u << porder) - 1u; left = len;
                    if (overrun) { left = 0u; parts_left = 0u; take = 0u; }
                    wp = b.pcm + (size_t)chan * cfg.pcm_stride + blk.smp_off;
                    idx = 0u;
                    chan++;
                } else {                                         /* done: the lane idles along (its window reads zeros past the block) */
                    running = false; take = 0u;
                    fin = r.wi * 32u - r.cnt;
                }
                sh = 31u - k2; k2p1 = k2 + 1u; k2mask = (1u << k2) - 1u;
            }
        }
        if (is_rice) {
            int32_t v;
            if (__builtin_expect(lz > sh, 0)) {
                /* code word longer than 32 bits (rare): walked here */
                uint32_t q = 0;
                while (r.hi == 0u && !overrun) {
                    q += 32u;
                    lnb_te_take(r, 32u);
                    if (r.wi > r.end_word + 2u) overrun = 1u;
                }
                if (!overrun) {
                    const uint32_t z = lnb_clz32(r.hi);
                    q += z;
                    lnb_te_take(r, z + 1u);
                }
                const uint32_t low = k2 ? (r.hi >> (32u - k2)) : 0u;
                lnb_te_take(r, k2);
                v = lnb_zz_dec((q == 0u) ? 0u : low + (2u << k2) + ((q - 1u) << k2));
                take = 0u;
                if (overrun) { left = 1u; parts_left = 0u; }
            } else {
                const uint32_t low = (h >> ((sh - ml) & 31u)) & k2mask;
                const uint32_t mult = lz ? lz + 1u : ((h >> 30) & 1u);
                v = lnb_zz_dec((mult << k2) + low);
            }
            a0 = a1; a1 = a2; a2 = a3; a3 = v;
            idx++;
            if ((idx & 3u) == 0u) *(int4 *)(wp + idx - 4u) = make_int4(a0, a1, a2, a3);
            left--;
        }
        lnb_te_take(r, take);
    }
    if (active) {
        const uint32_t used = (fin - rel_payload * 8u + 7u) >> 3;
        b.blocks[blk_i].na = used;                                               /* payload bytes consumed (reference Flush + Tell) */
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
