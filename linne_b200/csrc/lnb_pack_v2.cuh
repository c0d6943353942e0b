/* lnb_pack_v2.cuh -- cooperative block packer: one CTA per block.
 *
 * Replaces the one-thread-per-block packer (stage E9) for every block whose image fits the
 * shared-memory staging buffer.  Covers reference rows a18-a21:
 * libs/linne_encoder/src/linne_encoder.c:698-749 (side information + residual emission),
 * libs/linne_coder/src/linne_coder.c:281-302 (partition parameters + recursive Rice codes),
 * :556-585 (raw payload), :807-858 (block framing) and linne_utility.c:72-89 (CRC16).
 *
 * How
 *   - every code word's bit offset comes from an exclusive scan of code lengths (per thread: its
 *     contiguous run of samples; across threads: warp shuffles + one shared-memory pass);
 *   - code words are OR-ed into a zeroed big-endian word image in shared memory (a code of <= 32 bits
 *     touches at most two words; zero runs only advance the position);
 *   - the CRC16 is computed in parallel: 256 chunk CRCs combined by multiplying with x^(8*len) in
 *     GF(2)[x]/P (the CRC has init 0 and no final xor, so it is linear and leading zeros are free);
 *   - the finished image is copied to its final, byte-unaligned place in the stream with 128-bit
 *     stores (head/tail bytes singly).
 */
#pragma once
#include "lnb_common.cuh"
#include "lnb_encode_core.cuh"

#define LNB_PK_THREADS 256

/* exclusive scan of one value per thread; returns the thread's offset, total in `total` */
__device__ __forceinline__ uint32_t lnb_pk_scan(uint32_t v, uint32_t *warp_sums /* [9] */, uint32_t &total)
{
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, off);
        if (lane >= (uint32_t)off) inc += t;
    }
    __syncthreads();
    if (lane == 31u) warp_sums[warp] = inc;
    __syncthreads();
    uint32_t base = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < LNB_PK_THREADS / 32; w++) {
        const uint32_t s = warp_sums[w];
        if ((uint32_t)w < warp) base += s;
        tot += s;
    }
    total = tot;
    return base + inc - v;
}

/* OR `n` (<= 32) low bits of `val` into the payload at bit position `pos` (MSB-first) */
__device__ __forceinline__ void lnb_pk_emit(uint32_t *payload_words, uint32_t pos, uint32_t val, uint32_t n)
{
    if (n == 0) return;
    const uint32_t s = pos & 31u;
    const uint64_t v = ((uint64_t)val & ((n >= 32u) ? 0xFFFFFFFFull : ((1ull << n) - 1ull))) << (64u - n - s);
    const uint32_t hi = (uint32_t)(v >> 32), lo = (uint32_t)v;
    uint32_t *w = payload_words + (pos >> 5);
    if (hi) atomicOr(w, hi);
    if (lo) atomicOr(w + 1, lo);
}

/* ---- GF(2) arithmetic of the reflected CRC16 (polynomial 0xA001): bit 15 <-> x^0 ---- */
__device__ __forceinline__ uint32_t lnb_crc_mulx(uint32_t b) { return (b >> 1) ^ ((b & 1u) ? 0xA001u : 0u); }
__device__ __forceinline__ uint32_t lnb_crc_mulmod(uint32_t a, uint32_t b)
{
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) {
        if (a & (0x8000u >> i)) r ^= b;
        b = lnb_crc_mulx(b);
    }
    return r;
}
__device__ __forceinline__ uint32_t lnb_crc_xpow_bytes(uint32_t n)      /* x^(8n) */
{
    uint32_t base = 0x8000u, result = 0x8000u;
#pragma unroll
    for (int i = 0; i < 8; i++) base = lnb_crc_mulx(base);
    while (n) {
        if (n & 1u) result = lnb_crc_mulmod(result, base);
        base = lnb_crc_mulmod(base, base);
        n >>= 1;
    }
    return result;
}

/* side-information symbol j of a compressed block -> (value, bit count).
 * per channel, per layer: log2(units) 3 bits, shift 4 bits, then the layer's Huffman-coded taps */
__device__ __forceinline__ void lnb_pk_side_symbol(const LnbStreamCfg &cfg, const LnbDevTables &tab,
                                                   const LnbChanParams *params, uint32_t per_channel,
                                                   uint32_t j, uint32_t &val, uint32_t &nbits)
{
    const uint32_t c = j / per_channel;
    uint32_t r = j % per_channel, l = 0;
    while (r >= cfg.layer_params[l] + 2u) { r -= cfg.layer_params[l] + 2u; l++; }
    const LnbChanParams &p = params[c];
    if (r == 0) { val = p.log2_units[l]; nbits = 3; }
    else if (r == 1) { val = p.rshift[l]; nbits = 4; }
    else {
        const uint32_t sym = lnb_zz_enc(p.coef[l * LNB_MAX_PARAMS + (r - 2u)]) & 0xFFu;
        val = tab.huff_code[sym]; nbits = tab.huff_len[sym];
    }
}

__global__ void __launch_bounds__(LNB_PK_THREADS) lnb_pack_v2_kernel(LnbEncodeBatch b, uint32_t out_capacity, uint32_t img_words)
{
    extern __shared__ __align__(16) uint32_t lnb_pk_img[];          /* [img_words] block image, stream starts at byte 1 */
    __shared__ uint32_t warp_sums[LNB_PK_THREADS / 32 + 1];
    __shared__ uint16_t crc_tab[256];
    __shared__ uint32_t crc_part[LNB_PK_THREADS];

    const uint32_t tid = threadIdx.x;
    LnbBlockDesc &gblk = b.blocks[blockIdx.x];
    const LnbBlockDesc blk = gblk;
    const uint32_t C = b.cfg.num_channels, n = blk.nsmp;
    if ((uint64_t)blk.byte_off + blk.byte_size > out_capacity) return;     /* host reports INSUFFICIENT_BUFFER */
    const uint32_t payload_bytes = blk.byte_size - LNB_BLOCK_HEADER_SIZE;
    const uint32_t words_needed = 3u + (payload_bytes + 3u) / 4u + 2u;
    if (words_needed > img_words) return;                                  /* left to the flat packer */

    for (uint32_t w = tid; w < words_needed; w += LNB_PK_THREADS) lnb_pk_img[w] = 0;
    crc_tab[tid] = b.tab.crc_table[tid];
    __syncthreads();
    uint32_t *payload = lnb_pk_img + 3;
    uint8_t *bytes = (uint8_t *)lnb_pk_img;                                /* stream byte k lives at bytes[1 + k] */

    if (blk.type == LNB_BLOCK_COMPRESSED) {
        const LnbChanParams *params = b.params + (size_t)blockIdx.x * C;
        const LnbCoderPlan *plans = b.plans + (size_t)blockIdx.x * C;
        const uint32_t bps = b.cfg.bits_per_sample;
        /* (1) pre-emphasis state: fixed-size fields (linne_encoder.c:703-716) */
        for (uint32_t f = tid; f < 4u * C; f += LNB_PK_THREADS) {
            const uint32_t c = f >> 2, r = f & 3u;
            const uint32_t pos = c * 2u * (bps + 5u) + (r >> 1) * (bps + 5u) + ((r & 1u) ? bps + 1u : 0u);
            if (r & 1u) lnb_pk_emit(payload, pos, params[c].preem_coef[r >> 1], LNB_PREEM_SHIFT - 1);
            else lnb_pk_emit(payload, pos, lnb_zz_enc(params[c].preem_prev[r >> 1]), bps + 1u);
        }
        uint32_t bit_base = C * 2u * (bps + 5u);
        /* (2) units / shift / Huffman-coded taps (linne_encoder.c:718-735) */
        {
            uint32_t per_channel = 0;
            for (uint32_t l = 0; l < b.cfg.num_layers; l++) per_channel += b.cfg.layer_params[l] + 2u;
            const uint32_t nsym = per_channel * C;
            const uint32_t K = (nsym + LNB_PK_THREADS - 1u) / LNB_PK_THREADS;
            const uint32_t j0 = tid * K, j1 = (j0 + K < nsym) ? j0 + K : nsym;
            uint32_t mine = 0, val, nb, total;
            for (uint32_t j = j0; j < j1; j++) { lnb_pk_side_symbol(b.cfg, b.tab, params, per_channel, j, val, nb); mine += nb; }
            uint32_t pos = bit_base + lnb_pk_scan(mine, warp_sums, total);
            for (uint32_t j = j0; j < j1; j++) {
                lnb_pk_side_symbol(b.cfg, b.tab, params, per_channel, j, val, nb);
                lnb_pk_emit(payload, pos, val, nb);
                pos += nb;
            }
            bit_base += total;
        }
        /* (3) residuals, channel after channel (linne_coder.c:281-302) */
        const uint32_t T = (n + LNB_PK_THREADS - 1u) / LNB_PK_THREADS;
        for (uint32_t c = 0; c < C; c++) {
            const LnbCoderPlan &pl = plans[c];
            const int32_t *r = b.work + ((size_t)blockIdx.x * C + c) * b.cfg.work_stride;
            const uint32_t plen = n >> pl.porder;                      /* samples per partition */
            const uint32_t i0 = tid * T, i1 = (i0 + T < n) ? i0 + T : n;
            uint32_t mine = 0, total;
            for (uint32_t i = i0; i < i1; i++) {
                const uint32_t part = i / plen, k2 = pl.k2[part];
                if (i == part * plen) {
                    if (part == 0) mine += 10u + 5u;
                    else mine += lnb_gamma_bits(lnb_zz_enc((int32_t)k2 - (int32_t)pl.k2[part - 1u]));
                }
                mine += lnb_rice_len(k2, lnb_zz_enc(r[i]));
            }
            uint32_t pos = bit_base + lnb_pk_scan(mine, warp_sums, total);
            for (uint32_t i = i0; i < i1; i++) {
                const uint32_t part = i / plen, k2 = pl.k2[part], k1 = k2 + 1u;
                if (i == part * plen) {
                    if (part == 0) {
                        lnb_pk_emit(payload, pos, pl.porder, 10); pos += 10u;
                        lnb_pk_emit(payload, pos, k2, 5); pos += 5u;
                    } else {
                        const uint32_t v = lnb_zz_enc((int32_t)k2 - (int32_t)pl.k2[part - 1u]);
                        if (v == 0) { lnb_pk_emit(payload, pos, 1, 1); pos += 1u; }
                        else {
                            const uint32_t nd = lnb_log2_ceil(v + 2u);
                            lnb_pk_emit(payload, pos + nd - 1u, v + 1u, nd); pos += 2u * nd - 1u;
                        }
                    }
                }
                uint32_t uv = lnb_zz_enc(r[i]);
                if (uv < (1u << k1)) {
                    lnb_pk_emit(payload, pos, (1u << k1) | uv, k1 + 1u); pos += k1 + 1u;
                } else {
                    uv -= (1u << k1);
                    const uint32_t q = 1u + (uv >> k2);                /* zeros before the terminating one */
                    lnb_pk_emit(payload, pos + q, (1u << k2) | (uv & ((1u << k2) - 1u)), k2 + 1u);
                    pos += q + 1u + k2;
                }
            }
            bit_base += total;
        }
        __syncthreads();
        /* big-endian words -> stream byte order */
        for (uint32_t w = tid; w < (payload_bytes + 3u) / 4u; w += LNB_PK_THREADS) payload[w] = lnb_bswap32(payload[w]);
    } else if (blk.type == LNB_BLOCK_RAW) {
        const uint32_t nb = b.cfg.bits_per_sample >> 3;
        for (uint32_t j = tid; j < n * C; j += LNB_PK_THREADS) {
            const uint32_t i = j / C, c = j % C;
            const uint32_t v = lnb_zz_enc(b.pcm[(size_t)c * b.cfg.pcm_stride + blk.smp_off + i]);
            uint8_t *dst = bytes + 12u + (size_t)j * nb;
            for (uint32_t k = 0; k < nb; k++) dst[k] = (uint8_t)(v >> (8u * (nb - 1u - k)));
        }
    }
    __syncthreads();

    /* block header (linne_encoder.c:807-819, :848): stream bytes 0..10 = buffer bytes 1..11 */
    if (tid == 0) {
        bytes[1] = 0xFF; bytes[2] = 0xFF;
        const uint32_t sz = payload_bytes + 5u;
        bytes[3] = (uint8_t)(sz >> 24); bytes[4] = (uint8_t)(sz >> 16); bytes[5] = (uint8_t)(sz >> 8); bytes[6] = (uint8_t)sz;
        bytes[9] = (uint8_t)blk.type;
        bytes[10] = (uint8_t)(n >> 8); bytes[11] = (uint8_t)n;
    }
    __syncthreads();

    /* CRC16 over stream bytes [8, 11 + payload) = buffer bytes [9, 12 + payload) */
    {
        const uint32_t first = 9u, end = 12u + payload_bytes, L = end - first;
        const uint32_t Lc = (L + LNB_PK_THREADS - 1u) / LNB_PK_THREADS;
        const int32_t start = (int32_t)end - (int32_t)(Lc * LNB_PK_THREADS);   /* virtual leading zeros before `first` */
        int32_t lo = start + (int32_t)(tid * Lc), hi = lo + (int32_t)Lc;
        if (lo < (int32_t)first) lo = (int32_t)first;
        uint32_t crc = 0;
        for (int32_t i = lo; i < hi; i++) crc = (crc >> 8) ^ crc_tab[(crc ^ bytes[i]) & 0xFFu];
        crc_part[tid] = crc;
        uint32_t pw = lnb_crc_xpow_bytes(Lc);
        for (uint32_t stride = 1; stride < LNB_PK_THREADS; stride <<= 1) {
            __syncthreads();
            uint32_t v = 0;
            const bool active = (tid % (2u * stride)) == 0u;
            if (active) v = lnb_crc_mulmod(crc_part[tid], pw) ^ crc_part[tid + stride];
            __syncthreads();
            if (active) crc_part[tid] = v;
            pw = lnb_crc_mulmod(pw, pw);
        }
        __syncthreads();
        if (tid == 0) { bytes[7] = (uint8_t)(crc_part[0] >> 8); bytes[8] = (uint8_t)crc_part[0]; }
    }
    __syncthreads();

    /* copy buffer bytes [1, 1 + byte_size) to the stream: 128-bit stores on the aligned middle */
    {
        uint8_t *dst = b.out + blk.byte_off;
        const uint32_t nbytes = blk.byte_size;
        uint32_t head = (uint32_t)((16u - ((uintptr_t)dst & 15u)) & 15u);
        if (head > nbytes) head = nbytes;
        for (uint32_t i = tid; i < head; i += LNB_PK_THREADS) dst[i] = bytes[1u + i];
        const uint32_t nvec = (nbytes - head) / 16u;
        uint4 *dst4 = (uint4 *)(dst + head);
        for (uint32_t v = tid; v < nvec; v += LNB_PK_THREADS) {
            const uint32_t a = 1u + head + 16u * v;                    /* buffer byte address */
            const uint32_t *src = lnb_pk_img + (a >> 2);
            const uint32_t sh = (a & 3u) * 8u;
            const uint32_t w0 = src[0], w1 = src[1], w2 = src[2], w3 = src[3], w4 = src[4];
            uint4 o;
            o.x = __funnelshift_r(w0, w1, sh); o.y = __funnelshift_r(w1, w2, sh);
            o.z = __funnelshift_r(w2, w3, sh); o.w = __funnelshift_r(w3, w4, sh);
            dst4[v] = o;
        }
        for (uint32_t i = head + nvec * 16u + tid; i < nbytes; i += LNB_PK_THREADS) dst[i] = bytes[1u + i];
    }
    if (tid == 0) gblk.status = blk.status | LNB_ENC_FLAG_PACKED;
}
