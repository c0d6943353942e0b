/* linne_oracle.c -- CPU restatement of the LINNE block encode/decode path (see linne_oracle.h).
 *
 * TEST INFRASTRUCTURE ONLY: the checker the CUDA path is compared with, never the thing shipped.
 * Parity status: PINNED against the unmodified reference (tests/test_oracle_vs_reference.py).
 *
 * Written from the reference's algorithm, not from its text: one flat file, stage functions in
 * the order of the GPU pipeline, English comments, and a citation (reference-relative
 * file:line) on every function.  Floating-point statements keep the reference's evaluation
 * order on purpose -- with -O3 -ffp-contract=off on x86-64 this file and the reference produce
 * byte-identical streams, which is what pins it.
 */
#include "linne_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define LO_PI 3.1415926535897932384626433832795029

/* ------------------------------------------------------------------------------------------
 * Presets and the fixed coefficient-symbol statistics.
 * libs/linne_internal/src/linne_internal.c:16-41 (layer sizes, regulariser lists, preset map)
 * and :26-28 (256-entry frequency table: format constants, needed by any implementation).
 * ------------------------------------------------------------------------------------------ */
static const lo_preset k_presets[8] = {
    { 2, { 2,  32,  0 }, 1, { 0.0, 0, 0, 0 } },
    { 2, { 2,  32,  0 }, 2, { 0.0, 1.0 / 512.0, 0, 0 } },
    { 3, { 4,  64,  8 }, 1, { 0.0, 0, 0, 0 } },
    { 3, { 4,  64,  8 }, 2, { 0.0, 1.0 / 512.0, 0, 0 } },
    { 3, { 4,  64,  8 }, 4, { 0.0, 1.0 / 2048.0, 1.0 / 512.0, 1.0 / 128.0 } },
    { 3, { 4, 128, 16 }, 1, { 0.0, 0, 0, 0 } },
    { 3, { 4, 128, 16 }, 2, { 0.0, 1.0 / 512.0, 0, 0 } },
    { 3, { 4, 128, 16 }, 4, { 0.0, 1.0 / 2048.0, 1.0 / 512.0, 1.0 / 128.0 } },
};

static const uint32_t k_coef_freq[256] = {
    2944693,2417040,2500224,2220717,2361506,2005548,2161319,1804396,1961813,1628891,1774159,1471673,
    1604885,1335449,1451476,1218111,1316402,1112581,1200154,1019661,1094294,935533,1000598,861453,
    914647,793863,837607,733372,769686,679634,709504,630828,653277,583990,602876,545068,556612,507071,
    516014,473301,478009,441389,442848,415057,412045,389010,384623,364872,359578,343600,335976,322541,
    314173,304513,293388,286871,277191,271905,260699,256892,245269,243815,231142,231894,217938,220197,
    205798,209146,196061,199652,185811,189659,176121,181265,168122,173827,159699,167156,150968,158868,
    144276,152666,137117,146329,130245,141026,124044,134984,118946,130389,113141,125287,108826,120399,
    102664,116857,98953,112210,93718,109059,89757,106036,86363,102597,82554,99558,78306,96473,76105,
    92575,72428,89227,68911,85952,66258,82764,63571,80241,61196,78050,58502,75544,56329,73454,53557,
    71750,51667,81769,52853,90325,53934,86990,51338,83565,48756,80882,47304,78156,44823,75050,43129,
    72304,41339,70163,39767,67853,37538,65134,35572,62994,34367,61059,32981,58664,31690,56196,30505,
    54354,29091,52803,27750,50577,26523,49428,25414,47359,24109,46224,23419,44925,22167,43578,21336,
    42201,20551,41434,19640,39842,18815,38775,18200,37804,17159,36516,16591,35217,16053,34221,14962,
    33101,14533,32077,13842,31550,13427,30277,12962,29616,12296,29090,11678,27922,11467,27212,10733,
    26329,10270,25938,9930,24828,9336,24672,9085,23868,8616,23456,8430,22633,7892,21759,7594,21723,
    7430,20729,6988,20475,6673,20100,6489,19480,6100,18993,5912,18480,5599,17993,5292,17267,5100,
    17013,4919,16502,4721,16304,4471,16040,4313,16120,4090,17146,3921,28239,3817,49638,5544,7587,
};

const lo_preset *lo_get_preset(uint32_t preset) { return preset < 8 ? &k_presets[preset] : NULL; }
const uint32_t *lo_coef_freq_table(void) { return k_coef_freq; }

/* ------------------------------------------------------------------------------------------
 * Small integer helpers.  libs/linne_internal/include/linne_utility.h:30-32 (zig-zag), :55 (log2 ceil)
 * ------------------------------------------------------------------------------------------ */
static uint32_t zz_enc(int32_t s) { return s < 0 ? (uint32_t)(-(s << 1) - 1) : (uint32_t)(s << 1); }
static int32_t zz_dec(uint32_t u) { return (int32_t)(u >> 1) ^ -(int32_t)(u & 1); }
static uint32_t clz32(uint32_t x) { return x ? (uint32_t)__builtin_clz(x) : 32u; }
static uint32_t log2_ceil(uint32_t x) { return 32u - clz32(x - 1u); }
static double round_half_away(double d) { return d >= 0.0 ? floor(d + 0.5) : -floor(-d + 0.5); }
static double log2_via_ln(double d) { return log(d) * 1.4426950408889634; }  /* lpc.c:55-60 */

/* ------------------------------------------------------------------------------------------
 * CRC16-IBM, reflected polynomial 0xA001, init 0, no final xor.
 * libs/linne_internal/src/linne_utility.c:72-89 (table-driven there; bitwise here, same function).
 * ------------------------------------------------------------------------------------------ */
uint16_t lo_crc16(const uint8_t *data, size_t size)
{
    uint16_t crc = 0;
    size_t i;
    int b;
    for (i = 0; i < size; i++) {
        crc ^= data[i];
        for (b = 0; b < 8; b++) crc = (uint16_t)((crc & 1) ? (crc >> 1) ^ 0xA001 : crc >> 1);
    }
    return crc;
}

/* ------------------------------------------------------------------------------------------
 * Static Huffman code from symbol counts.
 * libs/static_huffman/src/static_huffman.c:28-92 (tree: repeatedly merge the two smallest live
 * nodes, scanning in index order, strict '<' so earlier index wins ties; first minimum becomes
 * child 0) and :95-131 (codes: child 0 appends bit 0, child 1 appends bit 1).
 * ------------------------------------------------------------------------------------------ */
void lo_huffman_build(const uint32_t *counts, uint32_t n, uint32_t *codes, uint8_t *lengths)
{
    uint64_t weight[512];
    uint32_t child0[512], child1[512];
    uint32_t live = n, next = n, i;
    /* iterative code assignment stack */
    uint32_t stack_node[512], stack_code[512];
    uint8_t stack_len[512];
    int sp = 0;

    for (i = 0; i < n; i++) weight[i] = counts[i] ? counts[i] : 1;   /* zero counts become 1 (:19-23) */
    for (i = n; i < 512; i++) weight[i] = 0;

    while (live > 1) {
        uint32_t a = UINT32_MAX, b = UINT32_MAX;
        for (i = 0; i < next; i++) {
            if (weight[i] == 0) continue;
            if (a == UINT32_MAX || weight[i] < weight[a]) { b = a; a = i; }
            else if (b == UINT32_MAX || weight[i] < weight[b]) { b = i; }
        }
        /* parent weight wraps in 32 bits in the reference (uint32 add); counts here never reach it */
        weight[next] = (uint32_t)(weight[a] + weight[b]);
        weight[a] = weight[b] = 0;
        child0[next] = a; child1[next] = b;
        next++; live--;
    }
    stack_node[0] = next - 1; stack_code[0] = 0; stack_len[0] = 0; sp = 1;
    while (sp > 0) {
        uint32_t node, code; uint8_t len;
        sp--; node = stack_node[sp]; code = stack_code[sp]; len = stack_len[sp];
        if (node < n) { codes[node] = code; lengths[node] = len; continue; }
        stack_node[sp] = child1[node]; stack_code[sp] = (code << 1) | 1u; stack_len[sp] = (uint8_t)(len + 1); sp++;
        stack_node[sp] = child0[node]; stack_code[sp] = (code << 1);      stack_len[sp] = (uint8_t)(len + 1); sp++;
    }
}

void lo_coef_huffman_table(uint32_t codes[256], uint8_t lengths[256])
{
    lo_huffman_build(k_coef_freq, 256, codes, lengths);
}

/* ------------------------------------------------------------------------------------------
 * MSB-first bit writer / reader over a byte buffer.
 * Bit order and flush-to-byte semantics of libs/bit_stream/include/bit_stream.h:240-302 (put,
 * zero run), :305-394 (get, zero-run get), :397-434 (flush).  The reference keeps a 32-bit
 * accumulator and spills big-endian words; the byte image that results is simply "first bit
 * written is bit 7 of byte 0", which is what is implemented here.
 * ------------------------------------------------------------------------------------------ */
typedef struct { uint8_t *buf; uint64_t cap_bits; uint64_t pos; int overflow; } bitw;

static void bw_init(bitw *w, uint8_t *buf, uint32_t cap_bytes)
{
    w->buf = buf; w->cap_bits = (uint64_t)cap_bytes * 8u; w->pos = 0; w->overflow = 0;
    memset(buf, 0, cap_bytes);
}
static void bw_put(bitw *w, uint32_t val, uint32_t nbits)   /* low nbits of val, MSB first */
{
    uint32_t i;
    if (nbits == 0) return;
    if (w->pos + nbits > w->cap_bits) { w->overflow = 1; w->pos += nbits; return; }
    for (i = 0; i < nbits; i++) {
        uint32_t bit = (val >> (nbits - 1 - i)) & 1u;
        if (bit) w->buf[w->pos >> 3] |= (uint8_t)(0x80u >> (w->pos & 7));
        w->pos++;
    }
}
static void bw_zero_run(bitw *w, uint64_t run)              /* `run` zeros then a one */
{
    if (w->pos + run + 1 > w->cap_bits) { w->overflow = 1; w->pos += run + 1; return; }
    w->pos += run;
    bw_put(w, 1, 1);
}
static uint32_t bw_bytes(const bitw *w) { return (uint32_t)((w->pos + 7) >> 3); }

typedef struct { const uint8_t *buf; uint64_t size_bits; uint64_t pos; } bitr;

static void br_init(bitr *r, const uint8_t *buf, uint32_t size_bytes)
{
    r->buf = buf; r->size_bits = (uint64_t)size_bytes * 8u; r->pos = 0;
}
static uint32_t br_bit(bitr *r)
{
    uint32_t bit = 0;
    if (r->pos < r->size_bits) bit = (r->buf[r->pos >> 3] >> (7 - (r->pos & 7))) & 1u;
    r->pos++;
    return bit;
}
static uint32_t br_get(bitr *r, uint32_t nbits)
{
    uint32_t v = 0, i;
    for (i = 0; i < nbits; i++) v = (v << 1) | br_bit(r);
    return v;
}
static uint32_t br_zero_run(bitr *r)                        /* count zeros, swallow the one */
{
    uint32_t run = 0;
    while (r->pos < r->size_bits) {
        if (br_bit(r)) return run;
        run++;
    }
    return run;
}

/* ------------------------------------------------------------------------------------------
 * Header.  libs/linne_encoder/src/linne_encoder.c:53-138, libs/linne_decoder/src/linne_decoder.c:60-131
 * ------------------------------------------------------------------------------------------ */
static void put_be(uint8_t **p, uint32_t v, int nbytes)
{
    int i;
    for (i = nbytes - 1; i >= 0; i--) *(*p)++ = (uint8_t)(v >> (8 * i));
}
static uint32_t get_be(const uint8_t **p, int nbytes)
{
    uint32_t v = 0; int i;
    for (i = 0; i < nbytes; i++) v = (v << 8) | *(*p)++;
    return v;
}

int lo_header_encode(const lo_stream_info *h, uint8_t *out, uint32_t cap)
{
    uint8_t *p = out;
    if (!h || !out) return LO_INVALID_ARGUMENT;
    if (cap < LO_HEADER_SIZE) return LO_INSUFFICIENT_BUFFER;
    if (!h->num_channels || !h->num_samples || !h->sampling_rate || !h->bits_per_sample
        || !h->block_size || h->preset >= 8 || h->ms >= 2 || (h->ms == 1 && h->num_channels == 1))
        return LO_INVALID_FORMAT;
    *p++ = 'I'; *p++ = 'B'; *p++ = 'R'; *p++ = 'A';
    put_be(&p, 1, 4); put_be(&p, 2, 4);                    /* versions are always the macros (:112-117) */
    put_be(&p, h->num_channels, 2); put_be(&p, h->num_samples, 4); put_be(&p, h->sampling_rate, 4);
    put_be(&p, h->bits_per_sample, 2); put_be(&p, h->block_size, 4);
    put_be(&p, h->preset, 1); put_be(&p, h->ms, 1);
    return LO_OK;
}

int lo_header_decode(const uint8_t *data, uint32_t size, lo_stream_info *h)
{
    const uint8_t *p = data;
    if (!data || !h) return LO_INVALID_ARGUMENT;
    if (size < LO_HEADER_SIZE) return LO_INSUFFICIENT_DATA;
    if (p[0] != 'I' || p[1] != 'B' || p[2] != 'R' || p[3] != 'A') return LO_INVALID_FORMAT;
    p += 4;
    h->format_version = get_be(&p, 4); h->codec_version = get_be(&p, 4);
    h->num_channels = get_be(&p, 2); h->num_samples = get_be(&p, 4); h->sampling_rate = get_be(&p, 4);
    h->bits_per_sample = get_be(&p, 2); h->block_size = get_be(&p, 4);
    h->preset = get_be(&p, 1); h->ms = get_be(&p, 1);
    return LO_OK;
}

/* linne_decoder.c:134-184 */
static int header_valid(const lo_stream_info *h)
{
    return h->format_version == 1 && h->codec_version == 2 && h->num_channels && h->num_samples
        && h->sampling_rate && h->bits_per_sample && h->block_size && h->preset < 8 && h->ms < 2
        && !(h->ms == 1 && h->num_channels == 1);
}

/* ------------------------------------------------------------------------------------------
 * Mid/side.  libs/linne_internal/src/linne_utility.c:120-132 (forward), :135-147 (inverse)
 * ------------------------------------------------------------------------------------------ */
void lo_ms_forward(int32_t *ch0, int32_t *ch1, uint32_t n)
{
    uint32_t i;
    for (i = 0; i < n; i++) { ch1[i] -= ch0[i]; ch0[i] += ch1[i] >> 1; }
}
void lo_ms_inverse(int32_t *ch0, int32_t *ch1, uint32_t n)
{
    uint32_t i;
    for (i = 0; i < n; i++) { ch0[i] -= ch1[i] >> 1; ch1[i] += ch0[i]; }
}

/* ------------------------------------------------------------------------------------------
 * Pre-/de-emphasis.  linne_utility.c:158-193 (coefficient), :196-212 (filter), :215-241 (inverse x2)
 * ------------------------------------------------------------------------------------------ */
int32_t lo_preemphasis_coef(const int32_t *x, uint32_t n)
{
    double c0 = 0.0, c1 = 0.0, cur;
    uint32_t i;
    int32_t coef;
    cur = x[0];
    for (i = 0; i + 1 < n; i++) {
        const double nxt = x[i + 1];
        c0 += cur * cur;
        c1 += cur * nxt;
        cur = nxt;
    }
    c1 /= c0;
    if (c0 < 1e-6 || c1 < 0.0) return 0;
    coef = (int32_t)round_half_away(c1 * 32.0);
    return coef >= 16 ? 15 : coef;
}

void lo_preemphasis(int32_t *x, uint32_t n, int32_t prev, int32_t coef)
{
    uint32_t i;
    for (i = 0; i < n; i++) {
        const int32_t cur = x[i];
        x[i] -= (prev * coef) >> 5;
        prev = cur;
    }
}

void lo_deemphasis2(int32_t *x, uint32_t n, const int32_t prev[2], const int32_t coef[2])
{
    /* Undo filter 1 (applied last by the encoder) then filter 0, interleaved one sample apart. */
    uint32_t i;
    const int32_t c0 = coef[0], c1 = coef[1];
    x[0] += (prev[1] * c1) >> 5;
    if (n >= 2) x[1] += (x[0] * c1) >> 5;
    x[0] += (prev[0] * c0) >> 5;
    for (i = 2; i < n; i++) {
        x[i] += (x[i - 1] * c1) >> 5;
        x[i - 1] += (x[i - 2] * c0) >> 5;
    }
    if (n >= 2) x[n - 1] += (x[n - 2] * c0) >> 5;
}

/* ------------------------------------------------------------------------------------------
 * Integer predictor / synthesiser with `num_units` independent sub-blocks.
 * libs/linne_encoder/src/linne_lpc_predict.c:7-38, libs/linne_decoder/src/linne_lpc_synthesize.c:8-83.
 * int32 arithmetic wraps (computed in uint32 here to keep that defined in C).
 * A unit shorter than its tap count predicts nothing (the reference underflows there).
 * ------------------------------------------------------------------------------------------ */
void lo_predict(const int32_t *in, uint32_t n, const int32_t *coef, uint32_t num_params,
                uint32_t rshift, uint32_t num_units, int32_t *residual)
{
    const uint32_t p = num_params / num_units, m = n / num_units;
    const uint32_t half = rshift ? (1u << (rshift - 1)) : 0u;
    uint32_t u, t, k;
    memcpy(residual, in, sizeof(int32_t) * n);
    if (m <= p) return;
    for (u = 0; u < num_units; u++) {
        const int32_t *x = in + u * m;
        const int32_t *c = coef + u * p;
        int32_t *r = residual + u * m;
        for (t = 0; t < m - p; t++) {
            uint32_t acc = half;
            for (k = 0; k < p; k++) acc += (uint32_t)c[k] * (uint32_t)x[t + k];
            r[t + p] = (int32_t)((uint32_t)r[t + p] + (uint32_t)((int32_t)acc >> rshift));
        }
    }
}

void lo_synthesize(int32_t *data, uint32_t n, const int32_t *coef, uint32_t num_params,
                   uint32_t rshift, uint32_t num_units)
{
    const uint32_t p = num_params / num_units, m = n / num_units;
    const uint32_t half = rshift ? (1u << (rshift - 1)) : 0u;
    uint32_t u, t, k;
    if (m <= p) return;
    for (u = 0; u < num_units; u++) {
        int32_t *x = data + u * m;
        const int32_t *c = coef + u * p;
        for (t = 0; t < m - p; t++) {
            uint32_t acc = half;
            for (k = 0; k < p; k++) acc += (uint32_t)c[k] * (uint32_t)x[t + k];
            x[t + p] = (int32_t)((uint32_t)x[t + p] - (uint32_t)((int32_t)acc >> rshift));
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * Coefficient quantiser: one shared shift per layer, error feedback from the last tap to the first.
 * libs/lpc/src/lpc.c:981-1040 (called with 8 bits from linne_encoder.c:680-682)
 * ------------------------------------------------------------------------------------------ */
void lo_quantize(const double *coef, uint32_t n, int32_t *q, uint32_t *rshift)
{
    double peak = 0.0, carry = 0.0;
    int exponent, i;
    uint32_t shift;
    for (i = 0; i < (int)n; i++) if (peak < fabs(coef[i])) peak = fabs(coef[i]);
    if (peak <= pow(2.0, -7)) {
        *rshift = 8;
        memset(q, 0, sizeof(int32_t) * n);
        return;
    }
    (void)frexp(peak, &exponent);
    shift = (uint32_t)(7 - exponent);
    for (i = (int)n - 1; i >= 0; i--) {
        int32_t v;
        carry += coef[i] * pow(2.0, shift);
        v = (int32_t)round_half_away(carry);
        if (v >= 128) v = 127; else if (v < -128) v = -128;
        carry -= v;
        q[i] = v;
    }
    *rshift = shift;
}

/* ------------------------------------------------------------------------------------------
 * LPC analysis state.  Mirrors what the reference keeps inside struct LPCCalculator
 * (libs/lpc/src/lpc.c:31-46): the window buffer and the PARCOR array persist between calls,
 * which is observable (stale parcor[P0]; Q2: un-rewritten centre sample of an
 * odd-length Welch window), so they persist here too.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    double *win;                         /* windowed samples, capacity = max block */
    uint32_t cap;
    double acorr[LO_MAX_PARAMS + 2];
    double lpc[LO_MAX_PARAMS + 2];
    double parcor[LO_MAX_PARAMS + 2];
    double a[LO_MAX_PARAMS + 2], u[LO_MAX_PARAMS + 2], v[LO_MAX_PARAMS + 2];
    double gram[LO_MAX_PARAMS + 1][LO_MAX_PARAMS + 1];
} lpc_state;

enum { WIN_SIN = 1, WIN_WELCH = 2 };

/* lpc.c:176-212 */
static void apply_window(int kind, const double *x, uint32_t n, double *out)
{
    uint32_t i;
    if (kind == WIN_SIN) {
        for (i = 0; i < n; i++) out[i] = x[i] * sin((LO_PI * i) / (n - 1));
    } else {
        const double scale = 4.0 * pow(n - 1, -2.0);
        for (i = 0; i < (n >> 1); i++) {
            const double w = scale * i * (n - 1 - i);
            out[i] = x[i] * w;
            out[n - i - 1] = x[n - i - 1] * w;
        }
    }
}

/* lpc.c:215-249: r[lag] = sum_i d[i] d[i+lag], every lag accumulated in ascending i.
 * Lags that do not fit in the data stay 0 (the reference underflows there). */
static void autocorrelation(const double *d, uint32_t n, double *r, uint32_t num_lags)
{
    uint32_t i, lag;
    for (lag = 0; lag < num_lags; lag++) r[lag] = 0.0;
    for (i = 0; i < n; i++) {
        const double s = d[i];
        const uint32_t lim = (n - i < num_lags) ? n - i : num_lags;
        for (lag = 0; lag < lim; lag++) r[lag] += s * d[i + lag];
    }
}

/* lpc.c:252-324 */
static void levinson(lpc_state *st, const double *r, uint32_t order)
{
    double *a = st->a, *u = st->u, *v = st->v;
    double err, gamma;
    uint32_t i, k;
    if (fabs(r[0]) < FLT_EPSILON) {
        for (i = 0; i < order + 1; i++) st->lpc[i] = st->parcor[i] = 0.0;
        return;
    }
    for (i = 0; i < order + 2; i++) a[i] = u[i] = v[i] = 0.0;
    a[0] = 1.0;
    err = r[0];
    a[1] = -r[1] / r[0];
    st->parcor[0] = r[1] / err;
    err += r[1] * a[1];
    u[0] = 1.0; u[1] = 0.0;
    v[0] = 0.0; v[1] = 1.0;
    for (k = 1; k < order; k++) {
        gamma = 0.0;
        for (i = 0; i < k + 1; i++) gamma += a[i] * r[k + 1 - i];
        gamma /= -err;
        err *= (1.0 - gamma * gamma);
        for (i = 0; i < k; i++) u[i + 1] = v[k - i] = a[i + 1];
        u[0] = 1.0; u[k + 1] = 0.0;
        v[0] = 0.0; v[k + 1] = 1.0;
        for (i = 0; i < k + 2; i++) a[i] = u[i] + gamma * v[i];
        st->parcor[k] = -gamma;
    }
    memcpy(st->lpc, &a[1], sizeof(double) * order);
}

/* lpc.c:327-366 */
static void lpc_levinson_path(lpc_state *st, const double *x, uint32_t n, uint32_t order,
                              int window, double lambda)
{
    uint32_t i;
    apply_window(window, x, n, st->win);
    autocorrelation(st->win, n, st->acorr, order + 1);
    if (n < order) {
        for (i = 0; i < order + 1; i++) st->lpc[i] = st->parcor[i] = 0.0;
        return;
    }
    st->acorr[0] *= (1.0 + lambda);
    levinson(st, st->acorr, order);
}

/* lpc.c:402-448: solve gram * x = b by Cholesky; returns 0 when a pivot is not positive */
static int cholesky_solve(lpc_state *st, int dim, double *x, const double *b, double *inv_diag)
{
    int i, j, k;
    double s;
    for (i = 0; i < dim; i++) {
        s = st->gram[i][i];
        for (k = i - 1; k >= 0; k--) s -= st->gram[i][k] * st->gram[i][k];
        if (s <= 0.0) return 0;
        inv_diag[i] = pow(s, -0.5);
        for (j = i + 1; j < dim; j++) {
            s = st->gram[i][j];
            for (k = i - 1; k >= 0; k--) s -= st->gram[i][k] * st->gram[j][k];
            st->gram[j][i] = s * inv_diag[i];
        }
    }
    for (i = 0; i < dim; i++) {
        s = b[i];
        for (j = i - 1; j >= 0; j--) s -= st->gram[i][j] * x[j];
        x[i] = s * inv_diag[i];
    }
    for (i = dim - 1; i >= 0; i--) {
        s = x[i];
        for (j = i + 1; j < dim; j++) s -= st->gram[j][i] * x[j];
        x[i] = s * inv_diag[i];
    }
    return 1;
}

/* lpc.c:452-509: IRLS weights 1/max(|res|,1e-6); weighted Gram matrix and right-hand side */
static double irls_system(lpc_state *st, const double *x, uint32_t n, const double *a,
                          double *rhs, uint32_t order)
{
    uint32_t t, i, j;
    double objective = 0.0;
    for (i = 0; i < order; i++) {
        rhs[i] = 0.0;
        for (j = 0; j < order; j++) st->gram[i][j] = 0.0;
    }
    for (t = order; t < n; t++) {
        double res = x[t], w;
        for (i = 0; i < order; i++) res += a[i] * x[t - i - 1];
        res = fabs(res);
        objective += res;
        res = (res < 1e-6) ? 1e-6 : res;
        w = 1.0 / res;
        for (i = 0; i < order; i++) {
            rhs[i] -= x[t] * x[t - i - 1] * w;
            for (j = i; j < order; j++) st->gram[i][j] += x[t - i - 1] * x[t - j - 1] * w;
        }
    }
    for (i = 0; i < order; i++)
        for (j = i + 1; j < order; j++) st->gram[j][i] = st->gram[i][j];
    return objective / (n - order);
}

/* lpc.c:578-661: Levinson start + optional IRLS refinement; result in out[0..order) (a_1..a_p) */
static void lpc_unit(lpc_state *st, const double *x, uint32_t n, uint32_t order,
                     uint32_t irls_iterations, double lambda, double *out)
{
    double *a = st->a, *rhs = st->u;
    double objective, prev_objective = FLT_MAX;
    uint32_t it, i;
    lpc_levinson_path(st, x, n, order, WIN_WELCH, lambda);
    memcpy(a, st->lpc, sizeof(double) * order);
    if (fabs(st->acorr[0]) < FLT_EPSILON) {
        for (i = 0; i < order + 1; i++) st->lpc[i] = 0.0;
        memmove(out, st->lpc, sizeof(double) * order);
        return;
    }
    for (it = 0; it < irls_iterations; it++) {
        objective = irls_system(st, x, n, a, rhs, order);
        if (!cholesky_solve(st, (int)order, a, rhs, st->v)) {
            for (i = 0; i < order; i++) st->lpc[i] = 0.0;
            memmove(out, st->lpc, sizeof(double) * order);
            return;
        }
        if (fabs(prev_objective - objective) < 1e-8) break;
        prev_objective = objective;
    }
    memmove(st->lpc, a, sizeof(double) * order);
    memmove(out, st->lpc, sizeof(double) * order);
}

/* lpc.c:810-865 via linne_network.c:680-696: entropy estimate used for the raw/compressed decision */
static double estimate_bits_per_sample(lpc_state *st, const double *x, uint32_t n,
                                       uint32_t bits, uint32_t order)
{
    double power, ratio = 0.0, est;
    uint32_t k;
    lpc_levinson_path(st, x, n, order, WIN_SIN, 0.0);
    power = st->acorr[0];
    power *= pow(2, (double)(2.0 * (bits - 1)));
    if (fabs(power) <= FLT_MIN) return 0.0;
    power = log2_via_ln(power) - log2_via_ln((double)n);
    for (k = 1; k <= order; k++) ratio += log2_via_ln(1.0 - st->parcor[k] * st->parcor[k]);
    est = 1.9426950408889634 + 0.5f * (power + ratio);
    return est <= 0 ? 1.0 : est;
}

/* ------------------------------------------------------------------------------------------
 * Encoder object
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    uint32_t num_params, num_units;
    double params[LO_MAX_PARAMS];
    double dparams[LO_MAX_PARAMS];
    double *din, *dout;
} net_layer;

struct lo_encoder {
    uint32_t max_channels, max_block;
    int configured;
    lo_stream_info info;
    uint32_t enable_learning, af_iterations;
    const lo_preset *preset;
    uint32_t huff_code[256]; uint8_t huff_len[256];
    lpc_state lpc;
    net_layer layers[LO_MAX_LAYERS];
    double *net_buf;                       /* working signal of the layer cascade */
    double *fbuf;                          /* normalised input */
    int32_t *work[LO_MAX_CHANNELS];        /* integer signal per channel */
    int32_t *resid[LO_MAX_CHANNELS];
    lo_channel_trace trace[LO_MAX_CHANNELS];
    int last_block_type;
    double momentum[LO_MAX_LAYERS][LO_MAX_PARAMS];
};

lo_encoder *lo_encoder_create(uint32_t max_channels, uint32_t max_block)
{
    lo_encoder *e;
    uint32_t c, l;
    if (!max_channels || max_channels > LO_MAX_CHANNELS || !max_block) return NULL;
    e = (lo_encoder *)calloc(1, sizeof(*e));
    if (!e) return NULL;
    e->max_channels = max_channels; e->max_block = max_block;
    e->lpc.win = (double *)calloc(max_block, sizeof(double)); e->lpc.cap = max_block;
    e->net_buf = (double *)calloc(max_block, sizeof(double));
    e->fbuf = (double *)calloc(max_block, sizeof(double));
    for (l = 0; l < LO_MAX_LAYERS; l++) {
        e->layers[l].din = (double *)calloc(max_block, sizeof(double));
        e->layers[l].dout = (double *)calloc(max_block, sizeof(double));
    }
    for (c = 0; c < max_channels; c++) {
        e->work[c] = (int32_t *)calloc(max_block, sizeof(int32_t));
        e->resid[c] = (int32_t *)calloc(max_block, sizeof(int32_t));
    }
    lo_coef_huffman_table(e->huff_code, e->huff_len);
    return e;
}

void lo_encoder_destroy(lo_encoder *e)
{
    uint32_t c, l;
    if (!e) return;
    free(e->lpc.win); free(e->net_buf); free(e->fbuf);
    for (l = 0; l < LO_MAX_LAYERS; l++) { free(e->layers[l].din); free(e->layers[l].dout); }
    for (c = 0; c < e->max_channels; c++) { free(e->work[c]); free(e->resid[c]); }
    free(e);
}

/* linne_encoder.c:141-198 (validation) + :410-477 */
int lo_encoder_configure(lo_encoder *e, uint32_t num_channels, uint32_t bits, uint32_t rate,
                         uint32_t block_size, uint32_t preset, uint32_t ms,
                         uint32_t enable_learning, uint32_t af_iterations)
{
    const lo_preset *p;
    int l;
    if (!e) return LO_INVALID_ARGUMENT;
    if (!num_channels || !bits || !rate || !block_size || preset >= 8 || ms >= 2) return LO_INVALID_FORMAT;
    p = &k_presets[preset];
    for (l = 0; l < p->num_layers; l++) if (block_size <= (uint32_t)p->layer_params[l]) return LO_INVALID_FORMAT;
    if (block_size > e->max_block || num_channels > e->max_channels) return LO_INSUFFICIENT_BUFFER;
    memset(&e->info, 0, sizeof(e->info));
    e->info.format_version = 1; e->info.codec_version = 2;
    e->info.num_channels = num_channels; e->info.bits_per_sample = bits; e->info.sampling_rate = rate;
    e->info.block_size = block_size; e->info.preset = preset; e->info.ms = ms;
    e->preset = p;
    for (l = 0; l < p->num_layers; l++) {
        e->layers[l].num_params = (uint32_t)p->layer_params[l];
        e->layers[l].num_units = 1;
        memset(e->layers[l].params, 0, sizeof(e->layers[l].params));
    }
    e->enable_learning = enable_learning; e->af_iterations = af_iterations;
    e->configured = 1;
    return LO_OK;
}

const lo_channel_trace *lo_encoder_trace(const lo_encoder *e, uint32_t ch) { return &e->trace[ch]; }
const int32_t *lo_encoder_last_residual(const lo_encoder *e, uint32_t ch) { return e->resid[ch]; }
int lo_encoder_last_block_type(const lo_encoder *e) { return e->last_block_type; }

/* ------------------------------------------------------------------------------------------
 * Layer cascade ("network") analysis.
 * ------------------------------------------------------------------------------------------ */

/* libs/linne_network/src/linne_network.c:165-210: replace data by the layer's residual (double) */
static void layer_forward(net_layer *L, double *data, uint32_t n)
{
    const uint32_t U = L->num_units, m = n / U, p = L->num_params / U;
    uint32_t u, i, j;
    memcpy(L->din, data, sizeof(double) * n);
    for (u = 0; u < U; u++) {
        const double *w = &L->params[u * p];
        const double *x = &L->din[u * m];
        double *r = &data[u * m];
        double acc;
        i = 0;
        if (u == 0) {
            /* ramp-in: history before the block is treated as zero */
            for (i = 1; i < p; i++) {
                acc = 0.0;
                for (j = 0; j < i; j++) acc += w[p - i + j] * x[j];
                r[i] += acc;
            }
        }
        for (; i < m; i++) {
            acc = 0.0;
            for (j = 0; j < p; j++) acc += w[j] * x[(int32_t)(i - p + j)];
            r[i] += acc;
        }
    }
}

/* linne_network.c:213-265 */
static void layer_backward(net_layer *L, double *data, uint32_t n)
{
    const uint32_t U = L->num_units, m = n / U, p = L->num_params / U;
    uint32_t u, i, j;
    memcpy(L->dout, data, sizeof(double) * n);
    for (u = 0; u < U; u++) {
        const double *x = &L->din[u * m];
        const double *g = &L->dout[u * m];
        const double *w = &L->params[u * p];
        double *back = &data[u * m];
        double *dw = &L->dparams[u * p];
        for (i = 0; i < p; i++) {
            dw[i] = 0.0;
            for (j = 0; j < (m - p + i); j++) dw[i] += x[j] * g[p - i + j];
        }
        for (i = 0; i < (m - p); i++) {
            double s = 0.0;
            for (j = 0; j < p; j++) s += w[j] * g[p + i - j];
            back[i] += s / p;
        }
        for (; i < m; i++) {
            double s = 0.0;
            for (j = 0; j < p; j++) if ((p + i - j) < m) s += w[j] * g[p + i - j];
            back[i] += s / p;
        }
    }
}

/* linne_network.c:50-63 */
static double mean_abs(const double *d, uint32_t n)
{
    double s = 0.0;
    uint32_t i;
    for (i = 0; i < n; i++) s += fabs(d[i]);
    return s / n;
}

static void reverse_in_place(double *v, uint32_t n)
{
    uint32_t k;
    for (k = 0; k < n / 2; k++) { double t = v[k]; v[k] = v[n - k - 1]; v[n - k - 1] = t; }
}

/* linne_network.c:268-347: try U = 1,2,4,..; per unit fit LPC (Welch, Levinson only), measure the
 * L1 residual with the unit-0 ramp-in rule, keep the first minimum. */
static uint32_t search_num_units(lo_encoder *e, net_layer *L, const double *x, uint32_t n,
                                 uint32_t max_units, double lambda)
{
    double best_loss = FLT_MAX;
    uint32_t best = 0, U;
    for (U = 1; U <= max_units; U <<= 1) {
        const uint32_t p = L->num_params / U, m = n / U;
        double loss = 0.0;
        uint32_t u, t, k;
        if ((L->num_params % U) != 0 || (n % U) != 0) continue;
        for (u = 0; u < U; u++) {
            const double *xs = &x[u * m];
            double *w = &L->params[u * p];
            double r;
            lpc_unit(&e->lpc, xs, m, p, 0, lambda, w);
            reverse_in_place(w, p);
            t = 0;
            if (u == 0) {
                for (t = 1; t < p; t++) {
                    r = xs[t];
                    for (k = 0; k < t; k++) r += w[p - t + k] * xs[k];
                    loss += (r > 0) ? r : -r;
                }
            }
            for (; t < m; t++) {
                r = xs[t];
                for (k = 0; k < p; k++) r += w[k] * xs[(int32_t)(t - p + k)];
                loss += (r > 0) ? r : -r;
            }
        }
        loss /= n;
        if (loss < best_loss) { best_loss = loss; best = U; }
    }
    return best ? best : 1;
}

/* linne_network.c:350-376 */
static void set_layer_parameters(lo_encoder *e, net_layer *L, const double *x, uint32_t n,
                                 uint32_t irls_iterations, double lambda)
{
    const uint32_t U = L->num_units, p = L->num_params / U, m = n / U;
    uint32_t u;
    for (u = 0; u < U; u++) {
        lpc_unit(&e->lpc, &x[u * m], m, p, irls_iterations, lambda, &L->params[u * p]);
        reverse_in_place(&L->params[u * p], p);
    }
}

/* linne_network.c:582-602 */
static double cascade_pass(lo_encoder *e, const double *x, uint32_t n, uint32_t irls_iterations, double lambda)
{
    int l;
    memcpy(e->net_buf, x, sizeof(double) * n);
    for (l = 0; l < e->preset->num_layers; l++) {
        net_layer *L = &e->layers[l];
        const uint32_t cap = L->num_params < 128 ? L->num_params : 128;
        L->num_units = search_num_units(e, L, e->net_buf, n, cap, lambda);
        set_layer_parameters(e, L, e->net_buf, n, irls_iterations, lambda);
        layer_forward(L, e->net_buf, n);
    }
    return mean_abs(e->net_buf, n);
}

/* linne_network.c:805-873 with constants from linne_internal.h:29-33 (2000 iterations,
 * learning rate 0.1f, epsilon 1e-7) and momentum 0.8f (:832) */
static void train_cascade(lo_encoder *e, const double *x, uint32_t n)
{
    const double lr = 0.1f, alpha = 0.8f, eps = 1.0e-7;
    double loss, prev = FLT_MAX;
    uint32_t it, i;
    int l;
    const int nl = e->preset->num_layers;
    for (l = 0; l < nl; l++) for (i = 0; i < e->layers[l].num_params; i++) e->momentum[l][i] = 0.0;
    for (it = 0; it < 2000; it++) {
        memcpy(e->net_buf, x, sizeof(double) * n);
        for (l = 0; l < nl; l++) layer_forward(&e->layers[l], e->net_buf, n);
        loss = mean_abs(e->net_buf, n);
        /* linne_network.c:66-75: gradient of the mean absolute value */
        for (i = 0; i < n; i++) {
            const double d = e->net_buf[i];
            e->net_buf[i] = (double)((d > 0) - (d < 0)) / n;
        }
        for (l = nl - 1; l >= 0; l--) layer_backward(&e->layers[l], e->net_buf, n);
        for (l = 0; l < nl; l++) {
            net_layer *L = &e->layers[l];
            for (i = 0; i < L->num_params; i++) {
                e->momentum[l][i] = alpha * e->momentum[l][i] + lr * L->dparams[i];
                L->params[i] -= e->momentum[l][i];
            }
        }
        if (fabs(loss - prev) < eps) break;
        prev = loss;
    }
}

/* linne_network.c:605-630 followed (optionally) by training, linne_encoder.c:665-675 */
void lo_analyze(lo_encoder *e, const double *x, uint32_t n,
                uint32_t num_units[LO_MAX_LAYERS], double params[LO_MAX_LAYERS][LO_MAX_PARAMS])
{
    double best_loss = FLT_MAX;
    int best = 0, i, l;
    for (i = 0; i < e->preset->num_lambdas; i++) {
        const double loss = cascade_pass(e, x, n, 0, e->preset->lambdas[i]);
        if (loss < best_loss) { best_loss = loss; best = i; }
    }
    (void)cascade_pass(e, x, n, e->af_iterations, e->preset->lambdas[best]);
    if (e->enable_learning) train_cascade(e, x, n);
    for (l = 0; l < e->preset->num_layers; l++) {
        num_units[l] = e->layers[l].num_units;
        memcpy(params[l], e->layers[l].params, sizeof(double) * e->layers[l].num_params);
    }
}

/* ------------------------------------------------------------------------------------------
 * Residual coder: partitioned recursive Rice.  libs/linne_coder/src/linne_coder.c
 * ------------------------------------------------------------------------------------------ */

/* linne_coder.c:172-200 */
void lo_rice_parameter(double mean, uint32_t *k1, uint32_t *k2)
{
    const double optx = 0.5127629514437670454896078808815218508243560791015625;
    const double rho = 1.0 / (1.0 + mean);
    const double f = floor(log2_via_ln(log(optx) / log(1.0 - rho)));
    const uint32_t k = (uint32_t)((0 > f) ? 0 : f);
    *k2 = k; *k1 = k + 1;
}

/* linne_coder.c:203-214 */
uint32_t lo_rice_code_length(uint32_t k1, uint32_t k2, uint32_t uval)
{
    const uint32_t thr = 1u << k1;
    return uval < thr ? k1 + 1 : k2 + 2 + ((uval - thr) >> k2);
}

static uint32_t gamma_bits(uint32_t v) { return v == 0 ? 1u : 2u * log2_ceil(v + 2u) - 1u; }  /* :16 */

/* linne_coder.c:85-103 */
static void put_gamma(bitw *w, uint32_t v)
{
    uint32_t nd;
    if (v == 0) { bw_put(w, 1, 1); return; }
    nd = log2_ceil(v + 2u);
    bw_put(w, 0, nd - 1);
    bw_put(w, v + 1u, nd);
}
/* linne_coder.c:106-127 */
static uint32_t get_gamma(bitr *r)
{
    const uint32_t nd = br_zero_run(r) + 1u;
    if (nd == 1) return 0;
    return (uint32_t)((1ul << (nd - 1)) + br_get(r, nd - 1) - 1);
}

static double part_mean_scratch[11][1024];

/* linne_coder.c:217-278 (search half).  Means: finest level from exact integer sums, coarser
 * levels by pairwise averaging; cost in wrapping uint32; first minimum wins. */
uint32_t lo_coder_plan(const int32_t *res, uint32_t n, uint32_t *porder_out, uint32_t *k2_out)
{
    uint32_t max_porder = 1, porder, part, s, best = 0, best_bits = UINT32_MAX;
    int lvl;
    while ((n % (1u << max_porder)) == 0) max_porder++;
    max_porder = (max_porder - 1 < 10) ? max_porder - 1 : 10;
    {
        const uint32_t parts = 1u << max_porder, len = n / parts;
        for (part = 0; part < parts; part++) {
            double sum = 0.0;
            for (s = 0; s < len; s++) sum += zz_enc(res[part * len + s]);
            part_mean_scratch[max_porder][part] = sum / len;
        }
        for (lvl = (int)max_porder - 1; lvl >= 0; lvl--)
            for (part = 0; part < (1u << lvl); part++)
                part_mean_scratch[lvl][part] =
                    (part_mean_scratch[lvl + 1][2 * part] + part_mean_scratch[lvl + 1][2 * part + 1]) / 2.0;
    }
    for (porder = 0; porder <= max_porder; porder++) {
        const uint32_t len = n >> porder;
        uint32_t bits = 0, k1, k2, prev_k2 = 0;
        for (part = 0; part < (1u << porder); part++) {
            lo_rice_parameter(part_mean_scratch[porder][part], &k1, &k2);
            for (s = 0; s < len; s++) bits += lo_rice_code_length(k1, k2, zz_enc(res[part * len + s]));
            if (part == 0) bits += 5;
            else bits += gamma_bits(zz_enc((int32_t)k2 - (int32_t)prev_k2));
            prev_k2 = k2;
        }
        if (best_bits > bits) { best_bits = bits; best = porder; }
    }
    *porder_out = best;
    if (k2_out) {
        uint32_t k1;
        for (part = 0; part < (1u << best); part++)
            lo_rice_parameter(part_mean_scratch[best][part], &k1, &k2_out[part]);
    }
    return best_bits + 10;
}

/* linne_coder.c:130-147 and :281-302 (emit half) */
static void coder_emit(bitw *w, const int32_t *res, uint32_t n, uint32_t *porder_used)
{
    static uint32_t k2s[1024];
    uint32_t porder, part, s, prev_k2 = 0;
    uint32_t len;
    (void)lo_coder_plan(res, n, &porder, k2s);
    len = n >> porder;
    bw_put(w, porder, 10);
    for (part = 0; part < (1u << porder); part++) {
        const uint32_t k2 = k2s[part], k1 = k2 + 1;
        if (part == 0) bw_put(w, k2, 5);
        else put_gamma(w, zz_enc((int32_t)k2 - (int32_t)prev_k2));
        prev_k2 = k2;
        for (s = 0; s < len; s++) {
            uint32_t uv = zz_enc(res[part * len + s]);
            if (uv < (1u << k1)) {
                bw_put(w, 1, 1);
                bw_put(w, uv, k1);
            } else {
                uv -= (1u << k1);
                bw_zero_run(w, 1u + (uint64_t)(uv >> k2));
                bw_put(w, uv & ((1u << k2) - 1u), k2);
            }
        }
    }
    if (porder_used) *porder_used = porder;
}

/* linne_coder.c:150-169 and :306-327 */
static void coder_decode(bitr *r, int32_t *out, uint32_t n)
{
    const uint32_t porder = br_get(r, 10);
    const uint32_t len = n >> porder;
    uint32_t part, s, k2 = 0;
    for (part = 0; part < (1u << porder); part++) {
        uint32_t k1;
        if (part == 0) k2 = br_get(r, 5);
        else k2 = (uint32_t)((int32_t)k2 + zz_dec(get_gamma(r)));
        k1 = k2 + 1;
        for (s = 0; s < len; s++) {
            const uint32_t q = br_zero_run(r);
            uint32_t uv;
            if (q == 0) uv = br_get(r, k1 & 31u);
            else uv = br_get(r, k2 & 31u) + (1u << (k1 & 31u)) + ((q - 1u) << (k2 & 31u));
            if (part * len + s < n) out[part * len + s] = zz_dec(uv);
        }
        if (r->pos > r->size_bits + 64) break;      /* ran off the payload: stop early (corrupt data) */
    }
}

/* ------------------------------------------------------------------------------------------
 * Block encode
 * ------------------------------------------------------------------------------------------ */

/* linne_encoder.c:480-529 */
static int decide_block_type(lo_encoder *e, const int32_t *const *in, uint32_t n)
{
    const lo_stream_info *h = &e->info;
    const double norm = pow(2.0, -(int32_t)(h->bits_per_sample - 1));
    double mean = 0.0;
    uint32_t c, i;
    for (c = 0; c < h->num_channels; c++) {
        for (i = 0; i < n; i++) e->fbuf[i] = in[c][i] * norm;
        mean += estimate_bits_per_sample(&e->lpc, e->fbuf, n, h->bits_per_sample, e->layers[0].num_params);
    }
    mean /= h->num_channels;
    mean /= h->bits_per_sample;
    if (mean >= 0.95f) return LO_BLOCK_RAW;
    for (c = 0; c < h->num_channels; c++)
        for (i = 0; i < n; i++) if (in[c][i] != 0) return LO_BLOCK_COMPRESSED;
    return LO_BLOCK_SILENT;
}

/* linne_encoder.c:532-591 */
static int encode_raw(const lo_stream_info *h, const int32_t *const *in, uint32_t n,
                      uint8_t *out, uint32_t cap, uint32_t *size)
{
    const uint32_t bytes = h->bits_per_sample / 8;
    uint32_t i, c;
    uint8_t *p = out;
    if (cap < (h->bits_per_sample * n * h->num_channels) / 8) return LO_INSUFFICIENT_BUFFER;
    for (i = 0; i < n; i++)
        for (c = 0; c < h->num_channels; c++) put_be(&p, zz_enc(in[c][i]), (int)bytes);
    *size = (uint32_t)(p - out);
    return LO_OK;
}

static uint32_t analysis_length(const lo_encoder *e, uint32_t n)   /* linne_encoder.c:644-655 */
{
    uint32_t maxp = 0, na;
    int l;
    for (l = 0; l < e->preset->num_layers; l++)
        if (maxp < (uint32_t)e->preset->layer_params[l]) maxp = (uint32_t)e->preset->layer_params[l];
    na = ((n + 7u) / 8u) * 8u;
    if (na < maxp) na = maxp;
    if (na > e->info.block_size) na = e->info.block_size;
    return na;
}

/* linne_encoder.c:594-752.  `forced` != NULL skips pre-emphasis-coefficient and analysis
 * computation and uses the supplied values instead. */
static int encode_compressed(lo_encoder *e, const int32_t *const *in, uint32_t n,
                             const lo_channel_trace *forced, uint8_t *out, uint32_t cap, uint32_t *size)
{
    const lo_stream_info *h = &e->info;
    const int nl = e->preset->num_layers;
    const double norm = pow(2.0, -(int32_t)(h->bits_per_sample - 1));
    uint32_t c, i, na;
    int l, f;
    bitw w;

    for (c = 0; c < h->num_channels; c++) {
        memcpy(e->work[c], in[c], sizeof(int32_t) * n);
        if (n < e->max_block) memset(&e->work[c][n], 0, sizeof(int32_t) * (e->max_block - n));
    }
    if (h->ms) {
        if (h->num_channels < 2) return LO_INVALID_FORMAT;
        lo_ms_forward(e->work[0], e->work[1], n);
    }
    for (c = 0; c < h->num_channels; c++) {
        lo_channel_trace *tr = &e->trace[c];
        for (f = 0; f < 2; f++) {
            tr->preem_prev[f] = e->work[c][0];
            tr->preem_coef[f] = forced ? forced[c].preem_coef[f] : lo_preemphasis_coef(e->work[c], n);
            lo_preemphasis(e->work[c], n, tr->preem_prev[f], tr->preem_coef[f]);
        }
    }
    na = analysis_length(e, n);
    for (c = 0; c < h->num_channels; c++) {
        lo_channel_trace *tr = &e->trace[c];
        if (forced) {
            for (l = 0; l < nl; l++) {
                tr->num_units[l] = forced[c].num_units[l];
                tr->rshift[l] = forced[c].rshift[l];
                memcpy(tr->coef[l], forced[c].coef[l], sizeof(int32_t) * (size_t)e->preset->layer_params[l]);
            }
        } else {
            for (i = 0; i < na; i++) e->fbuf[i] = e->work[c][i] * norm;
            lo_analyze(e, e->fbuf, na, tr->num_units, tr->coef_f64);
            for (l = 0; l < nl; l++)
                lo_quantize(tr->coef_f64[l], (uint32_t)e->preset->layer_params[l], tr->coef[l], &tr->rshift[l]);
        }
    }
    for (c = 0; c < h->num_channels; c++) {
        const lo_channel_trace *tr = &e->trace[c];
        for (l = 0; l < nl; l++) {
            lo_predict(e->work[c], n, tr->coef[l], (uint32_t)e->preset->layer_params[l],
                       tr->rshift[l], tr->num_units[l], e->resid[c]);
            memcpy(e->work[c], e->resid[c], sizeof(int32_t) * n);
        }
    }

    bw_init(&w, out, cap);
    for (c = 0; c < h->num_channels; c++)
        for (f = 0; f < 2; f++) {
            bw_put(&w, zz_enc(e->trace[c].preem_prev[f]), h->bits_per_sample + 1);
            bw_put(&w, (uint32_t)e->trace[c].preem_coef[f], 4);
        }
    for (c = 0; c < h->num_channels; c++)
        for (l = 0; l < nl; l++) {
            const lo_channel_trace *tr = &e->trace[c];
            bw_put(&w, log2_ceil(tr->num_units[l]), 3);
            bw_put(&w, tr->rshift[l], 4);
            for (i = 0; i < (uint32_t)e->preset->layer_params[l]; i++) {
                const uint32_t sym = zz_enc(tr->coef[l][i]) & 0xFFu;
                bw_put(&w, e->huff_code[sym], e->huff_len[sym]);
            }
        }
    for (c = 0; c < h->num_channels; c++) {
        const uint64_t before = w.pos;
        coder_emit(&w, e->resid[c], n, &e->trace[c].porder);
        e->trace[c].residual_bits = (uint32_t)(w.pos - before);
    }
    if (w.overflow) return LO_INSUFFICIENT_BUFFER;
    *size = bw_bytes(&w);
    return LO_OK;
}

/* linne_encoder.c:774-862 */
static int encode_block_impl(lo_encoder *e, const int32_t *const *input, uint32_t n,
                             const lo_channel_trace *forced, uint8_t *out, uint32_t cap, uint32_t *out_size)
{
    uint8_t *p = out;
    uint32_t payload = 0;
    int type, ret;
    if (!e || !input || !n || !out || !cap || !out_size) return LO_INVALID_ARGUMENT;
    if (!e->configured) return LO_PARAMETER_NOT_SET;
    if (n > e->info.block_size) return LO_INSUFFICIENT_BUFFER;
    if (cap < LO_BLOCK_HEADER) return LO_INSUFFICIENT_BUFFER;
    type = forced ? LO_BLOCK_COMPRESSED : decide_block_type(e, input, n);
    e->last_block_type = type;
    put_be(&p, 0xFFFF, 2); put_be(&p, 0, 4); put_be(&p, 0, 2);
    put_be(&p, (uint32_t)type, 1); put_be(&p, n, 2);
    switch (type) {
    case LO_BLOCK_RAW: ret = encode_raw(&e->info, input, n, p, cap - LO_BLOCK_HEADER, &payload); break;
    case LO_BLOCK_COMPRESSED: ret = encode_compressed(e, input, n, forced, p, cap - LO_BLOCK_HEADER, &payload); break;
    default: ret = LO_OK; payload = 0; break;
    }
    if (ret != LO_OK) return ret;
    p = out + 2; put_be(&p, payload + 5, 4);
    p = out + 6; put_be(&p, lo_crc16(out + 8, payload + 3), 2);
    *out_size = LO_BLOCK_HEADER + payload;
    return LO_OK;
}

int lo_encode_block(lo_encoder *e, const int32_t *const *input, uint32_t n,
                    uint8_t *out, uint32_t cap, uint32_t *out_size)
{
    return encode_block_impl(e, input, n, NULL, out, cap, out_size);
}

int lo_encode_block_forced(lo_encoder *e, const int32_t *const *input, uint32_t n,
                           const lo_channel_trace *forced, uint8_t *out, uint32_t cap, uint32_t *out_size)
{
    return encode_block_impl(e, input, n, forced, out, cap, out_size);
}

/* linne_encoder.c:865-932 */
int lo_encode_whole(lo_encoder *e, const int32_t *const *input, uint32_t num_samples,
                    uint8_t *out, uint32_t cap, uint32_t *out_size)
{
    const int32_t *ptr[LO_MAX_CHANNELS];
    uint32_t done = 0, off = LO_HEADER_SIZE, c, wrote;
    int ret;
    if (!e || !input || !out || !out_size) return LO_INVALID_ARGUMENT;
    if (!e->configured) return LO_PARAMETER_NOT_SET;
    e->info.num_samples = num_samples;
    if ((ret = lo_header_encode(&e->info, out, cap)) != LO_OK) return ret;
    while (done < num_samples) {
        const uint32_t n = (num_samples - done < e->info.block_size) ? num_samples - done : e->info.block_size;
        for (c = 0; c < e->info.num_channels; c++) ptr[c] = &input[c][done];
        if ((ret = lo_encode_block(e, ptr, n, out + off, cap - off, &wrote)) != LO_OK) return ret;
        off += wrote; done += n;
    }
    *out_size = off;
    return LO_OK;
}

/* ------------------------------------------------------------------------------------------
 * Block decode.  libs/linne_decoder/src/linne_decoder.c:564-668 (framing, check order),
 * :430-526 (compressed), :357-427 (raw), :529-561 (silent)
 * ------------------------------------------------------------------------------------------ */
typedef struct { uint32_t child0[512], child1[512], root; } huff_tree;

static void huff_tree_build(huff_tree *t)
{
    /* same merge order as lo_huffman_build; kept as a tree for the bit-serial walk (static_huffman.c:145-165) */
    uint64_t weight[512];
    uint32_t live = 256, next = 256, i;
    for (i = 0; i < 256; i++) weight[i] = k_coef_freq[i] ? k_coef_freq[i] : 1;
    for (i = 256; i < 512; i++) weight[i] = 0;
    while (live > 1) {
        uint32_t a = UINT32_MAX, b = UINT32_MAX;
        for (i = 0; i < next; i++) {
            if (weight[i] == 0) continue;
            if (a == UINT32_MAX || weight[i] < weight[a]) { b = a; a = i; }
            else if (b == UINT32_MAX || weight[i] < weight[b]) { b = i; }
        }
        weight[next] = (uint32_t)(weight[a] + weight[b]);
        weight[a] = weight[b] = 0;
        t->child0[next] = a; t->child1[next] = b;
        next++; live--;
    }
    t->root = next - 1;
}

int lo_decode_block(const lo_stream_info *h, int check_crc, const uint8_t *data, uint32_t size,
                    int32_t **out, uint32_t out_channels, uint32_t out_samples,
                    uint32_t *consumed, uint32_t *decoded_samples)
{
    static huff_tree tree; static int tree_ready = 0;
    const lo_preset *ps;
    const uint8_t *p = data;
    uint32_t blk_size, n, type, c, i, payload_used = 0;
    uint16_t crc;
    if (!h || !data || !out || !consumed || !decoded_samples) return LO_INVALID_ARGUMENT;
    if (out_channels < h->num_channels) return LO_INSUFFICIENT_BUFFER;
    if (size < LO_BLOCK_HEADER) return (size >= 2 && get_be(&p, 2) != 0xFFFF) ? LO_INVALID_FORMAT : LO_INSUFFICIENT_DATA;
    if (get_be(&p, 2) != 0xFFFF) return LO_INVALID_FORMAT;
    blk_size = get_be(&p, 4);
    if ((uint64_t)blk_size + 6 > size) return LO_INSUFFICIENT_DATA;
    crc = (uint16_t)get_be(&p, 2);
    if (check_crc && lo_crc16(p, blk_size - 2) != crc) return LO_DATA_CORRUPTION;
    type = get_be(&p, 1);
    n = get_be(&p, 2);
    if (n > out_samples) return LO_INSUFFICIENT_BUFFER;
    ps = &k_presets[h->preset];
    if (type == LO_BLOCK_RAW) {
        const uint32_t bytes = h->bits_per_sample / 8;
        if (size - LO_BLOCK_HEADER < (h->bits_per_sample * n * h->num_channels) / 8) return LO_INSUFFICIENT_DATA;
        for (i = 0; i < n; i++)
            for (c = 0; c < h->num_channels; c++) out[c][i] = zz_dec(get_be(&p, (int)bytes));
        payload_used = (uint32_t)(p - (data + LO_BLOCK_HEADER));
    } else if (type == LO_BLOCK_SILENT) {
        for (c = 0; c < h->num_channels; c++) memset(out[c], 0, sizeof(int32_t) * n);
    } else if (type == LO_BLOCK_COMPRESSED) {
        int32_t prev[LO_MAX_CHANNELS][2], pcoef[LO_MAX_CHANNELS][2];
        uint32_t units[LO_MAX_CHANNELS][LO_MAX_LAYERS], rsh[LO_MAX_CHANNELS][LO_MAX_LAYERS];
        static int32_t coef[LO_MAX_CHANNELS][LO_MAX_LAYERS][LO_MAX_PARAMS];
        bitr r;
        int l, f;
        if (!tree_ready) { huff_tree_build(&tree); tree_ready = 1; }
        br_init(&r, p, size - LO_BLOCK_HEADER);
        for (c = 0; c < h->num_channels; c++)
            for (f = 0; f < 2; f++) {
                prev[c][f] = zz_dec(br_get(&r, h->bits_per_sample + 1));
                pcoef[c][f] = (int32_t)br_get(&r, 4);
            }
        for (c = 0; c < h->num_channels; c++)
            for (l = 0; l < ps->num_layers; l++) {
                units[c][l] = 1u << br_get(&r, 3);
                rsh[c][l] = br_get(&r, 4);
                for (i = 0; i < (uint32_t)ps->layer_params[l]; i++) {
                    uint32_t node = tree.root;
                    do { node = br_bit(&r) ? tree.child1[node] : tree.child0[node]; }
                    while (node >= 256 && r.pos <= r.size_bits + 64);
                    coef[c][l][i] = zz_dec(node & 0xFFu);
                }
            }
        for (c = 0; c < h->num_channels; c++) coder_decode(&r, out[c], n);
        payload_used = (uint32_t)((r.pos + 7) >> 3);
        for (c = 0; c < h->num_channels; c++) {
            for (l = ps->num_layers - 1; l >= 0; l--)
                lo_synthesize(out[c], n, coef[c][l], (uint32_t)ps->layer_params[l], rsh[c][l], units[c][l]);
            lo_deemphasis2(out[c], n, prev[c], pcoef[c]);
        }
        if (h->ms) {
            if (h->num_channels < 2) return LO_INVALID_FORMAT;
            lo_ms_inverse(out[0], out[1], n);
        }
    } else {
        return LO_INVALID_FORMAT;
    }
    *consumed = LO_BLOCK_HEADER + payload_used;
    *decoded_samples = n;
    return LO_OK;
}

/* linne_decoder.c:671-730 */
int lo_decode_whole(const uint8_t *data, uint32_t size, int check_crc,
                    int32_t **out, uint32_t out_channels, uint32_t out_samples)
{
    lo_stream_info h;
    int32_t *ptr[LO_MAX_CHANNELS];
    uint32_t done = 0, off = LO_HEADER_SIZE, c, used, got;
    int ret;
    if (!data || !out) return LO_INVALID_ARGUMENT;
    if ((ret = lo_header_decode(data, size, &h)) != LO_OK) return ret;
    if (!header_valid(&h)) return LO_INVALID_FORMAT;
    if (h.num_channels > LO_MAX_CHANNELS) return LO_INSUFFICIENT_BUFFER;
    if (out_channels < h.num_channels || out_samples < h.num_samples) return LO_INSUFFICIENT_BUFFER;
    while (done < h.num_samples && off < size) {
        for (c = 0; c < h.num_channels; c++) ptr[c] = &out[c][done];
        ret = lo_decode_block(&h, check_crc, data + off, size - off, ptr, out_channels,
                              out_samples - done, &used, &got);
        if (ret != LO_OK) return ret;
        off += used; done += got;
    }
    return LO_OK;
}
