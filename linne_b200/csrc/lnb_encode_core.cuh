/* lnb_encode_core.cuh -- per-work-item bodies of the encode kernels.
 *
 * Stage order for a batch of blocks (every stage is one kernel over all blocks of the batch):
 *   E0 estimate     (block, channel)               entropy estimate for the raw/compressed decision
 *   E1 prepare      (block)                        block type, copy + M/S + 2x pre-emphasis
 *   E2 acorr        (block, ch, lambda, level, unit, lag)   windowed autocorrelation, one lag per item
 *      solve        (block, ch, lambda, level, unit)        Levinson-Durbin per unit
 *      loss         (block, ch, lambda, level, chunk)       L1 loss of every unit count
 *   E3 select       (block, ch, lambda)            argmin unit count, latch its coefficients
 *   E4 forward      (block, ch, lambda, chunk)     residual of the layer = next layer's input
 *      (E2..E4 repeat per layer)
 *   E5 finish       (block, ch)                    pick the regulariser, quantise coefficients
 *   E6 predict      (block, ch)                    integer predictor cascade -> residual
 *   E7 plan         (block, ch)                    residual-coder partition search, bit count
 *   E8 size/scan    (block)                        block byte sizes -> exclusive scan -> offsets
 *   E9 pack         (block)                        bit packing, block header, CRC16
 */
#pragma once
#include "lnb_common.cuh"

/* Multiply-accumulate of the O(N*P) signal loops (autocorrelation, residual evaluation, forward
 * filter).  LNB_EXACT_FP=1 keeps the reference's two roundings per term (bit-identical analysis
 * on the same inputs); LNB_EXACT_FP=0 uses one DFMA per term (half the FP64 pipe work; results
 * differ from the CPU reference in the last bits only, which the 8-bit quantiser absorbs). */
#ifndef LNB_EXACT_FP
#define LNB_EXACT_FP 0
#endif
LNB_HD double lnb_mac(double a, double b, double acc)
{
#if LNB_EXACT_FP
    return lnb_add_rn(acc, lnb_mul_rn(a, b));
#else
    return lnb_fma(a, b, acc);
#endif
}

/* a unit count 2^level takes part in the search of a layer with P parameters over na samples
 * (reference linne_network.c:291-296) */
LNB_HD bool lnb_level_valid(uint32_t level, uint32_t P, uint32_t na)
{
    const uint32_t U = 1u << level;
    return U <= P && U <= LNB_MAX_UNITS && (P % U) == 0u && (na % U) == 0u && na >= U;
}

/* ------------------------------------------------------------------------------------------
 * Levinson-Durbin on r[0..p] -> a[1..p] (returned in coef[0..p-1]); parcor optional.
 * reference libs/lpc/src/lpc.c:252-324 (same operation order, unfused arithmetic).
 * `a` is scratch of p+2 doubles.
 * ------------------------------------------------------------------------------------------ */
LNB_HD void lnb_levinson(const double *r, uint32_t p, double *a, double *coef, double *parcor)
{
    if (fabs(r[0]) < (double)FLT_EPSILON) {
        for (uint32_t i = 0; i < p; i++) coef[i] = 0.0;
        if (parcor) for (uint32_t i = 0; i <= p; i++) parcor[i] = 0.0;
        return;
    }
    for (uint32_t i = 0; i < p + 2u; i++) a[i] = 0.0;
    a[0] = 1.0;
    double err = r[0];
    a[1] = -r[1] / r[0];
    if (parcor) parcor[0] = r[1] / err;
    err = lnb_add_rn(err, lnb_mul_rn(r[1], a[1]));
    for (uint32_t k = 1; k < p; k++) {
        double gamma = 0.0;
        for (uint32_t i = 0; i < k + 1u; i++) gamma = lnb_add_rn(gamma, lnb_mul_rn(a[i], r[k + 1u - i]));
        gamma /= -err;
        err = lnb_mul_rn(err, lnb_add_rn(1.0, -lnb_mul_rn(gamma, gamma)));
        /* a_new[i] = a[i] + gamma * a[k+1-i], i = 1..k ; a_new[k+1] = gamma  (pairs updated together) */
        for (uint32_t i = 1, j = k; i <= j; i++, j--) {
            const double ai = a[i], aj = a[j];
            a[i] = lnb_add_rn(ai, lnb_mul_rn(gamma, aj));
            if (i != j) a[j] = lnb_add_rn(aj, lnb_mul_rn(gamma, ai));
        }
        a[k + 1u] = lnb_add_rn(0.0, lnb_mul_rn(gamma, 1.0));
        if (parcor) parcor[k] = -gamma;
    }
    for (uint32_t i = 0; i < p; i++) coef[i] = a[i + 1u];
}

/* ------------------------------------------------------------------------------------------
 * E0: estimated bits/sample of one channel of one block from a sine-windowed order-P0 LPC.
 * reference lpc.c:810-865 (+ window :188-195, autocorrelation :215-249).
 * The reference also adds log2(1 - parcor[P0]^2) with a value left over from an earlier call
 * (SURVEY Q1); a fresh reference handle has 0 there, and so do we.
 * ------------------------------------------------------------------------------------------ */
LNB_HD double lnb_estimate_bits(const int32_t *x, uint32_t n, uint32_t bits, uint32_t p0)
{
    const double norm = ldexp(1.0, -(int)(bits - 1u));
    const double pi = 3.1415926535897932384626433832795029;
    double r[10], hist[10], a[12], coef[10], parcor[10];
    if (p0 > 8u) p0 = 8u;                                  /* presets use 2 or 4 */
    for (uint32_t k = 0; k <= p0; k++) { r[k] = 0.0; hist[k] = 0.0; }
    /* r[k] = sum_i w[i] w[i+k]: walk i+k = j upward, keeping the last p0 windowed samples */
    for (uint32_t j = 0; j < n; j++) {
        const double w = ((double)x[j] * norm) * sin((pi * (double)j) / (double)(n - 1u));
        for (uint32_t k = p0; k >= 1u; k--) hist[k] = hist[k - 1u];
        hist[0] = w;
        for (uint32_t k = 0; k <= p0 && k <= j; k++) r[k] = lnb_add_rn(r[k], lnb_mul_rn(hist[k], w));
    }
    for (uint32_t k = 0; k <= p0; k++) parcor[k] = 0.0;
    if (n >= p0) lnb_levinson(r, p0, a, coef, parcor);
    double power = r[0] * ldexp(1.0, (int)(2u * (bits - 1u)));
    if (fabs(power) <= (double)FLT_MIN) return 0.0;
    power = log(power) * 1.4426950408889634 - log((double)n) * 1.4426950408889634;
    double ratio = 0.0;
    for (uint32_t k = 1; k < p0; k++) ratio += log(1.0 - parcor[k] * parcor[k]) * 1.4426950408889634;
    const double est = 1.9426950408889634 + 0.5 * (power + ratio);
    return est <= 0.0 ? 1.0 : est;
}

/* ------------------------------------------------------------------------------------------
 * E1: pre-emphasis coefficient and filter.  reference linne_utility.c:158-193, :196-212.
 * Sums run in sample order (exact for 16-bit input in any order; order matters for 24-bit).
 * ------------------------------------------------------------------------------------------ */
LNB_HD int32_t lnb_preemphasis_coef(const int32_t *x, uint32_t n)
{
    double c0 = 0.0, c1 = 0.0;
    double cur = (double)x[0];
    for (uint32_t i = 0; i + 1u < n; i++) {
        const double nxt = (double)x[i + 1u];
        c0 = lnb_add_rn(c0, lnb_mul_rn(cur, cur));
        c1 = lnb_add_rn(c1, lnb_mul_rn(cur, nxt));
        cur = nxt;
    }
    c1 /= c0;
    if (c0 < 1e-6 || c1 < 0.0) return 0;
    const int32_t coef = (int32_t)lnb_round_half_away(c1 * 32.0);
    return coef >= 16 ? 15 : coef;
}

LNB_HD void lnb_preemphasis(int32_t *x, uint32_t n, int32_t prev, int32_t coef)
{
    for (uint32_t i = 0; i < n; i++) {
        const int32_t cur = x[i];
        x[i] = cur - ((prev * coef) >> LNB_PREEM_SHIFT);
        prev = cur;
    }
}

/* One block: decide the type from the per-channel estimates, and for compressed blocks build the
 * integer work signal (copy, zero-pad to the analysis length, M/S, two pre-emphasis passes).
 * reference libs/linne_encoder/src/linne_encoder.c:504-528, :613-641 */
LNB_HD void lnb_prepare_block(const LnbStreamCfg &cfg, LnbBlockDesc &blk, const double *est /* [C] */,
                              const int32_t *pcm, int32_t *work /* [C][work_stride] */,
                              LnbChanParams *params /* [C] */)
{
    const uint32_t C = cfg.num_channels, n = blk.nsmp;
    double mean = 0.0;
    for (uint32_t c = 0; c < C; c++) mean += est[c];
    mean /= (double)C;
    mean /= (double)cfg.bits_per_sample;
    if (mean >= (double)LNB_RAW_THRESHOLD) { blk.type = LNB_BLOCK_RAW; return; }
    bool silent = true;
    for (uint32_t c = 0; c < C && silent; c++) {
        const int32_t *src = pcm + (size_t)c * cfg.pcm_stride + blk.smp_off;
        for (uint32_t i = 0; i < n; i++) if (src[i] != 0) { silent = false; break; }
    }
    if (silent) { blk.type = LNB_BLOCK_SILENT; return; }
    blk.type = LNB_BLOCK_COMPRESSED;

    for (uint32_t c = 0; c < C; c++) {
        const int32_t *src = pcm + (size_t)c * cfg.pcm_stride + blk.smp_off;
        int32_t *dst = work + (size_t)c * cfg.work_stride;
        for (uint32_t i = 0; i < n; i++) dst[i] = src[i];
        for (uint32_t i = n; i < cfg.work_stride; i++) dst[i] = 0;
    }
    if (cfg.ms && C >= 2u) {
        int32_t *l = work, *r = work + cfg.work_stride;
        for (uint32_t i = 0; i < n; i++) { r[i] -= l[i]; l[i] += r[i] >> 1; }     /* linne_utility.c:128-131 */
    }
    for (uint32_t c = 0; c < C; c++) {
        int32_t *x = work + (size_t)c * cfg.work_stride;
        for (int f = 0; f < LNB_NUM_PREEM; f++) {
            const int32_t prev = x[0];
            const int32_t coef = lnb_preemphasis_coef(x, n);
            params[c].preem_prev[f] = prev;
            params[c].preem_coef[f] = (uint8_t)coef;
            lnb_preemphasis(x, n, prev, coef);
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * E2 (flat path): the unit-count search of reference libs/linne_network/src/linne_network.c:268-347
 * cut into independent work items so that no single thread walks a whole block:
 *   lnb_acorr_lag      one autocorrelation lag of one unit (Welch window applied on the fly;
 *                      lpc.c:196-205, :215-249).  A lag is summed in ascending sample order, exactly
 *                      like the reference, so with LNB_EXACT_FP the value is bit-identical.
 *   lnb_solve_unit     regularise + Levinson-Durbin + coefficient reversal (lpc.c:358, :252-324,
 *                      linne_network.c:312-316)
 *   lnb_loss_chunk     sum of |residual| over a run of samples (linne_network.c:318-335)
 * The centre sample of an odd-length unit gets the window's true centre weight; the reference leaves
 * a stale value there.
 * ------------------------------------------------------------------------------------------ */
LNB_HD double lnb_welch_weight(double scale, uint32_t pos, uint32_t m)
{
    const uint32_t q = (pos < m - 1u - pos) ? pos : (m - 1u - pos);
    return lnb_mul_rn(lnb_mul_rn(scale, (double)q), (double)(m - 1u - q));
}

LNB_HD double lnb_acorr_lag(const double *xs, uint32_t m, uint32_t lag, double scale)
{
    double s = 0.0;
    if (lag >= m) return 0.0;
    for (uint32_t i = 0; i + lag < m; i++) {
        const double a = lnb_mul_rn(xs[i], lnb_welch_weight(scale, i, m));
        const double b = lnb_mul_rn(xs[i + lag], lnb_welch_weight(scale, i + lag, m));
        s = lnb_mac(a, b, s);
    }
    return s;
}

LNB_HD void lnb_solve_unit(const double *r_in, uint32_t p, uint32_t m, double lambda, double *out_w)
{
    double r[LNB_MAX_PARAMS + 1], a[LNB_MAX_PARAMS + 2], coef[LNB_MAX_PARAMS];
    if (m < p) {
        for (uint32_t i = 0; i < p; i++) coef[i] = 0.0;
    } else {
        for (uint32_t k = 0; k <= p; k++) r[k] = r_in[k];
        r[0] = lnb_mul_rn(r[0], lnb_add_rn(1.0, lambda));
        lnb_levinson(r, p, a, coef, (double *)0);
    }
    for (uint32_t j = 0; j < p; j++) out_w[j] = coef[p - 1u - j];
}

/* residual at sample t of the block-channel for the layer split into units of m samples, p taps each;
 * `w_all` = reversed coefficients of every unit, unit u at w_all + u*p.  History before the block is
 * zero (unit 0 ramp-in); later units reach back into the previous unit. */
LNB_HD double lnb_residual_at(const double *x, uint32_t t, uint32_t m, uint32_t p, const double *w_all, double init)
{
    const uint32_t u = t / m;
    const double *w = w_all + (size_t)u * p;
    double acc = init;
    const uint32_t first = (t >= p) ? 0u : p - t;            /* taps that would read before sample 0 are dropped */
    const double *h = x + t - p;
    for (uint32_t k = first; k < p; k++) acc = lnb_mac(w[k], h[k], acc);
    return acc;
}

LNB_HD double lnb_loss_chunk(const double *x, uint32_t t0, uint32_t t1, uint32_t m, uint32_t p, const double *w_all)
{
    double loss = 0.0;
    for (uint32_t t = (t0 == 0u) ? 1u : t0; t < t1; t++)     /* sample 0 is not counted (linne_network.c:319-321) */
        loss += fabs(lnb_residual_at(x, t, m, p, w_all, x[t]));
    return loss;
}

/* E4 (flat path): residual of a run of samples given the chosen coefficients, out of place.
 * reference linne_network.c:165-210 (the prediction is summed from zero, then added).  Returns sum |y|. */
LNB_HD double lnb_forward_chunk(const double *x, double *y, uint32_t t0, uint32_t t1, uint32_t m, uint32_t p,
                                const double *w_all)
{
    double sum = 0.0;
    for (uint32_t t = t0; t < t1; t++) {
        const double v = (t == 0u) ? x[0] : x[t] + lnb_residual_at(x, t, m, p, w_all, 0.0);
        y[t] = v;
        sum += fabs(v);
    }
    return sum;
}

/* ------------------------------------------------------------------------------------------
 * E5: quantise one layer's coefficients to 8 bits with a shared right shift.
 * reference lpc.c:981-1040.
 * ------------------------------------------------------------------------------------------ */
LNB_HD void lnb_quantize_layer(const double *w, uint32_t n, int8_t *q, uint8_t *rshift)
{
    double peak = 0.0;
    for (uint32_t i = 0; i < n; i++) { const double v = fabs(w[i]); if (peak < v) peak = v; }
    if (!(peak > 0.0078125)) {                         /* 2^-7; also catches NaN */
        *rshift = 8;
        for (uint32_t i = 0; i < n; i++) q[i] = 0;
        return;
    }
    int exponent;
    (void)frexp(peak, &exponent);
    int shift = 7 - exponent;
    if (shift < 1) shift = 1;                          /* |coef| >= 64: outside what the format can carry */
    if (shift > 15) shift = 15;
    double carry = 0.0;
    for (int i = (int)n - 1; i >= 0; i--) {
        carry = lnb_add_rn(carry, lnb_mul_rn(w[i], ldexp(1.0, shift)));
        int32_t v = (int32_t)lnb_round_half_away(carry);
        if (v >= 128) v = 127; else if (v < -128) v = -128;
        carry -= (double)v;
        q[i] = (int8_t)v;
    }
    *rshift = (uint8_t)shift;
}

/* ------------------------------------------------------------------------------------------
 * E6: integer predictor for one unit, in place (walks backwards so taps still see layer input).
 * reference libs/linne_encoder/src/linne_lpc_predict.c:7-38.
 * ------------------------------------------------------------------------------------------ */
LNB_HD void lnb_predict_unit_inplace(int32_t *x, uint32_t m, const int8_t *c, uint32_t p, uint32_t rshift)
{
    if (m <= p) return;
    const uint32_t half = rshift ? (1u << (rshift - 1u)) : 0u;
    for (uint32_t t = m - p; t-- > 0u;) {
        uint32_t acc = half;
        for (uint32_t k = 0; k < p; k++) acc += (uint32_t)(int32_t)c[k] * (uint32_t)x[t + k];
        x[t + p] = (int32_t)((uint32_t)x[t + p] + (uint32_t)((int32_t)acc >> rshift));
    }
}

/* ------------------------------------------------------------------------------------------
 * E7: residual coder plan.  reference libs/linne_coder/src/linne_coder.c:172-278.
 * ------------------------------------------------------------------------------------------ */
LNB_HD uint32_t lnb_rice_k2(const double *thr, double mean)
{
    /* k2 = number of thresholds (k >= 1) that mean has reached; thresholds ascend */
    uint32_t k = 0;
    while (k < 30u && mean >= thr[k + 1u]) k++;      /* k2 <= 30 keeps 1 << (k2+1) defined; needs means >= 2^31 to matter */
    return k;
}
LNB_HD uint32_t lnb_rice_len(uint32_t k2, uint32_t uval)
{
    const uint32_t k1 = k2 + 1u, thr = 1u << k1;
    return uval < thr ? k1 + 1u : k2 + 2u + ((uval - thr) >> k2);
}
LNB_HD uint32_t lnb_gamma_bits(uint32_t v) { return v == 0u ? 1u : 2u * lnb_log2_ceil(v + 2u) - 1u; }

LNB_HD uint32_t lnb_max_porder(uint32_t n)
{
    uint32_t mp = 0;
    while (mp < LNB_MAX_PORDER && (n % (1u << (mp + 1u))) == 0u) mp++;
    return mp;
}

/* `mean` scratch: 2 * LNB_MAX_PARTITIONS doubles (level j lives at offset (1<<j) - 1). */
LNB_HD void lnb_coder_plan(const double *thr, const int32_t *res, uint32_t n, double *mean, LnbCoderPlan &plan)
{
    const uint32_t maxp = lnb_max_porder(n);
    {
        const uint32_t parts = 1u << maxp, len = n / parts;
        double *top = mean + (parts - 1u);
        for (uint32_t part = 0; part < parts; part++) {
            uint64_t s = 0;                                  /* exact; equals the reference's double sum while < 2^53 */
            for (uint32_t i = 0; i < len; i++) s += lnb_zz_enc(res[part * len + i]);
            top[part] = (double)s / (double)len;
        }
        for (int lvl = (int)maxp - 1; lvl >= 0; lvl--) {
            const double *fine = mean + ((1u << (lvl + 1)) - 1u);
            double *coarse = mean + ((1u << lvl) - 1u);
            for (uint32_t part = 0; part < (1u << lvl); part++)
                coarse[part] = lnb_add_rn(fine[2u * part], fine[2u * part + 1u]) / 2.0;
        }
    }
    uint32_t best = 0, best_bits = 0xFFFFFFFFu;
    for (uint32_t porder = 0; porder <= maxp; porder++) {
        const uint32_t len = n >> porder;
        const double *lvl_mean = mean + ((1u << porder) - 1u);
        uint32_t bits = 0, prev_k2 = 0;
        for (uint32_t part = 0; part < (1u << porder); part++) {
            const uint32_t k2 = lnb_rice_k2(thr, lvl_mean[part]);
            for (uint32_t i = 0; i < len; i++) bits += lnb_rice_len(k2, lnb_zz_enc(res[part * len + i]));
            bits += (part == 0) ? 5u : lnb_gamma_bits(lnb_zz_enc((int32_t)k2 - (int32_t)prev_k2));
            prev_k2 = k2;
        }
        if (best_bits > bits) { best_bits = bits; best = porder; }
    }
    plan.porder = best;
    plan.bits = best_bits + 10u;
    const double *lvl_mean = mean + ((1u << best) - 1u);
    for (uint32_t part = 0; part < (1u << best); part++) plan.k2[part] = (uint8_t)lnb_rice_k2(thr, lvl_mean[part]);
}

/* bits of the side information of a compressed block: reference linne_encoder.c:703-735 */
LNB_HD uint32_t lnb_side_info_bits(const LnbStreamCfg &cfg, const LnbDevTables &tab, const LnbChanParams *params)
{
    uint32_t bits = cfg.num_channels * LNB_NUM_PREEM * (cfg.bits_per_sample + 1u + (LNB_PREEM_SHIFT - 1));
    for (uint32_t c = 0; c < cfg.num_channels; c++)
        for (uint32_t l = 0; l < cfg.num_layers; l++) {
            bits += 3u + 4u;
            const int8_t *q = params[c].coef + l * LNB_MAX_PARAMS;
            for (uint32_t i = 0; i < cfg.layer_params[l]; i++) bits += tab.huff_len[lnb_zz_enc(q[i]) & 0xFFu];
        }
    return bits;
}

/* ------------------------------------------------------------------------------------------
 * E9: serial MSB-first bit writer used by the one-thread-per-block packer.
 * Bit order of reference bit_stream.h:240-302; flush pads with zeros (:397-434).
 * ------------------------------------------------------------------------------------------ */
struct LnbBitWriter {
    uint8_t *dst;
    uint32_t pos;        /* bytes written */
    uint64_t acc;        /* pending bits, right-aligned */
    uint32_t nbits;      /* number of pending bits (< 8 between calls) */
};
LNB_HD void lnb_bw_open(LnbBitWriter &w, uint8_t *dst) { w.dst = dst; w.pos = 0; w.acc = 0; w.nbits = 0; }
LNB_HD void lnb_bw_put(LnbBitWriter &w, uint32_t val, uint32_t n)        /* n <= 32 */
{
    if (n == 0) return;
    const uint64_t mask = (n >= 32u) ? 0xFFFFFFFFull : ((1ull << n) - 1ull);
    w.acc = (w.acc << n) | ((uint64_t)val & mask);
    w.nbits += n;
    while (w.nbits >= 8u) {
        w.nbits -= 8u;
        w.dst[w.pos++] = (uint8_t)(w.acc >> w.nbits);
    }
    w.acc &= (1ull << w.nbits) - 1ull;
}
LNB_HD void lnb_bw_zero_run(LnbBitWriter &w, uint32_t run)               /* `run` zeros then a one */
{
    while (run >= 32u) { lnb_bw_put(w, 0, 32); run -= 32u; }
    lnb_bw_put(w, 1u, run + 1u);
}
LNB_HD uint32_t lnb_bw_close(LnbBitWriter &w)
{
    if (w.nbits) { w.dst[w.pos++] = (uint8_t)(w.acc << (8u - w.nbits)); w.nbits = 0; }
    return w.pos;
}

LNB_HD void lnb_put_gamma(LnbBitWriter &w, uint32_t v)                   /* linne_coder.c:85-103 */
{
    if (v == 0) { lnb_bw_put(w, 1, 1); return; }
    const uint32_t nd = lnb_log2_ceil(v + 2u);
    lnb_bw_put(w, 0, nd - 1u);
    lnb_bw_put(w, v + 1u, nd);
}

/* One whole block into its final place `dst` (blk.byte_size bytes): header, payload, size, CRC.
 * reference linne_encoder.c:807-858 (framing), :556-585 (raw), :699-749 (compressed), and
 * linne_coder.c:281-302 (residual emission). */
LNB_HD void lnb_pack_block(const LnbStreamCfg &cfg, const LnbDevTables &tab, const LnbBlockDesc &blk,
                           const LnbChanParams *params, const LnbCoderPlan *plans,
                           const int32_t *pcm, const int32_t *resid /* [C][work_stride] */, uint8_t *dst)
{
    const uint32_t C = cfg.num_channels, n = blk.nsmp;
    uint8_t *payload = dst + LNB_BLOCK_HEADER_SIZE;
    uint32_t payload_size = 0;
    lnb_put_be(dst, LNB_SYNC_CODE, 2);
    dst[8] = (uint8_t)blk.type;
    lnb_put_be(dst + 9, n, 2);

    if (blk.type == LNB_BLOCK_RAW) {
        const uint32_t bytes = cfg.bits_per_sample >> 3;
        uint8_t *p = payload;
        for (uint32_t i = 0; i < n; i++)
            for (uint32_t c = 0; c < C; c++) {
                lnb_put_be(p, lnb_zz_enc(pcm[(size_t)c * cfg.pcm_stride + blk.smp_off + i]), (int)bytes);
                p += bytes;
            }
        payload_size = (uint32_t)(p - payload);
    } else if (blk.type == LNB_BLOCK_COMPRESSED) {
        LnbBitWriter w;
        lnb_bw_open(w, payload);
        for (uint32_t c = 0; c < C; c++)
            for (int f = 0; f < LNB_NUM_PREEM; f++) {
                lnb_bw_put(w, lnb_zz_enc(params[c].preem_prev[f]), cfg.bits_per_sample + 1u);
                lnb_bw_put(w, params[c].preem_coef[f], LNB_PREEM_SHIFT - 1);
            }
        for (uint32_t c = 0; c < C; c++)
            for (uint32_t l = 0; l < cfg.num_layers; l++) {
                lnb_bw_put(w, params[c].log2_units[l], 3);
                lnb_bw_put(w, params[c].rshift[l], 4);
                const int8_t *q = params[c].coef + l * LNB_MAX_PARAMS;
                for (uint32_t i = 0; i < cfg.layer_params[l]; i++) {
                    const uint32_t sym = lnb_zz_enc(q[i]) & 0xFFu;
                    lnb_bw_put(w, tab.huff_code[sym], tab.huff_len[sym]);
                }
            }
        for (uint32_t c = 0; c < C; c++) {
            const LnbCoderPlan &pl = plans[c];
            const int32_t *r = resid + (size_t)c * cfg.work_stride;
            const uint32_t len = n >> pl.porder;
            uint32_t prev_k2 = 0;
            lnb_bw_put(w, pl.porder, 10);
            for (uint32_t part = 0; part < (1u << pl.porder); part++) {
                const uint32_t k2 = pl.k2[part], k1 = k2 + 1u;
                if (part == 0) lnb_bw_put(w, k2, 5);
                else lnb_put_gamma(w, lnb_zz_enc((int32_t)k2 - (int32_t)prev_k2));
                prev_k2 = k2;
                for (uint32_t i = 0; i < len; i++) {
                    uint32_t uv = lnb_zz_enc(r[part * len + i]);
                    if (uv < (1u << k1)) {
                        lnb_bw_put(w, (1u << k1) | uv, k1 + 1u);
                    } else {
                        uv -= (1u << k1);
                        lnb_bw_zero_run(w, 1u + (uv >> k2));
                        lnb_bw_put(w, uv & ((1u << k2) - 1u), k2);
                    }
                }
            }
        }
        payload_size = lnb_bw_close(w);
    }
    lnb_put_be(dst + 2, payload_size + 5u, 4);
    lnb_put_be(dst + 6, lnb_crc16_serial(tab.crc_table, dst + 8, payload_size + 3u), 2);
}
