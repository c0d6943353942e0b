/* lnb_analyze_v2.cuh -- cooperative analysis kernel: one CTA per slot (block, channel, regulariser).
 *
 * Replaces the flat search / select / forward kernels (stages E2..E4 of lnb_pipeline.cuh) for blocks
 * whose analysis length is a multiple of 2048 and at most 10240 samples (the CLI default block);
 * other blocks keep the flat kernels.  Covers reference rows a7-a13:
 * libs/linne_network/src/linne_network.c:268-347 (unit-count search), :350-376, :165-210 (forward),
 * libs/lpc/src/lpc.c:196-205 (Welch window), :215-249 (autocorrelation), :252-324 (Levinson).
 *
 * Layout and mapping
 *   - The block-channel's signal lives in shared memory for the whole layer cascade (2 x na doubles:
 *     A = layer input, B = windowed copy during the search / layer output after the forward pass).
 *   - 256 threads; thread c owns the T = na/256 consecutive samples [c*T, c*T+T).  Every unit
 *     count U in 1..128 divides 256, so a chunk never straddles a unit.
 *   - Shared-memory arrays are stored TRANSPOSED: sample i = c*T + e sits at [e*256 + c].  Lanes of a
 *     warp (consecutive c) then always touch consecutive 8-byte words -- in their own chunk and when a
 *     sliding window runs over into the following chunks -- so every LDS.64 is conflict-free.
 *   - Autocorrelation: per thread a 8-sample x 16-lag register tile (128 multiply-adds per 16 loads);
 *     chunk partials are reduced through shared memory in a fixed order (deterministic).
 *   - Levinson-Durbin: orders <= 16 one thread per unit with the reference's exact operation order;
 *     orders 32..128 one warp per unit (strided partial dot products + xor-butterfly).
 *   - Residual evaluation / forward filter: 8-output register tile, taps broadcast from shared memory.
 *
 * FP64 is the bounding pipe: ~2*(2P-1)+K multiply-adds per sample per layer pass.
 */
#pragma once
#include "lnb_common.cuh"
#include "lnb_encode_core.cuh"

#define LNB_AN_THREADS   256
#define LNB_AN_MAX_NA    10240
#define LNB_AN_LAGS      16          /* lags per register tile */

/* dynamic shared memory in doubles */
__host__ __device__ inline size_t lnb_an_smem_doubles(uint32_t na_max)
{
    return (size_t)2 * na_max                      /* A, B */
         + (size_t)LNB_MAX_LEVELS * LNB_MAX_PARAMS /* candidate coefficients per level */
         + (size_t)LNB_MAX_LEVELS * 256            /* autocorrelations per level: U*(p+1) = P+U <= 256 */
         + (size_t)LNB_AN_LAGS * LNB_AN_THREADS    /* per-thread partials of one lag tile; aliased by the
                                                      per-warp Levinson scratch (8 x 2 x 130 doubles) */
         + (size_t)LNB_AN_LAGS * 16                /* stage-1 segment sums */
         + 64;                                     /* level losses, misc */
}

struct LnbAnCtx {
    double *A, *B, *cand, *acorr, *part, *seg, *lev, *misc;
    uint32_t na, T;
};

/* sample at logical index i (>= 0) of a transposed array */
__device__ __forceinline__ double lnb_an_ld(const double *X, uint32_t i, uint32_t T)
{
    return X[(i % T) * LNB_AN_THREADS + (i / T)];
}

/* ---- Welch-windowed copy A -> B for unit length m (lpc.c:196-205) ---- */
__device__ __forceinline__ void lnb_an_window(const LnbAnCtx &cx, uint32_t m, double scale)
{
    const uint32_t c = threadIdx.x, T = cx.T;
    const uint32_t pos0 = (c * T) % m;
    for (uint32_t e = 0; e < T; e++) {
        const uint32_t pos = pos0 + e;
        const uint32_t q = (pos < m - 1u - pos) ? pos : (m - 1u - pos);
        const double w = __dmul_rn(__dmul_rn(scale, (double)q), (double)(m - 1u - q));
        cx.B[e * LNB_AN_THREADS + c] = __dmul_rn(cx.A[e * LNB_AN_THREADS + c], w);
    }
}

/* ---- autocorrelation of every unit of one level: r[u][0..p] into acorr_lvl[u*(p+1) + lag] ---- */
__device__ void lnb_an_autocorr(const LnbAnCtx &cx, uint32_t U, uint32_t p, double *acorr_lvl)
{
    const uint32_t c = threadIdx.x, T = cx.T;
    const uint32_t m = cx.na / U, cpu = LNB_AN_THREADS / U;   /* chunks per unit */
    const uint32_t pos0 = (c % cpu) * T;                      /* unit-local position of the chunk */
    const uint32_t groups = (p + 1u + LNB_AN_LAGS - 1u) / LNB_AN_LAGS;

    for (uint32_t g = 0; g < groups; g++) {
        const uint32_t k0 = g * LNB_AN_LAGS;
        double acc[LNB_AN_LAGS];
        double wv[LNB_AN_LAGS + 7];
#pragma unroll
        for (int k = 0; k < LNB_AN_LAGS; k++) acc[k] = 0.0;
        /* Cursor of the window's next element.  Its logical offset r from the chunk start is the same in
         * every thread, so the wrap logic (r -> element ew of chunk c + dq) runs on warp-uniform values;
         * per thread only the address add and the unit-end test (r < lim) remain. */
        const uint32_t lim = m - pos0;                       /* offsets at or past the unit end read as zero */
        const double *Bc = cx.B + c;
        uint32_t r = k0, ew = k0 % T, dq = k0 / T;
#pragma unroll
        for (int k = 0; k < LNB_AN_LAGS - 1; k++) {
            wv[k] = (r < lim) ? Bc[ew * LNB_AN_THREADS + dq] : 0.0;
            r++; if (++ew == T) { ew = 0; dq++; }
        }
        for (uint32_t e0 = 0; e0 < T; e0 += 8u) {
#pragma unroll
            for (int k = 0; k < 8; k++) {
                wv[LNB_AN_LAGS - 1 + k] = (r < lim) ? Bc[ew * LNB_AN_THREADS + dq] : 0.0;
                r++; if (++ew == T) { ew = 0; dq++; }
            }
#pragma unroll
            for (int o = 0; o < 8; o++) {
                const double w = Bc[(e0 + o) * LNB_AN_THREADS];
#pragma unroll
                for (int k = 0; k < LNB_AN_LAGS; k++) acc[k] = lnb_mac(w, wv[o + k], acc[k]);
            }
#pragma unroll
            for (int k = 0; k < LNB_AN_LAGS - 1; k++) wv[k] = wv[k + 8];
        }
        /* fixed-order reduction of the chunk partials: warp w sums lags 2w and 2w+1; lanes read
         * consecutive words (conflict-free), then an xor-butterfly inside each unit's lane segment */
#pragma unroll
        for (int k = 0; k < LNB_AN_LAGS; k++) cx.part[k * LNB_AN_THREADS + c] = acc[k];
        __syncthreads();
        {
            const uint32_t warp = c >> 5, lane = c & 31u;
            for (uint32_t kk = 0; kk < 2u; kk++) {
                const uint32_t k = 2u * warp + kk;
                const double *row = cx.part + k * LNB_AN_THREADS;
                if (cpu >= 32u) {
                    for (uint32_t u = 0; u < U; u++) {
                        double v = 0.0;
                        for (uint32_t q = 0; q < cpu; q += 32u) v += row[u * cpu + q + lane];
#pragma unroll
                        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
                        if (lane == 0 && k0 + k <= p) acorr_lvl[u * (p + 1u) + k0 + k] = v;
                    }
                } else {
                    for (uint32_t q = 0; q < LNB_AN_THREADS; q += 32u) {
                        double v = row[q + lane];
                        for (uint32_t off = cpu >> 1; off > 0u; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, (int)off);
                        if ((lane & (cpu - 1u)) == 0u && k0 + k <= p) acorr_lvl[((q + lane) / cpu) * (p + 1u) + k0 + k] = v;
                    }
                }
            }
        }
        __syncthreads();
    }
}

/* ---- Levinson-Durbin by one warp (orders 32..128); result reversed into out_w[0..p) ---- */
__device__ void lnb_an_levinson_warp(const double *r_in, uint32_t p, double lambda, double *scratch, double *out_w)
{
    const uint32_t lane = threadIdx.x & 31u;
    double *a = scratch, *an = scratch + (LNB_MAX_PARAMS + 2);
    const double r0 = __dmul_rn(r_in[0], __dadd_rn(1.0, lambda));
    if (fabs(r0) < (double)FLT_EPSILON) {
        for (uint32_t i = lane; i < p; i += 32u) out_w[i] = 0.0;
        return;
    }
    for (uint32_t i = lane; i < p + 2u; i += 32u) { a[i] = 0.0; an[i] = 0.0; }
    __syncwarp();
    double err = r0;
    const double a1 = -r_in[1] / r0;
    if (lane == 0) { a[0] = 1.0; a[1] = a1; }
    err = __dadd_rn(err, __dmul_rn(r_in[1], a1));
    __syncwarp();
    for (uint32_t k = 1; k < p; k++) {
        double part = 0.0;
        for (uint32_t i = lane; i <= k; i += 32u) {
            const double rv = (k + 1u - i == 0u) ? r0 : r_in[k + 1u - i];
            part = __dadd_rn(part, __dmul_rn(a[i], rv));
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) part += __shfl_xor_sync(0xffffffffu, part, off);
        const double gamma = part / -err;
        err = __dmul_rn(err, __dadd_rn(1.0, -__dmul_rn(gamma, gamma)));
        for (uint32_t i = 1u + lane; i <= k; i += 32u) an[i] = __dadd_rn(a[i], __dmul_rn(gamma, a[k + 1u - i]));
        if (lane == 0) { an[0] = 1.0; an[k + 1u] = gamma; }
        __syncwarp();
        double *t = a; a = an; an = t;
    }
    for (uint32_t j = lane; j < p; j += 32u) out_w[j] = a[p - j];
    __syncwarp();
}

/* ---- FIR over the thread's chunk: res[t] = init + sum_j w[j] * x[t-p+j]  (x[<0] = 0) ----
 * MODE 0: search loss (init = x[t], unit 0 skips t = 0)   linne_network.c:318-335
 * MODE 1: forward (y = x[t] + sum, written to Y, all samples counted)   linne_network.c:183-208 */
template <int MODE>
__device__ double lnb_an_fir(const LnbAnCtx &cx, const double *X, double *Y, uint32_t U, uint32_t p, const double *cand_lvl)
{
    const uint32_t c = threadIdx.x, T = cx.T;
    const uint32_t cpu = LNB_AN_THREADS / U, u = c / cpu;
    const double *w = cand_lvl + u * p;
    double loss = 0.0;
    const double *Xc = X + c;
    const int32_t first_valid = -(int32_t)(c * T);          /* chunk-relative offsets below this lie before the block: zero */
    for (uint32_t e0 = 0; e0 < T; e0 += 8u) {
        const uint32_t t0 = c * T + e0;
        double acc[8], xw[15];
#pragma unroll
        for (int o = 0; o < 8; o++) acc[o] = (MODE == 0) ? Xc[(e0 + o) * LNB_AN_THREADS] : 0.0;
        /* cursor over x starting at chunk-relative offset rr = e0 - p (same in every thread): element ew of
         * chunk c + dq, with floor division for negative offsets */
        int32_t rr = (int32_t)e0 - (int32_t)p;
        int32_t dq = (rr >= 0) ? rr / (int32_t)T : -(((int32_t)T - 1 - rr) / (int32_t)T);
        int32_t ew = rr - dq * (int32_t)T;
#pragma unroll
        for (int k = 0; k < 7; k++) {
            xw[k] = (rr >= first_valid) ? Xc[ew * LNB_AN_THREADS + dq] : 0.0;
            rr++; if (++ew == (int32_t)T) { ew = 0; dq++; }
        }
        for (uint32_t j0 = 0; j0 < p; j0 += 8u) {
            const uint32_t nj = (p - j0 < 8u) ? p - j0 : 8u;
#pragma unroll
            for (int k = 0; k < 8; k++) {
                /* element t0 - p + j0 + 7 + k; never read past t0 + 7 */
                const bool need = (uint32_t)k < nj;
                xw[7 + k] = (need && rr >= first_valid) ? Xc[ew * LNB_AN_THREADS + dq] : 0.0;
                if (need) { rr++; if (++ew == (int32_t)T) { ew = 0; dq++; } }
            }
#pragma unroll
            for (int jj = 0; jj < 8; jj++) {
                if ((uint32_t)jj < nj) {
                    const double wj = w[j0 + jj];
#pragma unroll
                    for (int o = 0; o < 8; o++) acc[o] = lnb_mac(wj, xw[jj + o], acc[o]);
                }
            }
#pragma unroll
            for (int k = 0; k < 7; k++) xw[k] = xw[k + 8];
        }
#pragma unroll
        for (int o = 0; o < 8; o++) {
            if (MODE == 0) {
                if (!(t0 + o == 0u)) loss += fabs(acc[o]);
            } else {
                const double y = X[(e0 + o) * LNB_AN_THREADS + c] + acc[o];
                Y[(e0 + o) * LNB_AN_THREADS + c] = y;
                loss += fabs(y);
            }
        }
    }
    return loss;
}

/* block-wide sum in a fixed order; result valid in every thread */
__device__ double lnb_an_block_sum(double v, double *scratch /* >= 9 doubles */)
{
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    __syncthreads();
    if ((threadIdx.x & 31u) == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
    for (int w = 0; w < LNB_AN_THREADS / 32; w++) s += scratch[w];
    __syncthreads();
    return s;
}

__global__ void __launch_bounds__(LNB_AN_THREADS, 1) lnb_analyze_v2_kernel(LnbEncodeBatch b, uint32_t na_max)
{
    extern __shared__ __align__(16) double lnb_an_smem[];
    const uint32_t s = blockIdx.x, c = threadIdx.x;
    const uint32_t bc = s / b.cfg.num_lambdas, lam = s % b.cfg.num_lambdas;
    const uint32_t blk_i = bc / b.cfg.num_channels;
    const LnbBlockDesc blk = b.blocks[blk_i];
    if (blk.type != LNB_BLOCK_COMPRESSED || !(blk.status & LNB_ENC_FLAG_FAST)) return;

    LnbAnCtx cx;
    cx.na = blk.na; cx.T = blk.na / LNB_AN_THREADS;
    cx.A = lnb_an_smem;
    cx.B = cx.A + na_max;
    cx.cand = cx.B + na_max;
    cx.acorr = cx.cand + LNB_MAX_LEVELS * LNB_MAX_PARAMS;
    cx.part = cx.acorr + LNB_MAX_LEVELS * 256;
    cx.seg = cx.part + LNB_AN_LAGS * LNB_AN_THREADS;
    cx.lev = cx.part;                              /* Levinson phase never overlaps the autocorrelation phase */
    cx.misc = cx.seg + LNB_AN_LAGS * 16;
    const uint32_t T = cx.T, na = cx.na;
    const double lambda = b.cfg.lambdas[lam];

    /* layer-0 input: normalised work signal (linne_encoder.c:661-663) */
    {
        const int32_t *src = b.work + (size_t)bc * b.cfg.work_stride + (size_t)c * T;
        const double norm = ldexp(1.0, -(int)(b.cfg.bits_per_sample - 1u));
        for (uint32_t e = 0; e < T; e++) cx.A[e * LNB_AN_THREADS + c] = (double)src[e] * norm;
    }
    __syncthreads();

    double final_loss = 0.0;
    for (uint32_t l = 0; l < b.cfg.num_layers; l++) {
        const uint32_t P = b.cfg.layer_params[l];
        uint32_t nlev = 0;
        while (nlev < LNB_MAX_LEVELS && (1u << nlev) <= P) nlev++;        /* U = 1 .. min(128, P) */

        /* ---- autocorrelation of every unit of every level ---- */
        for (uint32_t lv = 0; lv < nlev; lv++) {
            const uint32_t U = 1u << lv, p = P / U, m = na / U;
            lnb_an_window(cx, m, b.welch[(size_t)blk_i * LNB_MAX_LEVELS + lv]);
            __syncthreads();
            lnb_an_autocorr(cx, U, p, cx.acorr + lv * 256);
            __syncthreads();
        }

        /* ---- Levinson-Durbin: thread-serial for p <= 16, warp-cooperative above ---- */
        {
            /* enumerate thread tasks: (level, unit) with p <= 16 */
            uint32_t task = c;
            for (uint32_t lv = 0; lv < nlev; lv++) {
                const uint32_t U = 1u << lv, p = P / U;
                if (p > 16u) continue;
                if (task < U) {
                    double r[17], a[18], coef[16];
                    const double *src = cx.acorr + lv * 256 + task * (p + 1u);
                    for (uint32_t k = 0; k <= p; k++) r[k] = src[k];
                    r[0] = __dmul_rn(r[0], __dadd_rn(1.0, lambda));
                    lnb_levinson(r, p, a, coef, (double *)0);
                    double *dst = cx.cand + lv * LNB_MAX_PARAMS + task * p;
                    for (uint32_t j = 0; j < p; j++) dst[j] = coef[p - 1u - j];
                    task = 0xFFFFFFFFu;
                } else if (task != 0xFFFFFFFFu) {
                    task -= U;
                }
            }
            /* warp tasks */
            const uint32_t warp = c >> 5;
            uint32_t wt = warp;
            for (uint32_t lv = 0; lv < nlev; lv++) {
                const uint32_t U = 1u << lv, p = P / U;
                if (p <= 16u) continue;
                if (wt < U) {
                    lnb_an_levinson_warp(cx.acorr + lv * 256 + wt * (p + 1u), p, lambda,
                                         cx.lev + warp * 2 * (LNB_MAX_PARAMS + 2),
                                         cx.cand + lv * LNB_MAX_PARAMS + wt * p);
                    wt = 0xFFFFFFFFu;
                } else if (wt != 0xFFFFFFFFu) {
                    wt -= U;
                }
            }
        }
        __syncthreads();

        /* ---- L1 loss of every level, first minimum wins (linne_network.c:337-341) ---- */
        for (uint32_t lv = 0; lv < nlev; lv++) {
            const uint32_t U = 1u << lv, p = P / U;
            const double part = lnb_an_fir<0>(cx, cx.A, (double *)0, U, p, cx.cand + lv * LNB_MAX_PARAMS);
            const double tot = lnb_an_block_sum(part, cx.misc + 16);
            if (c == 0) cx.misc[lv] = tot / (double)na;
        }
        __syncthreads();
        uint32_t best = 0;
        {
            double best_loss = (double)FLT_MAX;
            bool found = false;
            for (uint32_t lv = 0; lv < nlev; lv++)
                if (cx.misc[lv] < best_loss) { best_loss = cx.misc[lv]; best = lv; found = true; }
            if (!found) best = 0;
        }
        if (c == 0) b.chosen_log2u[(size_t)s * LNB_MAX_LAYERS + l] = (uint8_t)best;
        {
            double *dst = b.chosen_w + ((size_t)s * LNB_MAX_LAYERS + l) * LNB_MAX_PARAMS;
            const double *src = cx.cand + best * LNB_MAX_PARAMS;
            for (uint32_t k = c; k < P; k += LNB_AN_THREADS) dst[k] = src[k];
        }

        /* ---- forward: this layer's residual becomes the next layer's input ---- */
        {
            const uint32_t U = 1u << best, p = P / U;
            const double part = lnb_an_fir<1>(cx, cx.A, cx.B, U, p, cx.cand + best * LNB_MAX_PARAMS);
            final_loss = lnb_an_block_sum(part, cx.misc + 16);
        }
        __syncthreads();
        double *t = cx.A; cx.A = cx.B; cx.B = t;
    }
    /* total |residual| of the cascade: what picks the regulariser (linne_network.c:618-626) */
    {   /* the finish stage sums ceil(na/64) chunk slots of this analysis slot: total in slot 0, zeros after */
        const uint32_t chunks_per_slot = (b.cfg.work_stride + 63u) / 64u, nch = (na + 63u) / 64u;
        if (c < nch) b.final_sum[(size_t)s * chunks_per_slot + c] = (c == 0) ? final_loss : 0.0;
    }
}
