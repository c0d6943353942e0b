/* lnb_refine_v2.cuh -- the non-default analysis paths: IRLS refinement and momentum-SGD training.
 *
 * One CTA per (block, channel).  Runs the FINAL cascade pass of
 * reference libs/linne_network/src/linne_network.c:628-629 (after the regulariser has been chosen) with
 *   - num_afmethod_iterations > 0: auxiliary-function (IRLS) refinement of every unit's coefficients,
 *     reference libs/lpc/src/lpc.c:578-633 (driver), :452-509 (weighted Gram matrix / right-hand side),
 *     :402-448 (Cholesky solve)                                           -- SURVEY row a14
 *   - enable_learning != 0: momentum-SGD on the L1 loss of the whole cascade,
 *     reference linne_network.c:805-873 (trainer), :213-265 (backward), :66-75 (L1 gradient);
 *     constants linne_internal.h:29-33 (2000 iterations, lr 0.1f, eps 1e-7), momentum 0.8f (:832)
 *                                                                         -- SURVEY row a15
 * and overwrites the chosen slot's unit counts / coefficients, which the finish stage then quantises.
 *
 * Written for generality rather than peak speed (these paths are off by default and used by no
 * BASELINE config): any analysis length (signals row-major in shared memory up to LNB_RF_MAX_NA samples, in a
 * global scratch beyond),
 * thread-strided loops, deterministic reductions.  Operation order inside a unit's Levinson recursion
 * and inside each residual is the reference's; sums across samples are tree-reduced.
 */
#pragma once
#include "lnb_common.cuh"
#include "lnb_encode_core.cuh"

#define LNB_RF_THREADS 256
#define LNB_RF_MAX_NA  10240
#define LNB_RF_HIST    LNB_MAX_PARAMS          /* zero history in front of the signal (unit 0 ramp-in) */

struct LnbRefineSmem {
    double cand[LNB_MAX_LEVELS * LNB_MAX_PARAMS];       /* reversed coefficients per unit-count level */
    double acorr[LNB_MAX_LEVELS * 256];
    double level_loss[LNB_MAX_LEVELS];
    double red[LNB_RF_THREADS / 32];
    double w[LNB_MAX_LAYERS][LNB_MAX_PARAMS];           /* chosen (reversed) coefficients per layer */
    double dw[LNB_MAX_LAYERS][LNB_MAX_PARAMS];          /* SGD: gradients */
    double mom[LNB_MAX_LAYERS][LNB_MAX_PARAMS];         /* SGD: momentum */
    double avec[LNB_MAX_PARAMS], rhs[LNB_MAX_PARAMS], inv_diag[LNB_MAX_PARAMS];
    uint32_t log2u[LNB_MAX_LAYERS];
    uint32_t flag;
};

__device__ __forceinline__ double lnb_rf_sum(double v, double *red)
{
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    __syncthreads();
    if ((threadIdx.x & 31u) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < LNB_RF_THREADS / 32; w++) s += red[w];
    __syncthreads();
    return s;
}

/* residual of sample t (x has LNB_RF_HIST zeros in front, so unit 0 simply sees zero history) */
__device__ __forceinline__ double lnb_rf_residual(const double *x, uint32_t t, uint32_t m, uint32_t p,
                                                  const double *w_all, double init)
{
    const double *w = w_all + (t / m) * p;
    const double *h = x + (int32_t)t - (int32_t)p;
    double acc = init;
    for (uint32_t k = 0; k < p; k++) acc = lnb_mac(w[k], h[k], acc);
    return acc;
}

/* packed upper-triangular index of (i, j), i <= j, dimension p */
__device__ __forceinline__ uint32_t lnb_rf_tri(uint32_t i, uint32_t j, uint32_t p) { return i * p - (i * (i - 1u)) / 2u + (j - i); }

/* IRLS refinement of one unit: xs = unit samples (unwindowed), a = coefficients a_1..a_p (model
 * res[t] = x[t] + sum_i a[i] x[t-i-1]); G = packed p x p scratch.  lpc.c:604-630 */
__device__ void lnb_rf_irls_unit(LnbRefineSmem &sm, const double *xs, uint32_t m, uint32_t p, uint32_t iters, double *G, double *wt)
{
    const uint32_t tid = threadIdx.x;
    const uint32_t ntri = p * (p + 1u) / 2u;
    double prev_obj = (double)FLT_MAX;
    for (uint32_t it = 0; it < iters; it++) {
        double obj = 0.0;                                       /* wt[m]: weights 1/max(|res|, 1e-6), global scratch */
        for (uint32_t t = p + tid; t < m; t += LNB_RF_THREADS) {
            double res = xs[t];
            for (uint32_t i = 0; i < p; i++) res = lnb_mac(sm.avec[i], xs[t - i - 1u], res);
            res = fabs(res);
            obj += res;
            wt[t] = 1.0 / ((res < 1e-6) ? 1e-6 : res);
        }
        obj = lnb_rf_sum(obj, sm.red) / (double)(m - p);
        __threadfence_block();
        /* right-hand side and upper triangle of the weighted Gram matrix (lpc.c:490-495) */
        for (uint32_t i = tid; i < p; i += LNB_RF_THREADS) {
            double s = 0.0;
            for (uint32_t t = p; t < m; t++) s -= xs[t] * xs[t - i - 1u] * wt[t];
            sm.rhs[i] = s;
        }
        for (uint32_t e = tid; e < ntri; e += LNB_RF_THREADS) {
            /* unrank e -> (i, j) */
            uint32_t i = 0, rem = e;
            while (rem >= p - i) { rem -= p - i; i++; }
            const uint32_t j = i + rem;
            double s = 0.0;
            for (uint32_t t = p; t < m; t++) s += xs[t - i - 1u] * xs[t - j - 1u] * wt[t];
            G[e] = s;
        }
        __syncthreads();
        /* Cholesky in place on the packed triangle: L(j,i) overwrites A(i,j) (lpc.c:412-429) */
        if (tid == 0) sm.flag = 0;
        __syncthreads();
        for (uint32_t i = 0; i < p; i++) {
            if (tid == 0) {
                double s = G[lnb_rf_tri(i, i, p)];
                for (int32_t k = (int32_t)i - 1; k >= 0; k--) { const double l = G[lnb_rf_tri((uint32_t)k, i, p)]; s -= l * l; }
                if (s <= 0.0) sm.flag = 1; else sm.inv_diag[i] = pow(s, -0.5);
            }
            __syncthreads();
            if (sm.flag) break;
            for (uint32_t j = i + 1u + tid; j < p; j += LNB_RF_THREADS) {
                double s = G[lnb_rf_tri(i, j, p)];
                for (int32_t k = (int32_t)i - 1; k >= 0; k--) s -= G[lnb_rf_tri((uint32_t)k, i, p)] * G[lnb_rf_tri((uint32_t)k, j, p)];
                G[lnb_rf_tri(i, j, p)] = s * sm.inv_diag[i];
            }
            __syncthreads();
        }
        if (sm.flag) {                                          /* singular: all-zero coefficients (lpc.c:614-619) */
            for (uint32_t i = tid; i < p; i += LNB_RF_THREADS) sm.avec[i] = 0.0;
            __syncthreads();
            return;
        }
        if (tid == 0) {                                         /* forward / backward substitution (lpc.c:432-445) */
            for (uint32_t i = 0; i < p; i++) {
                double s = sm.rhs[i];
                for (int32_t j = (int32_t)i - 1; j >= 0; j--) s -= G[lnb_rf_tri((uint32_t)j, i, p)] * sm.avec[j];
                sm.avec[i] = s * sm.inv_diag[i];
            }
            for (int32_t i = (int32_t)p - 1; i >= 0; i--) {
                double s = sm.avec[i];
                for (uint32_t j = (uint32_t)i + 1u; j < p; j++) s -= G[lnb_rf_tri((uint32_t)i, j, p)] * sm.avec[j];
                sm.avec[i] = s * sm.inv_diag[i];
            }
        }
        __syncthreads();
        if (fabs(prev_obj - obj) < 1e-8) break;
        prev_obj = obj;
    }
}

__global__ void __launch_bounds__(LNB_RF_THREADS, 1) lnb_refine_v2_kernel(LnbEncodeBatch b, uint32_t na_max,
                                                                          uint32_t af_iters, uint32_t learning,
                                                                          double *train_scratch, uint32_t chunks_per_slot,
                                                                          double *xy_global)
{
    extern __shared__ __align__(16) double lnb_rf_smem[];
    const uint32_t tid = threadIdx.x, bc = blockIdx.x;
    const uint32_t blk_i = bc / b.cfg.num_channels;
    const LnbBlockDesc blk = b.blocks[blk_i];
    if (blk.type != LNB_BLOCK_COMPRESSED) return;
    const uint32_t na = blk.na;
    /* the two signal buffers live in shared memory for blocks of up to LNB_RF_MAX_NA samples and in a global scratch
     * (L2-resident: 2 x 8 B per sample and block-channel) for longer ones -- the code below only sees pointers */
    double *xy = xy_global ? xy_global + (size_t)bc * 2u * (na_max + LNB_RF_HIST) : lnb_rf_smem;
    double *X = xy + LNB_RF_HIST;                                /* [na] layer input, zero history in front */
    double *Y = X + na_max + LNB_RF_HIST;                        /* [na] layer output / IRLS scratch, zero history in front */
    LnbRefineSmem &sm = *(LnbRefineSmem *)(xy_global ? lnb_rf_smem : Y + na_max);

    /* which regulariser won (linne_network.c:618-626): same rule as the finish stage */
    uint32_t best_lam = 0;
    {
        const uint32_t nch = (na + 63u) / 64u;
        double best_loss = (double)FLT_MAX;
        for (uint32_t lam = 0; lam < b.cfg.num_lambdas; lam++) {
            const size_t s = (size_t)bc * b.cfg.num_lambdas + lam;
            double sum = 0.0;
            for (uint32_t g = 0; g < nch; g++) sum += b.final_sum[s * chunks_per_slot + g];
            const double loss = sum / (double)na;
            if (loss < best_loss) { best_loss = loss; best_lam = lam; }
        }
    }
    const double lambda = b.cfg.lambdas[best_lam];
    const size_t slot = (size_t)bc * b.cfg.num_lambdas + best_lam;

    for (uint32_t i = tid; i < LNB_RF_HIST; i += LNB_RF_THREADS) { X[(int32_t)i - LNB_RF_HIST] = 0.0; Y[(int32_t)i - LNB_RF_HIST] = 0.0; }
    {
        const int32_t *src = b.work + (size_t)bc * b.cfg.work_stride;
        const double norm = ldexp(1.0, -(int)(b.cfg.bits_per_sample - 1u));
        for (uint32_t t = tid; t < na; t += LNB_RF_THREADS) X[t] = (double)src[t] * norm;
    }
    __syncthreads();

    const uint32_t L = b.cfg.num_layers;
    double *xin = X, *xout = Y;
    for (uint32_t l = 0; l < L; l++) {
        const uint32_t P = b.cfg.layer_params[l];
        /* ---- unit-count search with plain Levinson coefficients (linne_network.c:268-347) ---- */
        for (uint32_t cell = tid; cell < LNB_MAX_LEVELS * 256u; cell += LNB_RF_THREADS) {
            const uint32_t lv = cell / 256u, r = cell % 256u;
            if (!((1u << lv) <= P && (1u << lv) <= LNB_MAX_UNITS && (na % (1u << lv)) == 0u)) continue;
            const uint32_t U = 1u << lv, p = P / U, m = na / U;
            if (r >= U * (p + 1u)) continue;
            const uint32_t u = r / (p + 1u), lag = r % (p + 1u);
            sm.acorr[cell] = lnb_acorr_lag(xin + (size_t)u * m, m, lag, b.welch[(size_t)blk_i * LNB_MAX_LEVELS + lv]);
        }
        __syncthreads();
        for (uint32_t id = tid + 1u; id <= 255u; id += LNB_RF_THREADS) {
            const uint32_t lv = 31u - lnb_clz32(id), u = id - (1u << lv);
            if (!((1u << lv) <= P && (na % (1u << lv)) == 0u)) continue;
            const uint32_t U = 1u << lv, p = P / U, m = na / U;
            lnb_solve_unit(sm.acorr + lv * 256u + u * (p + 1u), p, m, lambda, sm.cand + lv * LNB_MAX_PARAMS + u * p);
        }
        __syncthreads();
        uint32_t best = 0;
        {
            double best_loss = (double)FLT_MAX;
            for (uint32_t lv = 0; lv < LNB_MAX_LEVELS; lv++) {
                if (!((1u << lv) <= P && (1u << lv) <= LNB_MAX_UNITS && (na % (1u << lv)) == 0u)) continue;
                const uint32_t U = 1u << lv, p = P / U, m = na / U;
                double part = 0.0;
                for (uint32_t t = (tid == 0) ? LNB_RF_THREADS : tid; t < na; t += LNB_RF_THREADS)      /* t = 0 is not counted */
                    part += fabs(lnb_rf_residual(xin, t, m, p, sm.cand + lv * LNB_MAX_PARAMS, xin[t]));
                const double loss = lnb_rf_sum(part, sm.red) / (double)na;
                if (loss < best_loss) { best_loss = loss; best = lv; }
            }
        }
        const uint32_t U = 1u << best, p = P / U, m = na / U;
        if (tid == 0) sm.log2u[l] = best;
        for (uint32_t k = tid; k < P; k += LNB_RF_THREADS) sm.w[l][k] = sm.cand[best * LNB_MAX_PARAMS + k];
        __syncthreads();

        /* ---- IRLS refinement of the chosen units (linne_network.c:350-376 with lpc.c:578-633) ---- */
        if (af_iters > 0u && m > p) {
            for (uint32_t u = 0; u < U; u++) {
                const double *xs = xin + (size_t)u * m;
                double r0 = lnb_acorr_lag(xs, m, 0, b.welch[(size_t)blk_i * LNB_MAX_LEVELS + best]);      /* cheap re-evaluation: uniform across threads */
                r0 = lnb_mul_rn(r0, lnb_add_rn(1.0, lambda));
                if (fabs(r0) < (double)FLT_EPSILON) continue;     /* silent unit keeps all-zero coefficients (lpc.c:597-602) */
                for (uint32_t i = tid; i < p; i += LNB_RF_THREADS) sm.avec[i] = sm.w[l][u * p + (p - 1u - i)];   /* un-reverse */
                __syncthreads();
                lnb_rf_irls_unit(sm, xs, m, p, af_iters, xout, b.sig_a + slot * b.cfg.work_stride);
                for (uint32_t i = tid; i < p; i += LNB_RF_THREADS) sm.w[l][u * p + (p - 1u - i)] = sm.avec[i];
                __syncthreads();
            }
        }
        /* ---- forward (linne_network.c:165-210) ---- */
        for (uint32_t t = tid; t < na; t += LNB_RF_THREADS)
            xout[t] = (t == 0u) ? xin[0] : xin[t] + lnb_rf_residual(xin, t, m, p, sm.w[l], 0.0);
        __syncthreads();
        double *tmp = xin; xin = xout; xout = tmp;
    }

    /* ---- momentum SGD on the whole cascade (linne_network.c:805-873) ---- */
    if (learning) {
        const double lr = 0.1f, alpha = 0.8f, eps = 1.0e-7;
        double *scr = train_scratch + (size_t)bc * (2u * LNB_MAX_LAYERS + 1u) * b.cfg.work_stride;
        double *xorig = scr;                                    /* normalised input */
        double *din[LNB_MAX_LAYERS], *dout[LNB_MAX_LAYERS];
        for (uint32_t l = 0; l < L; l++) { din[l] = scr + (size_t)(1u + 2u * l) * b.cfg.work_stride; dout[l] = din[l] + b.cfg.work_stride; }
        {
            const int32_t *src = b.work + (size_t)bc * b.cfg.work_stride;
            const double norm = ldexp(1.0, -(int)(b.cfg.bits_per_sample - 1u));
            for (uint32_t t = tid; t < na; t += LNB_RF_THREADS) xorig[t] = (double)src[t] * norm;
        }
        for (uint32_t l = 0; l < L; l++) for (uint32_t k = tid; k < LNB_MAX_PARAMS; k += LNB_RF_THREADS) sm.mom[l][k] = 0.0;
        __syncthreads();
        double prev_loss = (double)FLT_MAX;
        double *buf = X, *tmpb = Y;                              /* buf: running signal in shared memory */
        for (uint32_t itr = 0; itr < 2000u; itr++) {
            for (uint32_t t = tid; t < na; t += LNB_RF_THREADS) buf[t] = xorig[t];
            __syncthreads();
            /* forward through all layers, remembering each layer's input */
            for (uint32_t l = 0; l < L; l++) {
                const uint32_t P = b.cfg.layer_params[l], U = 1u << sm.log2u[l], p = P / U, m = na / U;
                for (uint32_t t = tid; t < na; t += LNB_RF_THREADS) {
                    din[l][t] = buf[t];
                    tmpb[t] = (t == 0u) ? buf[0] : buf[t] + lnb_rf_residual(buf, t, m, p, sm.w[l], 0.0);
                }
                __syncthreads();
                double *sw = buf; buf = tmpb; tmpb = sw;
            }
            double part = 0.0;
            for (uint32_t t = tid; t < na; t += LNB_RF_THREADS) part += fabs(buf[t]);
            const double loss = lnb_rf_sum(part, sm.red) / (double)na;
            /* gradient of the mean absolute value (linne_network.c:66-75) */
            for (uint32_t t = tid; t < na; t += LNB_RF_THREADS) {
                const double d = buf[t];
                buf[t] = (double)((d > 0.0) - (d < 0.0)) / (double)na;
            }
            __syncthreads();
            /* backward (linne_network.c:213-265) */
            for (int32_t l = (int32_t)L - 1; l >= 0; l--) {
                const uint32_t P = b.cfg.layer_params[l], U = 1u << sm.log2u[l], p = P / U, m = na / U;
                for (uint32_t t = tid; t < na; t += LNB_RF_THREADS) dout[l][t] = buf[t];
                __syncthreads();
                /* dparams[u][i] = sum_{j < m-p+i} din[j] * dout[p-i+j] */
                for (uint32_t k = tid; k < P; k += LNB_RF_THREADS) {
                    const uint32_t u = k / p, i = k % p;
                    const double *pin = din[l] + (size_t)u * m, *pout = dout[l] + (size_t)u * m;
                    double s = 0.0;
                    for (uint32_t j = 0; j + p < m + i; j++) s += pin[j] * pout[p - i + j];
                    sm.dw[l][k] = s;
                }
                /* back-propagated signal: back[t] += (sum_j w[j] * dout[p+t-j], terms inside the unit) / p */
                for (uint32_t t = tid; t < U * m; t += LNB_RF_THREADS) {
                    const uint32_t u = t / m, tl = t % m;
                    const double *pout = dout[l] + (size_t)u * m;
                    const double *w = sm.w[l] + u * p;
                    double s = 0.0;
                    for (uint32_t j = 0; j < p; j++) if (p + tl - j < m) s += w[j] * pout[p + tl - j];
                    tmpb[t] = buf[t] + s / (double)p;
                }
                for (uint32_t t = U * m + tid; t < na; t += LNB_RF_THREADS) tmpb[t] = buf[t];
                __syncthreads();
                double *sw = buf; buf = tmpb; tmpb = sw;
            }
            for (uint32_t l = 0; l < L; l++)
                for (uint32_t k = tid; k < b.cfg.layer_params[l]; k += LNB_RF_THREADS) {
                    sm.mom[l][k] = alpha * sm.mom[l][k] + lr * sm.dw[l][k];
                    sm.w[l][k] -= sm.mom[l][k];
                }
            __syncthreads();
            if (fabs(loss - prev_loss) < eps) break;
            prev_loss = loss;
        }
    }

    /* hand the result to the finish stage through the winning slot */
    for (uint32_t l = 0; l < L; l++) {
        if (tid == 0) b.chosen_log2u[slot * LNB_MAX_LAYERS + l] = (uint8_t)sm.log2u[l];
        double *dst = b.chosen_w + (slot * LNB_MAX_LAYERS + l) * LNB_MAX_PARAMS;
        for (uint32_t k = tid; k < b.cfg.layer_params[l]; k += LNB_RF_THREADS) dst[k] = sm.w[l][k];
    }
}
