/* lnb_analyze_v3.cuh -- cooperative analysis kernel: one CTA per slot (block, channel, regulariser).
 *
 * Runs the whole layer cascade of one slot with the signal resident in shared memory.  Covers
 * reference rows a7-a13: libs/linne_network/src/linne_network.c:268-347
 * (unit-count search), :350-376 (set parameter), :165-210 (forward), :50-63 (L1 loss),
 * libs/lpc/src/lpc.c:196-205 (Welch window), :215-249 (autocorrelation), :252-324 (Levinson-Durbin),
 * :327-366 (regularised solve).  Blocks whose analysis length is a multiple of 1024 and at most 10240
 * samples (the CLI default block) take this kernel; other shapes (in practice a file's tail block)
 * keep the flat kernels of lnb_pipeline.cuh.
 *
 * Design (what bounds it: the FP64 pipe, SURVEY 8d -- everything here exists to keep DFMA issuing)
 *   - Layout.  A signal is stored in GROUPS of 8 samples padded to 9 doubles: sample i lives at
 *     FRONT + i + (i >> 3).  A thread always works on whole groups and lanes of a warp on consecutive
 *     groups, so lanes are 9 doubles apart: every LDS.64 is bank-conflict free, and because all tile
 *     shapes are multiples of 8 every window element sits at a COMPILE-TIME offset from the group base
 *     (no cursor arithmetic, no bounds predicates: the arrays carry zeroed pads in front and behind).
 *   - Autocorrelation.  Work item = (unit, group of 16 lags[, sample split]) per warp.  A lane takes
 *     8 samples x L lags (L = 16, or 17 for the group holding lag p): 8 + L+7 loads feed 8*L DFMA.
 *     Lane partials are reduced once per work item (xor butterfly, fixed order => deterministic).
 *     Window elements past the end of a unit are masked only in the warp iterations that reach it.
 *   - Residual evaluation / forward filter.  A lane takes 8 consecutive outputs and slides an
 *     8-tap x 15-sample register window over the unit's taps: 8 + 8 loads feed 64 DFMA.
 *   - Levinson-Durbin.  Orders <= 16: one thread per unit in the reference's operation order.
 *     Orders 32..128: one warp per unit, coefficients in registers, the update of step k fused with the
 *     dot product of step k+1, reciprocal of the error off the critical path.
 *   - Signal loops use fused multiply-add (results differ from the unfused CPU reference in the last
 *     bits only; the 8-bit quantiser absorbs that -- tests/test_gpu_parity.py checks byte identity).
 */
#pragma once
#include "lnb_common.cuh"
#include "lnb_encode_core.cuh"

#define LNB_A3_THREADS  256
#define LNB_A3_WARPS    8
#define LNB_A3_MAX_NA   10240
#define LNB_A3_FRONT    160          /* zeroed doubles in front of sample 0 (>= 9 * 128 / 8 + window) */
#define LNB_A3_BACK     176          /* zeroed doubles behind the last sample (>= 9 * (128 + 24) / 8) */
#define LNB_A3_PART     24           /* doubles per work-item partial (>= 17 lags) */
#define LNB_A3_ITEMS    64           /* work items whose partials are kept when a task's samples are split */
#define LNB_A3_MIRROR   (LNB_MAX_PARAMS + 8)   /* doubles per Levinson mirror buffer; a warp task uses two */

__host__ __device__ inline uint32_t lnb_a3_array_doubles(uint32_t na_max)
{
    if (na_max < 4096u) na_max = 4096u;            /* the sample area doubles as solver scratch: 8 warps x 3 x 136 doubles */
    return LNB_A3_FRONT + na_max / 8u * 9u + LNB_A3_BACK;
}
__host__ __device__ inline size_t lnb_a3_smem_doubles(uint32_t na_max)
{
    return (size_t)2 * lnb_a3_array_doubles(na_max)
         + (size_t)LNB_MAX_LEVELS * LNB_MAX_PARAMS        /* candidate coefficients per level */
         + (size_t)LNB_MAX_LEVELS * 256                   /* autocorrelations per level: U*(p+1) = P+U <= 256 */
         + (size_t)LNB_A3_ITEMS * LNB_A3_PART             /* work-item partials */
         + (size_t)3 * LNB_A3_MIRROR                      /* scratch of the level-0 Toeplitz solve (runs beside the rest) */
         + 128;                                           /* level losses, block-sum scratch, per-warp level partials */
}

struct LnbA3Ctx {
    double *A, *B;               /* point at sample 0 (front pad lies below) */
    double *cand, *acorr, *part, *lev0, *misc;
    uint32_t na, ng;             /* samples, groups of 8 */
};

/* A team = the warps that run a phase together: the whole CTA (barrier 0 = __syncthreads) or warps 0..6
 * (named barrier 1) while warp 7 runs the long level-0 Levinson recursion beside them. */
struct LnbA3Team {
    uint32_t tid, nthr, warp, nwarps, bar;
    __device__ __forceinline__ void sync() const
    {
        if (bar == 0u) __syncthreads();
        else asm volatile("bar.sync %0, %1;" :: "r"(bar), "r"(nthr) : "memory");
    }
};

/* ---- Welch-windowed copy A -> B for unit length m (lpc.c:196-205), one group of 8 at a time ----
 * w = (scale * q) * (m-1-q) with q = min(pos, m-1-pos): positions are exact in double, so they are stepped by
 * additions from one conversion per group, and min / max are resolved once per group (a group crosses the
 * window's centre at most once). */
__device__ __forceinline__ void lnb_a3_window(const LnbA3Ctx &cx, const LnbA3Team &tm, uint32_t m, double scale)
{
    const uint32_t mg = m >> 3;
    for (uint32_t G = tm.tid; G < cx.ng; G += tm.nthr) {
        const uint32_t pos0 = (G % mg) * 8u;
        const double *src = cx.A + G * 9u;
        double *dst = cx.B + G * 9u;
        const uint32_t back0 = m - 1u - pos0;                        /* m-1-pos of the group's first sample */
        const double a0 = (double)pos0, b0 = (double)back0;
        if (pos0 + 7u <= back0 - 7u || pos0 >= back0) {              /* whole group on one side of the centre */
            const bool rising = pos0 < back0;
#pragma unroll
            for (uint32_t e = 0; e < 8u; e++) {
                const double a = a0 + (double)e, b = b0 - (double)e;
                const double w = rising ? __dmul_rn(__dmul_rn(scale, a), b) : __dmul_rn(__dmul_rn(scale, b), a);
                dst[e] = __dmul_rn(src[e], w);
            }
        } else {
#pragma unroll
            for (uint32_t e = 0; e < 8u; e++) {
                const double a = a0 + (double)e, b = b0 - (double)e;
                const double q = (a < b) ? a : b, r = (a < b) ? b : a;
                dst[e] = __dmul_rn(src[e], __dmul_rn(__dmul_rn(scale, q), r));
            }
        }
    }
}

/* ---- one 8-sample x L-lag tile: acc[k] += sum_o B[8g+o] * B[8g+o+k0+k] ----
 * Bg = address of the group's first sample, qoff = padded offset of lag k0 (= 18 * k0/16),
 * lim = number of window elements (from lag k0 of sample 0 on) that still lie inside the unit. */
template <int L, bool MASK>
__device__ __forceinline__ void lnb_a3_tile(const double *Bg, uint32_t qoff, int32_t lim, double (&acc)[L])
{
    double s[8], wv[L + 7];
#pragma unroll
    for (int o = 0; o < 8; o++) s[o] = Bg[o];
#pragma unroll
    for (int t = 0; t < L + 7; t++) {
        const double v = (!MASK || t < lim) ? Bg[qoff + t + (t >> 3)] : 0.0;
        wv[t] = v;
    }
#pragma unroll
    for (int o = 0; o < 8; o++)
#pragma unroll
        for (int k = 0; k < L; k++) acc[k] = fma(s[o], wv[o + k], acc[k]);
}

/* one work item: lags [16q, 16q+L) of unit u over the groups [g_lo, g_hi) of that unit */
template <int L>
__device__ __forceinline__ void lnb_a3_autocorr_item(const LnbA3Ctx &cx, uint32_t u, uint32_t q, uint32_t m, bool last_unit,
                                                     uint32_t g_lo, uint32_t g_hi, double *out /* L sums, lane 0 writes */)
{
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t mg = m >> 3;
    double acc[L];
#pragma unroll
    for (int k = 0; k < L; k++) acc[k] = 0.0;
    const double *Bu = cx.B + (size_t)u * mg * 9u;
    const uint32_t qoff = 18u * q;
    for (uint32_t g0 = g_lo; g0 < g_hi; g0 += 32u) {                /* warp-uniform trip count */
        const uint32_t g = g0 + lane;
        const bool active = g < g_hi;
        const uint32_t gg = active ? g : g_lo;
        const int32_t lim = (int32_t)m - (int32_t)(8u * gg + 16u * q);
        const bool need_mask = !last_unit && lim < L + 7;
        if (__any_sync(0xffffffffu, need_mask)) {
            double tmp[L];
#pragma unroll
            for (int k = 0; k < L; k++) tmp[k] = 0.0;
            lnb_a3_tile<L, true>(Bu + gg * 9u, qoff, last_unit ? (int32_t)(L + 7) : lim, tmp);
            if (active) {
#pragma unroll
                for (int k = 0; k < L; k++) acc[k] += tmp[k];
            }
        } else if (active) {
            lnb_a3_tile<L, false>(Bu + gg * 9u, qoff, 0, acc);
        }
    }
    /* lane partials -> sums, fixed order.  The first N = 2^n lags are reduced by a transposing butterfly: at each
     * of the first n steps a lane hands half of its values to its partner and keeps the other half, so 2N-ish
     * exchanges replace 5N; lane l ends up with lag (l >> (5-n)).  A seventeenth / ninth / .. lag takes the plain
     * butterfly. */
    constexpr int N = (L >= 16) ? 16 : (L >= 8) ? 8 : (L >= 4) ? 4 : 2;
    constexpr int LOGN = (N == 16) ? 4 : (N == 8) ? 3 : (N == 4) ? 2 : 1;
    {
        int cnt = N;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            if (cnt > 1) {
                const bool upper = (lane & (uint32_t)off) != 0u;
                const int half = cnt / 2;
#pragma unroll
                for (int k = 0; k < N / 2; k++) {
                    if (k < half) {
                        const double send = upper ? acc[k] : acc[k + half];
                        const double keep = upper ? acc[k + half] : acc[k];
                        acc[k] = keep + __shfl_xor_sync(0xffffffffu, send, off);
                    }
                }
                cnt = half;
            } else {
                acc[0] += __shfl_xor_sync(0xffffffffu, acc[0], off);
            }
        }
        if ((lane & ((1u << (5 - LOGN)) - 1u)) == 0u) out[lane >> (5 - LOGN)] = acc[0];
    }
    if (L > N) {
        double v = acc[L - 1];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if (lane == 0) out[L - 1] = v;
    }
}

/* ---- autocorrelation of every unit of one level: r[u][0..p] into acorr_lvl[u*(p+1) + lag] ----
 * Tasks (unit, lag group) are dealt to the team's warps; when they do not deal out evenly a task's samples
 * are split S ways and the partials summed afterwards in a fixed order. */
__device__ void lnb_a3_autocorr(const LnbA3Ctx &cx, const LnbA3Team &tm, uint32_t U, uint32_t p, double *acorr_lvl)
{
    const uint32_t m = cx.na / U, mg = m >> 3;
    const uint32_t nq = (p >= 16u) ? p / 16u : 1u;                   /* lag groups per unit */
    const uint32_t tasks = U * nq;
    uint32_t S = 1u;                                                 /* sample splits per task */
    if ((tasks % tm.nwarps) != 0u && tasks < 4u * tm.nwarps) S = (3u * tm.nwarps + tasks - 1u) / tasks;
    if (tasks * S > LNB_A3_ITEMS) S = LNB_A3_ITEMS / tasks;
    const uint32_t items = tasks * S;
    const uint32_t per = (mg + S - 1u) / S;
    for (uint32_t wi = tm.warp; wi < items; wi += tm.nwarps) {
        const uint32_t task = wi / S, split = wi % S;
        const uint32_t u = task / nq, q = task % nq;
        const uint32_t g_lo = split * per, g_hi = (g_lo + per < mg) ? g_lo + per : mg;
        const bool last_unit = (u + 1u == U);
        double *out = (S == 1u) ? acorr_lvl + u * (p + 1u) + 16u * q : cx.part + wi * LNB_A3_PART;
        if (p >= 16u) {
            if (q + 1u < nq) lnb_a3_autocorr_item<16>(cx, u, q, m, last_unit, g_lo, g_hi, out);
            else             lnb_a3_autocorr_item<17>(cx, u, q, m, last_unit, g_lo, g_hi, out);
        } else if (p == 8u)  lnb_a3_autocorr_item<9>(cx, u, q, m, last_unit, g_lo, g_hi, out);
        else if (p == 4u)    lnb_a3_autocorr_item<5>(cx, u, q, m, last_unit, g_lo, g_hi, out);
        else if (p == 2u)    lnb_a3_autocorr_item<3>(cx, u, q, m, last_unit, g_lo, g_hi, out);
        else                 lnb_a3_autocorr_item<2>(cx, u, q, m, last_unit, g_lo, g_hi, out);
    }
    if (S > 1u) {                                                    /* sum the sample splits in a fixed order */
        tm.sync();
        const uint32_t L_last = (p >= 16u) ? 17u : p + 1u;
        for (uint32_t i = tm.tid; i < tasks * 17u; i += tm.nthr) {
            const uint32_t task = i / 17u, k = i % 17u;
            const uint32_t u = task / nq, q = task % nq;
            const uint32_t Lq = (q + 1u == nq) ? L_last : 16u;
            if (k >= Lq) continue;
            double v = 0.0;
            for (uint32_t sp = 0; sp < S; sp++) v += cx.part[(task * S + sp) * LNB_A3_PART + k];
            acorr_lvl[u * (p + 1u) + 16u * q + k] = v;
        }
    }
}

/* reciprocal from a single-precision seed and three Newton steps: straight-line code (no slow-path call),
 * within an ulp or two of 1/x -- used only where the operation order already differs from the reference */
__device__ __forceinline__ double lnb_a3_rcp(double x)
{
    float seed;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(seed) : "f"(__double2float_rn(x)));
    double y = (double)seed;
    y = fma(y, fma(-x, y, 1.0), y);
    y = fma(y, fma(-x, y, 1.0), y);
    y = fma(y, fma(-x, y, 1.0), y);
    return y;
}

/* ---- regularised Toeplitz solve by one warp (orders 32..128); reversed coefficients into out_w[0..p) ----
 * Same system as the Levinson-Durbin recursion of lpc.c:252-324, computed in two dot-product-free passes so that
 * a step costs a handful of instructions instead of a 32-lane reduction:
 *   1. Schur (Le Roux-Gueguen) recursion for the reflection coefficients.  With F_m[i] = e^f_m[i+m] and
 *      B_m[i] = e^b_m[i+m]:  k_m = -F_{m-1}[1] / B_{m-1}[0],  B_m[i] = B_{m-1}[i] + k_m F_{m-1}[i+1],
 *      F_m[i] = F_{m-1}[i+1] + k_m B_{m-1}[i]  (B_m[0] is the prediction error).  B stays in registers (lane l
 *      holds i = l, l+32, ..), F is double-buffered in shared memory for the shifted read; the work shrinks by
 *      one element per step.
 *   2. step-up: a_m[i] = a_{m-1}[i] + k_m a_{m-1}[m-i], a_m[m] = k_m, the reversed read again from a
 *      double-buffered mirror.
 * `scr` = 3 * LNB_A3_MIRROR doubles (two buffers + the reflection coefficients). */
__device__ void lnb_a3_levinson_warp(const double *r_in, uint32_t p, double lambda, double *scr, double *out_w)
{
    const uint32_t lane = threadIdx.x & 31u;
    const double r0 = __dmul_rn(r_in[0], __dadd_rn(1.0, lambda));
    if (fabs(r0) < (double)FLT_EPSILON) {
        for (uint32_t i = lane; i < p; i += 32u) out_w[i] = 0.0;
        return;
    }
    double *cur = scr, *nxt = scr + LNB_A3_MIRROR, *gam = scr + 2 * LNB_A3_MIRROR;
    double B[5];
#pragma unroll
    for (int s = 0; s < 5; s++) {
        const uint32_t i = lane + 32u * (uint32_t)s;
        const double ri = (i <= p) ? r_in[i] : 0.0;
        B[s] = (i == 0u) ? r0 : ri;
        if (i < LNB_A3_MIRROR) { cur[i] = ri; nxt[i] = 0.0; }
    }
    __syncwarp();
    double rinv = lnb_a3_rcp(r0);
    for (uint32_t m = 1; m <= p; m++) {
        const double gamma = -cur[1] * rinv;
        if (lane == 0) gam[m] = gamma;
        const uint32_t cnt = p - m;                                  /* B_m[0..cnt], F_m[1..cnt] are what later steps read */
        const uint32_t ns = (cnt >> 5) + 1u;
#pragma unroll
        for (int s = 0; s < 5; s++) {
            if ((uint32_t)s < ns) {
                const uint32_t i = lane + 32u * (uint32_t)s;
                if (i <= cnt) {
                    const double fs = cur[i + 1u];
                    const double b = B[s];
                    B[s] = fma(gamma, fs, b);
                    nxt[i] = fma(gamma, b, fs);
                }
            }
        }
        const double err = __shfl_sync(0xffffffffu, B[0], 0);        /* B_m[0] */
        rinv = lnb_a3_rcp(err);
        __syncwarp();
        double *t = cur; cur = nxt; nxt = t;
    }
    /* step-up on the same buffers (now mirrors of a) */
    double a[5];
#pragma unroll
    for (int s = 0; s < 5; s++) {
        const uint32_t i = lane + 32u * (uint32_t)s;
        a[s] = (i == 0u) ? 1.0 : 0.0;
        if (i < LNB_A3_MIRROR) { cur[i] = a[s]; nxt[i] = a[s]; }
    }
    __syncwarp();
    for (uint32_t m = 1; m <= p; m++) {
        const double g = gam[m];
        const uint32_t ns = (m >> 5) + 1u;
#pragma unroll
        for (int s = 0; s < 5; s++) {
            if ((uint32_t)s < ns) {
                const uint32_t i = lane + 32u * (uint32_t)s;
                if (i >= 1u && i <= m) {
                    const double v = (i == m) ? g : fma(g, cur[m - i], a[s]);
                    a[s] = v;
                    nxt[i] = v;
                }
            }
        }
        __syncwarp();
        double *t = cur; cur = nxt; nxt = t;
    }
#pragma unroll
    for (int s = 0; s < 5; s++) {
        const uint32_t i = lane + 32u * (uint32_t)s;                 /* a[i], i = 1..p -> out_w[p - i] */
        if (i >= 1u && i <= p) out_w[p - i] = a[s];
    }
    __syncwarp();
}

/* eight taps of the sliding FIR tile: window = prev[1..7] ++ next[0..7] (next = the group at base + 7 .. + 15) */
__device__ __forceinline__ void lnb_a3_fir8(const double *base, const double *w, const double (&prev)[8], double (&next)[8],
                                            double (&acc)[8])
{
#pragma unroll
    for (int k = 0; k < 8; k++) next[k] = base[7 + k + ((7 + k) >> 3)];      /* window samples 7..14 (the pad slot is skipped) */
#pragma unroll
    for (int jj = 0; jj < 8; jj++) {
        const double wj = w[jj];
#pragma unroll
        for (int o = 0; o < 8; o++) {
            const int t = jj + o;
            acc[o] = fma(wj, (t < 7) ? prev[t + 1] : next[t - 7], acc[o]);
        }
    }
}

/* ---- FIR over whole groups: res[t] = init + sum_j w[j] * x[t-p+j]  (x[<0] = 0) ----
 * MODE 0: search loss (init = x[t], sample 0 not counted)          linne_network.c:318-335
 * MODE 1: forward (y = x[t] + sum, written to Y, all samples counted)   linne_network.c:183-208
 * PS = 0: p is a multiple of 8; PS = 1, 2, 4: p = PS. */
template <int MODE, int PS>
__device__ __forceinline__ double lnb_a3_fir(const LnbA3Ctx &cx, const LnbA3Team &tm, const double *X, double *Y, uint32_t p,
                                             uint32_t m, const double *cand_lvl)
{
    const uint32_t mg = m >> 3;
    double loss = 0.0;
    for (uint32_t G = tm.tid; G < cx.ng; G += tm.nthr) {
        const uint32_t u = G / mg;
        const double *w = cand_lvl + u * p;
        const double *Xg = X + G * 9u;
        double acc[8];
#pragma unroll
        for (int o = 0; o < 8; o++) acc[o] = (MODE == 0) ? Xg[o] : 0.0;
        if (PS == 0) {
            const double *base = Xg - (p >> 3) * 9u;                 /* sample 8G - p */
            /* sliding window of 15 samples = the last 7 of the previous group + the 8 of the next one, kept in two
             * register sets that swap roles every 8 taps (no register moves) */
            double P[8], Q[8];
#pragma unroll
            for (int k = 0; k < 7; k++) P[k + 1] = base[k];
            uint32_t j0 = 0;
            for (; j0 + 16u <= p; j0 += 16u) {
                lnb_a3_fir8(base, w + j0, P, Q, acc);
                base += 9;
                lnb_a3_fir8(base, w + j0 + 8u, Q, P, acc);
                base += 9;
            }
            if (j0 < p) lnb_a3_fir8(base, w + j0, P, Q, acc);
        } else {
            double xw[PS + 7];
#pragma unroll
            for (int t = 0; t < PS + 7; t++) xw[t] = (t < PS) ? Xg[t - 1 - PS] : Xg[t - PS];
#pragma unroll
            for (int jj = 0; jj < PS; jj++) {
                const double wj = w[jj];
#pragma unroll
                for (int o = 0; o < 8; o++) acc[o] = fma(wj, xw[jj + o], acc[o]);
            }
        }
        if (MODE == 0) {
#pragma unroll
            for (int o = 0; o < 8; o++) if (!(G == 0u && o == 0)) loss += fabs(acc[o]);
        } else {
            double *Yg = Y + G * 9u;
#pragma unroll
            for (int o = 0; o < 8; o++) {
                const double y = Xg[o] + acc[o];
                Yg[o] = y;
                loss += fabs(y);
            }
        }
    }
    return loss;
}

template <int MODE>
__device__ double lnb_a3_fir_any(const LnbA3Ctx &cx, const LnbA3Team &tm, const double *X, double *Y, uint32_t p, uint32_t m,
                                 const double *cand_lvl)
{
    if (p >= 8u) return lnb_a3_fir<MODE, 0>(cx, tm, X, Y, p, m, cand_lvl);
    if (p == 4u) return lnb_a3_fir<MODE, 4>(cx, tm, X, Y, p, m, cand_lvl);
    if (p == 2u) return lnb_a3_fir<MODE, 2>(cx, tm, X, Y, p, m, cand_lvl);
    return lnb_a3_fir<MODE, 1>(cx, tm, X, Y, p, m, cand_lvl);
}

/* team-wide sum in a fixed order; result valid in every thread of the team */
__device__ double lnb_a3_team_sum(const LnbA3Team &tm, double v, double *scratch /* >= 8 doubles */)
{
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    tm.sync();
    if ((tm.tid & 31u) == 0) scratch[tm.warp] = v;
    tm.sync();
    double s = 0.0;
    for (uint32_t w = 0; w < tm.nwarps; w++) s += scratch[w];
    tm.sync();
    return s;
}
__device__ double lnb_a3_block_sum(double v, double *scratch)
{
    const LnbA3Team all = { threadIdx.x, LNB_A3_THREADS, threadIdx.x >> 5, LNB_A3_WARPS, 0u };
    return lnb_a3_team_sum(all, v, scratch);
}

/* L1 loss of the levels [lv_first, nlev) (linne_network.c:318-335) by a team, mean losses into cx.misc[lv].
 * Every warp leaves its partial of every level in shared memory and ONE team barrier precedes the final sums
 * (same summation order as a per-level reduction, a fraction of the barriers).  The caller synchronises after. */
__device__ void lnb_a3_level_losses(const LnbA3Ctx &cx, const LnbA3Team &tm, uint32_t P, uint32_t lv_first, uint32_t nlev)
{
    double *parts = cx.misc + 32;                                    /* [warp][LNB_MAX_LEVELS] */
    for (uint32_t lv = lv_first; lv < nlev; lv++) {
        const uint32_t U = 1u << lv, p = P / U;
        double v = lnb_a3_fir_any<0>(cx, tm, cx.A, (double *)0, p, cx.na / U, cx.cand + lv * LNB_MAX_PARAMS);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if ((tm.tid & 31u) == 0u) parts[tm.warp * LNB_MAX_LEVELS + lv] = v;
    }
    tm.sync();
    if (tm.tid >= lv_first && tm.tid < nlev) {
        double sum = 0.0;
        for (uint32_t w = 0; w < tm.nwarps; w++) sum += parts[w * LNB_MAX_LEVELS + tm.tid];
        cx.misc[tm.tid] = sum / (double)cx.na;
    }
}

/* Levinson-Durbin of the levels [lv_first, nlev) by a team: threads take the orders <= 16 (reference operation
 * order), warps the orders 32..128.  `scratch` = LNB_A3_WARPS x 3 x LNB_A3_MIRROR doubles (only used by the warp tasks). */
__device__ void lnb_a3_solve_levels(const LnbA3Ctx &cx, const LnbA3Team &tm, uint32_t P, uint32_t lv_first, uint32_t nlev,
                                    double lambda, double *scratch)
{
    uint32_t thread_tasks = 0, warp_tasks = 0;
    for (uint32_t lv = lv_first; lv < nlev; lv++) { if ((P >> lv) <= 16u) thread_tasks += 1u << lv; else warp_tasks += 1u << lv; }
    for (uint32_t t = tm.tid; t < thread_tasks; t += tm.nthr) {
        uint32_t rest = t;
        for (uint32_t lv = lv_first; lv < nlev; lv++) {
            const uint32_t U = 1u << lv, p = P / U;
            if (p > 16u) continue;
            if (rest < U) {
                double r[17], a[18], coef[16];
                const double *src = cx.acorr + lv * 256 + rest * (p + 1u);
                for (uint32_t k = 0; k <= p; k++) r[k] = src[k];
                r[0] = __dmul_rn(r[0], __dadd_rn(1.0, lambda));
                lnb_levinson(r, p, a, coef, (double *)0);
                double *dst = cx.cand + lv * LNB_MAX_PARAMS + rest * p;
                for (uint32_t j = 0; j < p; j++) dst[j] = coef[p - 1u - j];
                break;
            }
            rest -= U;
        }
    }
    for (uint32_t t = tm.warp; t < warp_tasks; t += tm.nwarps) {
        uint32_t rest = t;
        for (uint32_t lv = lv_first; lv < nlev; lv++) {
            const uint32_t U = 1u << lv, p = P / U;
            if (p <= 16u) continue;
            if (rest < U) {
                lnb_a3_levinson_warp(cx.acorr + lv * 256 + rest * (p + 1u), p, lambda,
                                     scratch + tm.warp * 3 * LNB_A3_MIRROR, cx.cand + lv * LNB_MAX_PARAMS + rest * p);
                break;
            }
            rest -= U;
        }
    }
}

/* ---- any other shape up to LNB_A3_MAX_NA samples (a file's tail block): same cascade, plain layout ----
 * One thread per autocorrelation cell / sample, work-item bodies shared with the flat kernels
 * (lnb_encode_core.cuh), so the tail costs one CTA of the same launch instead of a chain of small ones. */
__device__ void lnb_a3_generic(const LnbEncodeBatch &b, uint32_t s, uint32_t bc, uint32_t blk_i, uint32_t na,
                               double lambda, uint32_t na_max, double *smem)
{
    const uint32_t c = threadIdx.x, warp = c >> 5;
    const uint32_t arr = lnb_a3_array_doubles(na_max);
    double *A = smem, *B = smem + arr;
    double *cand = smem + 2u * arr, *acorr = cand + LNB_MAX_LEVELS * LNB_MAX_PARAMS;
    double *misc = acorr + LNB_MAX_LEVELS * 256 + LNB_A3_ITEMS * LNB_A3_PART + 3 * LNB_A3_MIRROR;
    {
        const int32_t *src = b.work + (size_t)bc * b.cfg.work_stride;
        const double norm = ldexp(1.0, -(int)(b.cfg.bits_per_sample - 1u));
        for (uint32_t i = c; i < na; i += LNB_A3_THREADS) A[i] = (double)src[i] * norm;
    }
    __syncthreads();
    double final_loss = 0.0;
    for (uint32_t l = 0; l < b.cfg.num_layers; l++) {
        const uint32_t P = b.cfg.layer_params[l];
        for (uint32_t lv = 0; lv < LNB_MAX_LEVELS; lv++) {           /* lpc.c:196-249 */
            if (!lnb_level_valid(lv, P, na)) continue;
            const uint32_t U = 1u << lv, p = P / U, m = na / U;
            const double scale = b.welch[(size_t)blk_i * LNB_MAX_LEVELS + lv];
            for (uint32_t i = c; i < na; i += LNB_A3_THREADS)        /* windowed copy (same products as lnb_acorr_lag) */
                B[i] = lnb_mul_rn(A[i], lnb_welch_weight(scale, i % m, m));
            __syncthreads();
            for (uint32_t cell = c; cell < U * (p + 1u); cell += LNB_A3_THREADS) {
                const uint32_t lag = cell % (p + 1u);
                const double *xs = B + (size_t)(cell / (p + 1u)) * m;
                double sum = 0.0;
                for (uint32_t i = 0; i + lag < m; i++) sum = lnb_mac(xs[i], xs[i + lag], sum);
                acorr[lv * 256u + cell] = sum;
            }
            __syncthreads();
        }
        {   /* regularised solve: threads take the short orders, warps the long ones (B is free: mirror scratch) */
            uint32_t task = c, wt = warp;
            for (uint32_t lv = 0; lv < LNB_MAX_LEVELS; lv++) {
                if (!lnb_level_valid(lv, P, na)) continue;
                const uint32_t U = 1u << lv, p = P / U, m = na / U;
                if (p <= 16u) {
                    if (task < U) {
                        double r[17], a[18], coef[16];
                        const double *src = acorr + lv * 256u + task * (p + 1u);
                        for (uint32_t k = 0; k <= p; k++) r[k] = src[k];
                        r[0] = __dmul_rn(r[0], __dadd_rn(1.0, lambda));
                        if (m < p) { for (uint32_t k = 0; k < p; k++) coef[k] = 0.0; }
                        else lnb_levinson(r, p, a, coef, (double *)0);
                        double *dst = cand + lv * LNB_MAX_PARAMS + task * p;
                        for (uint32_t j = 0; j < p; j++) dst[j] = coef[p - 1u - j];
                        task = 0xFFFFFFFFu;
                    } else if (task != 0xFFFFFFFFu) task -= U;
                } else {
                    if (wt < U) {
                        double *dst = cand + lv * LNB_MAX_PARAMS + wt * p;
                        if (m < p) { for (uint32_t j = c & 31u; j < p; j += 32u) dst[j] = 0.0; }
                        else lnb_a3_levinson_warp(acorr + lv * 256u + wt * (p + 1u), p, lambda, B + warp * 3 * LNB_A3_MIRROR, dst);
                        wt = 0xFFFFFFFFu;
                    } else if (wt != 0xFFFFFFFFu) wt -= U;
                }
            }
        }
        __syncthreads();
        uint32_t best = 0;
        {   /* L1 loss of every valid level, first minimum wins (linne_network.c:318-341) */
            double best_loss = (double)FLT_MAX;
            for (uint32_t lv = 0; lv < LNB_MAX_LEVELS; lv++) {
                if (!lnb_level_valid(lv, P, na)) continue;
                const uint32_t U = 1u << lv, p = P / U, m = na / U;
                double part = 0.0;
                for (uint32_t t = c; t < na; t += LNB_A3_THREADS)
                    if (t) part += fabs(lnb_residual_at(A, t, m, p, cand + lv * LNB_MAX_PARAMS, A[t]));
                const double loss = lnb_a3_block_sum(part, misc + 16) / (double)na;
                if (loss < best_loss) { best_loss = loss; best = lv; }
            }
        }
        if (c == 0) b.chosen_log2u[(size_t)s * LNB_MAX_LAYERS + l] = (uint8_t)best;
        {
            double *dst = b.chosen_w + ((size_t)s * LNB_MAX_LAYERS + l) * LNB_MAX_PARAMS;
            for (uint32_t k = c; k < P; k += LNB_A3_THREADS) dst[k] = cand[best * LNB_MAX_PARAMS + k];
        }
        {   /* forward (linne_network.c:165-210) */
            const uint32_t U = 1u << best, p = P / U, m = na / U;
            double part = 0.0;
            for (uint32_t t = c; t < na; t += LNB_A3_THREADS) {
                const double v = (t == 0u) ? A[0] : A[t] + lnb_residual_at(A, t, m, p, cand + best * LNB_MAX_PARAMS, 0.0);
                B[t] = v;
                part += fabs(v);
            }
            final_loss = lnb_a3_block_sum(part, misc + 16);
        }
        __syncthreads();
        double *t = A; A = B; B = t;
    }
    const uint32_t chunks_per_slot = (b.cfg.work_stride + 63u) / 64u, nch = (na + 63u) / 64u;
    for (uint32_t i = c; i < nch; i += LNB_A3_THREADS) b.final_sum[(size_t)s * chunks_per_slot + i] = (i == 0) ? final_loss : 0.0;
}

__global__ void __launch_bounds__(LNB_A3_THREADS, 1) lnb_analyze_v3_kernel(LnbEncodeBatch b, uint32_t na_max)
{
    extern __shared__ __align__(16) double lnb_a3_smem[];
    const uint32_t s = blockIdx.x, c = threadIdx.x;
    const uint32_t bc = s / b.cfg.num_lambdas, lam = s % b.cfg.num_lambdas;
    const uint32_t blk_i = bc / b.cfg.num_channels;
    const LnbBlockDesc blk = b.blocks[blk_i];
    if (blk.type != LNB_BLOCK_COMPRESSED || !(blk.status & (LNB_ENC_FLAG_FAST | LNB_ENC_FLAG_GENERIC))) return;
    if (!(blk.status & LNB_ENC_FLAG_FAST)) {
        lnb_a3_generic(b, s, bc, blk_i, blk.na, b.cfg.lambdas[lam], na_max, lnb_a3_smem);
        return;
    }

    const uint32_t arr = lnb_a3_array_doubles(na_max);
    LnbA3Ctx cx;
    cx.na = blk.na; cx.ng = blk.na >> 3;
    cx.A = lnb_a3_smem + LNB_A3_FRONT;
    cx.B = lnb_a3_smem + arr + LNB_A3_FRONT;
    cx.cand = lnb_a3_smem + 2u * arr;
    cx.acorr = cx.cand + LNB_MAX_LEVELS * LNB_MAX_PARAMS;
    cx.part = cx.acorr + LNB_MAX_LEVELS * 256;
    cx.lev0 = cx.part + LNB_A3_ITEMS * LNB_A3_PART;
    cx.misc = cx.lev0 + 3 * LNB_A3_MIRROR;
    const uint32_t na = cx.na;
    const double lambda = b.cfg.lambdas[lam];

    /* zero pads of both arrays (everything outside the samples), then the layer-0 input:
     * normalised work signal (linne_encoder.c:661-663) */
    {
        const uint32_t data = cx.ng * 9u;
        for (uint32_t i = c; i < LNB_A3_FRONT; i += LNB_A3_THREADS) { lnb_a3_smem[i] = 0.0; lnb_a3_smem[arr + i] = 0.0; }
        for (uint32_t i = LNB_A3_FRONT + data + c; i < arr; i += LNB_A3_THREADS) { lnb_a3_smem[i] = 0.0; lnb_a3_smem[arr + i] = 0.0; }
        const int32_t *src = b.work + (size_t)bc * b.cfg.work_stride;
        const double norm = ldexp(1.0, -(int)(b.cfg.bits_per_sample - 1u));
        for (uint32_t G = c; G < cx.ng; G += LNB_A3_THREADS) {
            const int4 v0 = *(const int4 *)(src + 8u * G), v1 = *(const int4 *)(src + 8u * G + 4u);
            double *dst = cx.A + G * 9u;
            dst[0] = (double)v0.x * norm; dst[1] = (double)v0.y * norm; dst[2] = (double)v0.z * norm; dst[3] = (double)v0.w * norm;
            dst[4] = (double)v1.x * norm; dst[5] = (double)v1.y * norm; dst[6] = (double)v1.z * norm; dst[7] = (double)v1.w * norm;
            dst[8] = 0.0;
            cx.B[G * 9u + 8u] = 0.0;
        }
    }
    __syncthreads();

    double final_loss = 0.0;
    for (uint32_t l = 0; l < b.cfg.num_layers; l++) {
        const uint32_t P = b.cfg.layer_params[l];
        uint32_t nlev = 0;
        while (nlev < LNB_MAX_LEVELS && (1u << nlev) <= P) nlev++;        /* U = 1 .. min(128, P) */

        const LnbA3Team all = { c, LNB_A3_THREADS, c >> 5, LNB_A3_WARPS, 0u };
        /* The order-P recursion of level 0 (one unit) is 127 dependent steps on ONE warp when P = 128.  It runs on
         * warp 7 while warps 0..6 do everything else of the search that does not depend on it. */
        const bool split = (P >= 32u) && nlev >= 2u;
        const double *welch = b.welch + (size_t)blk_i * LNB_MAX_LEVELS;

        if (split) {
            lnb_a3_window(cx, all, na, welch[0]);
            __syncthreads();
            lnb_a3_autocorr(cx, all, 1u, P, cx.acorr);
            __syncthreads();
            if ((c >> 5) == LNB_A3_WARPS - 1u) {
                lnb_a3_levinson_warp(cx.acorr, P, lambda, cx.lev0, cx.cand);
            } else {
                const LnbA3Team rest = { c, LNB_A3_THREADS - 32u, c >> 5, LNB_A3_WARPS - 1u, 1u };
                for (uint32_t lv = 1; lv < nlev; lv++) {
                    const uint32_t U = 1u << lv, p = P / U, m = na / U;
                    lnb_a3_window(cx, rest, m, welch[lv]);
                    rest.sync();
                    lnb_a3_autocorr(cx, rest, U, p, cx.acorr + lv * 256);
                    rest.sync();
                }
                lnb_a3_solve_levels(cx, rest, P, 1u, nlev, lambda, cx.B);       /* B is free again: mirror scratch */
                rest.sync();
                lnb_a3_level_losses(cx, rest, P, 1u, nlev);
            }
            __syncthreads();
            lnb_a3_level_losses(cx, all, P, 0u, 1u);
        } else {
            /* ---- autocorrelation of every unit of every level ---- */
            for (uint32_t lv = 0; lv < nlev; lv++) {
                const uint32_t U = 1u << lv, p = P / U, m = na / U;
                lnb_a3_window(cx, all, m, welch[lv]);
                __syncthreads();
                lnb_a3_autocorr(cx, all, U, p, cx.acorr + lv * 256);
                __syncthreads();
            }
            /* ---- Levinson-Durbin (B is free: mirror scratch) ---- */
            lnb_a3_solve_levels(cx, all, P, 0u, nlev, lambda, cx.B);
            __syncthreads();
            /* ---- L1 loss of every level (linne_network.c:318-335) ---- */
            lnb_a3_level_losses(cx, all, P, 0u, nlev);
        }
        __syncthreads();
        /* first minimum wins (linne_network.c:337-341) */
        uint32_t best = 0;
        {
            double best_loss = (double)FLT_MAX;
            for (uint32_t lv = 0; lv < nlev; lv++)
                if (cx.misc[lv] < best_loss) { best_loss = cx.misc[lv]; best = lv; }
        }
        if (c == 0) b.chosen_log2u[(size_t)s * LNB_MAX_LAYERS + l] = (uint8_t)best;
        {
            double *dst = b.chosen_w + ((size_t)s * LNB_MAX_LAYERS + l) * LNB_MAX_PARAMS;
            const double *src = cx.cand + best * LNB_MAX_PARAMS;
            for (uint32_t k = c; k < P; k += LNB_A3_THREADS) dst[k] = src[k];
        }

        /* ---- forward: this layer's residual becomes the next layer's input ---- */
        {
            const uint32_t U = 1u << best, p = P / U;
            const double part = lnb_a3_fir_any<1>(cx, all, cx.A, cx.B, p, na / U, cx.cand + best * LNB_MAX_PARAMS);
            final_loss = lnb_a3_block_sum(part, cx.misc + 16);
        }
        __syncthreads();
        double *t = cx.A; cx.A = cx.B; cx.B = t;
    }
    /* total |residual| of the cascade: what picks the regulariser (linne_network.c:618-626) */
    {   /* the finish stage sums ceil(na/64) chunk slots of this analysis slot: total in slot 0, zeros after */
        const uint32_t chunks_per_slot = (b.cfg.work_stride + 63u) / 64u, nch = (na + 63u) / 64u;
        for (uint32_t i = c; i < nch; i += LNB_A3_THREADS) b.final_sum[(size_t)s * chunks_per_slot + i] = (i == 0) ? final_loss : 0.0;
    }
}
