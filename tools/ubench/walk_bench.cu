// walk_bench.cu -- micro-benchmarks of the serial code-word walk (recursive-Rice lengths) on one lane of one warp.
// Measures cycles per code word of several formulations of the same chain, on a synthetic payload held in shared
// memory.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o walk_bench walk_bench.cu ; run on a B200.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

#define WORDS 8192            // 32 KB of payload in shared memory
#define NCW   20000

__device__ __forceinline__ uint32_t bfind(uint32_t x) { uint32_t r; asm("bfind.u32 %0, %1;" : "=r"(r) : "r"(x)); return r; }
__device__ __forceinline__ void sts32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t lds32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }

struct Result { long long cycles; uint32_t endpos; uint32_t check; };

// ---- variant 0: hi/lo/nw register window, FLO, branchy refill (what lnb_stream_v2 does) ----
template <int VAR>
__global__ void walk(const uint32_t *g, uint32_t k2, uint32_t ncw, Result *res, uint32_t *sink)
{
    extern __shared__ uint32_t sm[];        // [WORDS] payload (already big-endian corrected) + [NCW] output
    for (uint32_t i = threadIdx.x; i < WORDS; i += blockDim.x) sm[i] = g[i];
    __syncthreads();
    if (threadIdx.x != 0) return;
    const uint32_t ring = (uint32_t)__cvta_generic_to_shared(sm);
    const uint32_t out = ring + WORDS * 4;
    uint32_t hi = sm[0], lo = sm[1], nw = sm[2], wi = 2, cnt = 64, check = 0;
    const uint32_t K = k2 + 32u;
    long long t0 = clock64();
    if (VAR == 0) {
        for (uint32_t i = 0; i < ncw; i += 8) {
#pragma unroll
            for (int s = 0; s < 8; s++) {
                const uint32_t f = bfind(hi);
                if ((int32_t)f < (int32_t)k2) { check++; }
                sts32(out + 4u * (i + s), hi);
                const uint32_t L = K - (f < 30u ? f : 30u);
                hi = __funnelshift_lc(lo, hi, L); lo = __funnelshift_lc(0u, lo, L); cnt -= L;
                if (cnt < 32u) { hi |= nw >> cnt; lo = __funnelshift_r(0u, nw, cnt); cnt += 32u; wi++; nw = lds32(ring + wi * 4u); }
            }
        }
    } else if (VAR == 1) {       // branch-free refill (selects)
        for (uint32_t i = 0; i < ncw; i += 8) {
#pragma unroll
            for (int s = 0; s < 8; s++) {
                const uint32_t f = bfind(hi);
                if ((int32_t)f < (int32_t)k2) { check++; }
                sts32(out + 4u * (i + s), hi);
                const uint32_t L = K - (f < 30u ? f : 30u);
                hi = __funnelshift_lc(lo, hi, L); lo = __funnelshift_lc(0u, lo, L); cnt -= L;
                const bool re = cnt < 32u;
                const uint32_t c31 = cnt & 31u;
                const uint32_t hi2 = hi | (nw >> c31), lo2 = __funnelshift_r(0u, nw, c31);
                hi = re ? hi2 : hi; lo = re ? lo2 : lo; cnt += re ? 32u : 0u; wi += re ? 1u : 0u;
                nw = lds32(ring + wi * 4u);
            }
        }
    } else if (VAR == 2) {       // position based: window re-read from shared memory for every code word
        uint32_t pos = 0;
        for (uint32_t i = 0; i < ncw; i += 8) {
#pragma unroll
            for (int s = 0; s < 8; s++) {
                const uint32_t a = ring + ((pos >> 5) << 2);
                const uint32_t h = __funnelshift_l(lds32(a + 4u), lds32(a), pos & 31u);
                const uint32_t f = bfind(h);
                if ((int32_t)f < (int32_t)k2) { check++; }
                sts32(out + 4u * (i + s), h);
                pos += K - (f < 30u ? f : 30u);
            }
        }
        wi = (pos >> 5) + 2; cnt = 64 - (pos & 31); hi = pos;
    } else if (VAR == 3) {       // no FLO: length from compares (lz <= 3 fast, else FLO)
        for (uint32_t i = 0; i < ncw; i += 8) {
#pragma unroll
            for (int s = 0; s < 8; s++) {
                uint32_t L = k2 + 2u + (hi < 0x40000000u) + (hi < 0x20000000u) + (hi < 0x10000000u);
                if (hi < 0x08000000u) { const uint32_t f = bfind(hi); if ((int32_t)f < (int32_t)k2) check++; L = K - f; }
                sts32(out + 4u * (i + s), hi);
                hi = __funnelshift_lc(lo, hi, L); lo = __funnelshift_lc(0u, lo, L); cnt -= L;
                if (cnt < 32u) { hi |= nw >> cnt; lo = __funnelshift_r(0u, nw, cnt); cnt += 32u; wi++; nw = lds32(ring + wi * 4u); }
            }
        }
    } else if (VAR == 4) {       // pairs: two short code words (lz <= 1) per step when both are
        uint32_t i = 0;
        const uint32_t s2 = k2 + 2u;
        while (i + 2 <= ncw) {
            const uint32_t h2 = __funnelshift_l(lo, hi, s2);
            if (hi >= 0x40000000u && h2 >= 0x40000000u && s2 <= 16u) {
                sts32(out + 4u * i, hi); sts32(out + 4u * i + 4u, h2);
                const uint32_t L = 2u * s2;
                hi = __funnelshift_lc(lo, hi, L); lo = __funnelshift_lc(0u, lo, L); cnt -= L;
                i += 2;
            } else {
                const uint32_t f = bfind(hi);
                if ((int32_t)f < (int32_t)k2) { check++; }
                sts32(out + 4u * i, hi);
                const uint32_t L = K - (f < 30u ? f : 30u);
                hi = __funnelshift_lc(lo, hi, L); lo = __funnelshift_lc(0u, lo, L); cnt -= L;
                i += 1;
            }
            if (cnt < 32u) { hi |= nw >> cnt; lo = __funnelshift_r(0u, nw, cnt); cnt += 32u; wi++; nw = lds32(ring + wi * 4u); }
        }
    } else if (VAR == 5) {       // chain only: no store, no long check (lower bound of the formulation)
        for (uint32_t i = 0; i < ncw; i += 8) {
#pragma unroll
            for (int s = 0; s < 8; s++) {
                const uint32_t f = bfind(hi);
                const uint32_t L = K - (f < 30u ? f : 30u);
                check += L;
                hi = __funnelshift_lc(lo, hi, L); lo = __funnelshift_lc(0u, lo, L); cnt -= L;
                if (cnt < 32u) { hi |= nw >> cnt; lo = __funnelshift_r(0u, nw, cnt); cnt += 32u; wi++; nw = lds32(ring + wi * 4u); }
            }
        }
    } else if (VAR == 6) {       // 64-bit window in one register pair, shifts as 64-bit ops, clzll
        unsigned long long w = ((unsigned long long)hi << 32) | lo;
        for (uint32_t i = 0; i < ncw; i += 8) {
#pragma unroll
            for (int s = 0; s < 8; s++) {
                const uint32_t h = (uint32_t)(w >> 32);
                const uint32_t f = bfind(h);
                if ((int32_t)f < (int32_t)k2) { check++; }
                sts32(out + 4u * (i + s), h);
                const uint32_t L = K - (f < 30u ? f : 30u);
                w <<= L; cnt -= L;
                if (cnt <= 32u) { w |= (unsigned long long)nw << (32u - cnt); cnt += 32u; wi++; nw = lds32(ring + wi * 4u); }
            }
        }
        hi = (uint32_t)(w >> 32);
    }
    else if (VAR == 7) {       // branch-free groups of 8: two words of lookahead, clamped shifts merge the refill, one branch per group
        uint32_t nw0 = nw, nw1 = lds32(ring + 3u * 4u);
        for (uint32_t i = 0; i < ncw; i += 8) {
            uint32_t bad = 0;
#pragma unroll
            for (int s = 0; s < 8; s++) {
                const uint32_t ldv = lds32(ring + ((wi + 2u) & (WORDS - 1u)) * 4u);
                const uint32_t cK = cnt - K;
                const uint32_t f = bfind(hi);
                bad |= ((int32_t)f < (int32_t)k2) ? 1u : 0u;
                sts32(out + 4u * (i + s), hi);
                const uint32_t m = f < 30u ? f : 30u;
                const uint32_t L = K - m, c2 = cK + m;
                hi = __funnelshift_lc(lo, hi, L) | __funnelshift_rc(nw0, 0u, c2);
                const bool re = c2 < 32u;
                const uint32_t lo1 = __funnelshift_lc(0u, lo, L), lo2 = __funnelshift_r(0u, nw0, c2);
                lo = re ? lo2 : lo1;
                cnt = re ? c2 + 32u : c2;
                wi = re ? wi + 1u : wi;
                nw0 = re ? nw1 : nw0;
                nw1 = re ? ldv : nw1;
            }
            if (bad) check++;
        }
    }
    else if (VAR >= 8 && VAR < 16) {
        constexpr int FL = VAR - 8;       // 32-bit window + two raw words and a bit offset; branch-free groups of 8; flag on the FMA pipe
        uint32_t nw0 = lo, nw1 = nw, o = 0, wa = ring + 4u, bad = 0;     // window = word 0, next bits start at bit 0 of word 1
        for (uint32_t i = 0; i < ncw; i += 8) {
#pragma unroll
            for (int s = 0; s < 8; s++) {
                const uint32_t ldv = (FL & 2) ? (nw1 * 2654435761u + 12345u) : lds32(wa + 8u);
                const uint32_t X = __funnelshift_l(nw1, nw0, o);
                const uint32_t f = bfind(hi);
                if (!(FL & 1)) sts32(out + 4u * (i + s), hi);
                const uint32_t m = f < 30u ? f : 30u;
                const uint32_t L = K - m;
                if (!(FL & 4)) bad = __umulhi(f - k2, 2u) + bad;        // + 1 for every f < k2 (IMAD.HI)
                hi = __funnelshift_lc(X, hi, L);
                o += L;
                if (o >= 32u) { o -= 32u; nw0 = nw1; nw1 = ldv; wa += 4u; }
            }
            if (bad) { check++; bad = 0; }
        }
        wi = (wa - ring) / 4u; cnt = 32u - o + 32u; wi += 1;   // position = wi_word*32 + o - 32
        cnt = 0; wi = 0; { const uint32_t j = (wa - ring) / 4u; wi = j; cnt = 32u - o; }
    }
    else if (VAR == 16) {      // 32-bit window + two raw words and a bit offset; no predicates, no min: masks and a 3-input add
        uint32_t nw0 = lo, nw1 = nw, o = 0, wa = ring + 4u, bad = 0;
        for (uint32_t i = 0; i < ncw; i += 8) {
#pragma unroll
            for (int s = 0; s < 8; s++) {
                const uint32_t ldv = lds32(wa + 8u);
                const uint32_t X = __funnelshift_l(nw1, nw0, o);
                const uint32_t f = bfind(hi);
                const uint32_t t = hi >> 31;
                sts32(out + 4u * (i + s), hi);
                const uint32_t L = K - f + t;                           // k2 + 1 + max(lz, 1)
                bad = __umulhi(f - k2, 2u) + bad;                       // + 1 for every f < k2 (IMAD.HI)
                hi = __funnelshift_lc(X, hi, L);
                const uint32_t o2 = o + L;
                const uint32_t u = o2 >> 5;
                o = o2 & 31u;
                const uint32_t M = 0u - u;
                nw0 = (nw0 & ~M) | (nw1 & M);
                nw1 = (nw1 & ~M) | (ldv & M);
                wa = u * 4u + wa;
            }
            if (bad) { check++; bad = 0; }
        }
        { const uint32_t j = (wa - ring) / 4u; wi = j; cnt = 32u - o; }
    }
    else if (VAR >= 17 && VAR < 25) {
        constexpr int FL = VAR - 17;      // 64-bit window (hi, mid) + X one step ahead; masks via single LOP3 selects
        // hi = word 0, mid = word 1, next bits start at bit 0 of word 2
        uint32_t mid = lo, nw0 = nw, nw1 = lds32(ring + 12u), o = 0, wa = ring + 8u, bad = 0;
        uint32_t X = nw0;
        for (uint32_t i = 0; i < ncw; i += 8) {
#pragma unroll
            for (int s = 0; s < 8; s++) {
                const uint32_t ldv = (FL & 2) ? (nw1 * 2654435761u + 12345u) : lds32(wa + 8u);
                const uint32_t f = bfind(hi);
                const uint32_t t = hi >> 31;
                if (!(FL & 1)) sts32(out + 4u * (i + s), hi);
                const uint32_t L = K - f + t;                           // k2 + 1 + max(lz, 1)
                if (!(FL & 4)) bad = __umulhi(f - k2, 2u) + bad;
                hi = __funnelshift_lc(mid, hi, L);
                mid = __funnelshift_lc(X, mid, L);
                const uint32_t o2 = o + L;
                const uint32_t u = o2 >> 5;
                o = o2 & 31u;
                const uint32_t M = 0u - u;
                uint32_t a0, a1;
                asm("lop3.b32 %0, %1, %2, %3, 0xD8;" : "=r"(a0) : "r"(nw0), "r"(nw1), "r"(M));   // M ? nw1 : nw0
                asm("lop3.b32 %0, %1, %2, %3, 0xD8;" : "=r"(a1) : "r"(nw1), "r"(ldv), "r"(M));   // M ? ldv : nw1
                nw0 = a0; nw1 = a1;
                wa = u * 4u + wa;
                X = __funnelshift_l(nw1, nw0, o);
            }
            if (bad) { check++; bad = 0; }
        }
        { const uint32_t j = (wa - ring) / 4u; wi = j; cnt = 32u - o + 32u; }
    }
    else if (VAR == 25 || VAR == 26) {      // 128-bit window, groups of G code words, window reloaded by position after each group
        constexpr int G = (VAR == 25) ? 8 : 4;
        uint32_t pos = 0, w0 = sm[0], w1 = sm[1], w2 = sm[2], w3 = sm[3];
        for (uint32_t i = 0; i < ncw; i += G) {
            uint32_t bad = 0, T = 0;
#pragma unroll
            for (int s = 0; s < G; s++) {
                const uint32_t f = bfind(w0);
                const uint32_t t = w0 >> 31;
                sts32(out + 4u * (i + s), w0);
                const uint32_t L = K - f + t;
                bad = __umulhi(f - k2, 2u) + bad;
                w0 = __funnelshift_lc(w1, w0, L); w1 = __funnelshift_lc(w2, w1, L); w2 = __funnelshift_lc(w3, w2, L); w3 = __funnelshift_lc(0u, w3, L);
                T += L;
            }
            pos += T;
            if (bad | (T > 96u)) { check++; }
            const uint32_t a = ring + ((pos >> 5) & (WORDS - 1u)) * 4u, sh = pos & 31u;
            const uint32_t v0 = lds32(a), v1 = lds32(a + 4u), v2 = lds32(a + 8u), v3 = lds32(a + 12u), v4 = lds32(a + 16u);
            w0 = __funnelshift_l(v1, v0, sh); w1 = __funnelshift_l(v2, v1, sh); w2 = __funnelshift_l(v3, v2, sh); w3 = __funnelshift_l(v4, v3, sh);
        }
        wi = (pos >> 5) + 2; cnt = 64 - (pos & 31); hi = w0;
    }
    else if (VAR == 27) {      // v25 with the leading-one index from the FP32 adder instead of FLO
        constexpr int G = 8;
        uint32_t pos = 0, w0 = sm[0], w1 = sm[1], w2 = sm[2], w3 = sm[3];
        const uint32_t K95 = K + 95u;
        for (uint32_t i = 0; i < ncw; i += G) {
            uint32_t bad = 0, T = 0;
#pragma unroll
            for (int s = 0; s < G; s++) {
                // 1.m - 1.0 with m = the window's top 23 bits: the result's exponent field is f + 95 (f = index of the leading one)
                const float x = __uint_as_float(__funnelshift_r(w0, 0x7Fu, 9)) - 1.0f;
                const uint32_t E = __float_as_uint(x) >> 23;
                const uint32_t t = w0 >> 31;
                sts32(out + 4u * (i + s), w0);
                const uint32_t L = K95 - E + t;
                bad = __umulhi(E - 95u - k2, 2u) + bad;
                w0 = __funnelshift_lc(w1, w0, L); w1 = __funnelshift_lc(w2, w1, L); w2 = __funnelshift_lc(w3, w2, L); w3 = __funnelshift_lc(0u, w3, L);
                T += L;
            }
            pos += T;
            if (bad | (T > 96u)) { check++; }
            const uint32_t a = ring + ((pos >> 5) & (WORDS - 1u)) * 4u, sh = pos & 31u;
            const uint32_t v0 = lds32(a), v1 = lds32(a + 4u), v2 = lds32(a + 8u), v3 = lds32(a + 12u), v4 = lds32(a + 16u);
            w0 = __funnelshift_l(v1, v0, sh); w1 = __funnelshift_l(v2, v1, sh); w2 = __funnelshift_l(v3, v2, sh); w3 = __funnelshift_l(v4, v3, sh);
        }
        wi = (pos >> 5) + 2; cnt = 64 - (pos & 31); hi = w0;
    }
    long long t1 = clock64();
    res->cycles = t1 - t0;
    res->endpos = wi * 32u - cnt;
    res->check = check + hi;
    sink[0] = check;
}

// ---- warp-wide variant: lane j speculates code word j of a round at pos + j*s + E, E in 0..7 candidates, and the
//      round resolves by composing the per-lane offset tables with shuffles (see DESIGN notes) -- omitted here ----

static void gen(std::vector<uint32_t> &words, uint32_t k2, double p_long, uint32_t &bits_out, uint32_t ncw)
{
    std::vector<uint8_t> bits;
    bits.reserve(WORDS * 32);
    auto put = [&](uint32_t v, int n) { for (int i = n - 1; i >= 0; i--) bits.push_back((v >> i) & 1); };
    srand(12345);
    for (uint32_t i = 0; i < ncw; i++) {
        double u = rand() / (double)RAND_MAX;
        uint32_t lz = 0;
        if (u < p_long) { lz = 2; while (rand() % 2 && lz < 6) lz++; }
        else lz = rand() % 2;
        if (lz == 0) { put(1, 1); put(rand(), k2 + 1); }
        else { put(0, lz); put(1, 1); put(rand(), k2); }
    }
    bits_out = (uint32_t)bits.size();
    while (bits.size() < (size_t)WORDS * 32) bits.push_back(0);
    words.assign(WORDS, 0);
    for (size_t i = 0; i < (size_t)WORDS * 32; i++) if (bits[i]) words[i >> 5] |= 1u << (31 - (i & 31));
}

template <int VAR> static void run(const char *name, const uint32_t *d, uint32_t k2, uint32_t ncw, uint32_t bits, Result *dres, uint32_t *dsink)
{
    size_t smem = (WORDS + NCW + 64) * 4;
    cudaFuncSetAttribute(walk<VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int rep = 0; rep < 2; rep++) walk<VAR><<<1, 32, smem>>>(d, k2, ncw, dres, dsink);
    Result r;
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(&r, dres, sizeof(r), cudaMemcpyDeviceToHost);
    printf("%-28s k2=%u: %7.1f cycles/cw  endpos=%u (want %u) %s %s\n", name, k2, (double)r.cycles / ncw, r.endpos, bits,
           r.endpos == bits ? "OK" : "MISMATCH", e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main()
{
    uint32_t *d; Result *dres; uint32_t *dsink;
    cudaMalloc(&d, WORDS * 4); cudaMalloc(&dres, sizeof(Result)); cudaMalloc(&dsink, 64);
    for (uint32_t k2 : {6u, 8u}) {
        for (double pl : {0.2}) {
            std::vector<uint32_t> words; uint32_t bits;
            gen(words, k2, pl, bits, NCW);
            cudaMemcpy(d, words.data(), WORDS * 4, cudaMemcpyHostToDevice);
            printf("--- k2=%u p_long=%.2f bits/cw=%.2f\n", k2, pl, (double)bits / NCW);
            run<0>("v0 hi/lo/nw FLO branchy", d, k2, NCW, bits, dres, dsink);
            run<1>("v1 branch-free refill", d, k2, NCW, bits, dres, dsink);
            run<2>("v2 position + 2 LDS", d, k2, NCW, bits, dres, dsink);
            run<3>("v3 compares, no FLO", d, k2, NCW, bits, dres, dsink);
            run<4>("v4 pairs of short cws", d, k2, NCW, bits, dres, dsink);
            run<5>("v5 chain only", d, k2, NCW, bits, dres, dsink);
            run<6>("v6 64-bit window", d, k2, NCW, bits, dres, dsink);
            run<7>("v7 branch-free groups of 8", d, k2, NCW, bits, dres, dsink);
            run<8>("v8 window + 2 words + offset", d, k2, NCW, bits, dres, dsink);
            run<16>("v16 window+2 words, no predicates", d, k2, NCW, bits, dres, dsink);
            run<17>("v17 (hi,mid)+X ahead, LOP3 selects", d, k2, NCW, bits, dres, dsink);
            run<25>("v25 128-bit window, groups of 8", d, k2, NCW, bits, dres, dsink);
            run<26>("v26 128-bit window, groups of 4", d, k2, NCW, bits, dres, dsink);
            run<27>("v27 groups of 8, exponent of 1.m-1", d, k2, NCW, bits, dres, dsink);
            run<18>("v17 no STS", d, k2, NCW, bits, dres, dsink);
            run<19>("v17 no LDS", d, k2, NCW, bits, dres, dsink);
            run<21>("v17 no flag", d, k2, NCW, bits, dres, dsink);
            run<24>("v17 none", d, k2, NCW, bits, dres, dsink);
            run<9>("v8 no STS", d, k2, NCW, bits, dres, dsink);
            run<10>("v8 no LDS", d, k2, NCW, bits, dres, dsink);
            run<12>("v8 no flag", d, k2, NCW, bits, dres, dsink);
            run<15>("v8 none of them", d, k2, NCW, bits, dres, dsink);
        }
    }
    return 0;
}
