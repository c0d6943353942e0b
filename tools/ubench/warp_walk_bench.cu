// warp_walk_bench.cu -- micro-benchmark of the WARP-WIDE code-word walk (recursive-Rice lengths): lane j guesses that
// code word j of a round of 32 starts at pos + j * (k2 + 2); every lane keeps a 64-bit window at its guess, so a long
// code word found at lane j1 (extra = max(lz, 1) - 1 bits) only shifts the lanes behind it inside their windows --
// one redux per long code word, no memory access -- until nothing is left to fix.  Measures cycles per code word of
// one warp alone (latency) and of a full machine of warps (throughput), against the serial walk of walk_bench.cu.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o warp_walk_bench warp_walk_bench.cu ; run on a B200.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

#define RING   256u          // words of the per-warp payload ring
#define CHUNK  128u          // words per refill (four per lane, one 16-byte load)
#define MIRROR 4u

__device__ __forceinline__ uint32_t lds32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ int32_t zz_dec(uint32_t u) { return (int32_t)(u >> 1) ^ -(int32_t)(u & 1u); }
__device__ __forceinline__ uint32_t bswap(uint32_t x) { return __byte_perm(x, 0, 0x0123); }

struct Ring {
    uint32_t *buf;           // [RING + MIRROR] shared
    uint32_t saddr;
    const uint4 *g;          // 16-byte aligned start of the payload
    uint32_t nchunk16;       // readable 16-byte lines
    uint32_t filled;         // words [0, filled) have been stored into the ring
    uint4 pre;               // this lane's part of chunk [filled, filled + CHUNK)
};
__device__ __forceinline__ uint4 ring_fetch(const Ring &r, uint32_t first_word, uint32_t lane)
{
    const uint32_t i16 = first_word / 4u + lane;
    return (i16 < r.nchunk16) ? __ldg(r.g + i16) : make_uint4(0, 0, 0, 0);
}
__device__ __forceinline__ void ring_advance(Ring &r, uint32_t lane)
{
    const uint32_t slot = r.filled % RING;
    const uint4 v = make_uint4(bswap(r.pre.x), bswap(r.pre.y), bswap(r.pre.z), bswap(r.pre.w));
    *(uint4 *)(r.buf + slot + lane * 4u) = v;
    if (slot == 0u && lane == 0u) *(uint4 *)(r.buf + RING) = v;
    r.filled += CHUNK;
    r.pre = ring_fetch(r, r.filled, lane);
    __syncwarp();
}
// make words up to (and including) `need_word` readable
__device__ __forceinline__ void ring_ensure(Ring &r, uint32_t need_word, uint32_t lane)
{
    while (need_word >= r.filled) ring_advance(r, lane);
}

// One round: up to R (<= 32) code words of Rice parameter k2 from bit position pos.  Returns the number of code words
// that are final (n_ok <= R; less when a code word does not fit its 32-bit window or the extras outgrow the lanes'
// windows) and the bits they take; lane j < n_ok holds its code word's first 32 bits in v.
template <int VAR>
__device__ __forceinline__ uint32_t round32(const Ring &ring, uint32_t pos, uint32_t R, uint32_t k2, uint32_t lane, uint32_t &v, uint32_t &bits, uint32_t &iters)
{
    const uint32_t L = k2 + 2u;
    const uint32_t p = pos + lane * L;
    const uint32_t a = ring.saddr + ((p >> 5) % RING) * 4u;
    const uint32_t w0 = lds32(a), w1 = lds32(a + 4u), w2 = lds32(a + 8u);
    const uint32_t hi = __funnelshift_l(w1, w0, p), lo = __funnelshift_l(w2, w1, p);
    const uint32_t xmax = 30u - k2;                       // a long code word of extra x fits 32 bits iff x + 1 + k2 <= 31
    uint32_t e = 0, first = 0, E = 0, n_ok = R;
    const uint32_t key_hi = lane << 8;
    for (;;) {
        v = __funnelshift_lc(lo, hi, e);
        const uint32_t lz = (uint32_t)__clz((int)v);
        const uint32_t x = (lz > 1u ? lz : 1u) - 1u;
        const bool cand = lane >= first && lane < R && x != 0u;
        iters++;
        uint32_t j1, xv;
        if (VAR == 0) {
            const uint32_t r = __reduce_min_sync(0xffffffffu, cand ? (key_hi | x) : 0xFFFFFFFFu);
            if (r == 0xFFFFFFFFu) break;
            j1 = r >> 8; xv = r & 255u;
        } else {
            const uint32_t m = __ballot_sync(0xffffffffu, cand);
            if (m == 0u) break;
            j1 = (uint32_t)__ffs((int)m) - 1u;
            xv = __shfl_sync(0xffffffffu, x, (int)j1);
        }
        if (xv > xmax) { n_ok = j1; break; }               // does not fit: the caller finishes it serially
        E += xv;
        if (E > 32u) { n_ok = j1 + 1u; break; }            // the lanes behind cannot follow inside their windows
        if (lane > j1) e += xv;
        first = j1 + 1u;
    }
    bits = n_ok * L + E;
    return n_ok;
}

__device__ __forceinline__ int32_t value_of(uint32_t hi, uint32_t k2)
{
    const uint32_t lz = (uint32_t)__clz((int)hi);
    const uint32_t ml = (lz > 1u) ? lz : 1u;
    const uint32_t low = (hi >> ((31u - k2 - ml) & 31u)) & ((1u << k2) - 1u);
    const uint32_t mult = lz ? lz + 1u : ((hi >> 30) & 1u);
    return zz_dec((mult << k2) + low);
}

struct Result { long long cycles; uint32_t endpos; uint32_t rounds; uint32_t iters; uint32_t serial; };

template <int VAR>
__global__ void __launch_bounds__(128) wwalk(const uint32_t *g, uint32_t nwords, uint32_t k2, uint32_t ncw, int32_t *out, Result *res)
{
    __shared__ __align__(16) uint32_t s_ring[4][RING + MIRROR];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const uint32_t gw = blockIdx.x * (blockDim.x >> 5) + warp;
    Ring ring;
    ring.buf = s_ring[warp];
    ring.saddr = (uint32_t)__cvta_generic_to_shared(ring.buf);
    ring.g = (const uint4 *)g; ring.nchunk16 = nwords / 4u; ring.filled = 0;
    ring.pre = ring_fetch(ring, 0u, lane);
    int32_t *o = out + (size_t)gw * ncw;
    uint32_t pos = 0, done = 0, rounds = 0, iters = 0, serial = 0;
    const long long t0 = clock64();
    while (done < ncw) {
        const uint32_t R = (ncw - done < 32u) ? ncw - done : 32u;
        ring_ensure(ring, ((pos + 32u * (k2 + 2u) + 128u) >> 5) + 1u, lane);
        uint32_t v, bits;
        const uint32_t n_ok = round32<VAR>(ring, pos, R, k2, lane, v, bits, iters);
        if (lane < n_ok) o[done + lane] = value_of(v, k2);
        pos += bits; done += n_ok; rounds++;
        if (n_ok < R && n_ok < 32u) {
            // maybe a code word that does not fit 32 bits: one lane reads it bit by bit (rare)
            const uint32_t a = ring.saddr;
            uint32_t q = 0, pp = pos;
            for (;;) {
                ring_ensure(ring, (pp >> 5) + 3u, lane);
                const uint32_t ad = a + ((pp >> 5) % RING) * 4u;
                const uint32_t h = __funnelshift_l(lds32(ad + 4u), lds32(ad), pp);
                const uint32_t z = (uint32_t)__clz((int)h);
                q += z; pp += z;
                if (z < 32u) break;
            }
            pp += 1u;
            if (q + k2 > 30u) {                              // really long: finish it here
                ring_ensure(ring, (pp >> 5) + 3u, lane);
                const uint32_t ad = a + ((pp >> 5) % RING) * 4u;
                const uint32_t kk = q ? k2 : k2 + 1u;
                const uint32_t h = __funnelshift_l(lds32(ad + 4u), lds32(ad), pp);
                const uint32_t low = kk ? h >> (32u - kk) : 0u;
                pp += kk;
                const uint32_t u = q ? low + (2u << k2) + ((q - 1u) << k2) : low;
                if (lane == 0) o[done] = zz_dec(u);
                done++; pos = pp; serial++;
            }
        }
    }
    const long long t1 = clock64();
    if (gw == 0 && lane == 0) { res->cycles = t1 - t0; res->endpos = pos; res->rounds = rounds; res->iters = iters; res->serial = serial; }
}

static void gen(std::vector<uint32_t> &words, std::vector<int32_t> &vals, uint32_t k2, double p_long, uint32_t &bits_out, uint32_t ncw)
{
    std::vector<uint8_t> bits;
    auto put = [&](uint32_t v, int n) { for (int i = n - 1; i >= 0; i--) bits.push_back((v >> i) & 1); };
    srand(12345);
    vals.resize(ncw);
    for (uint32_t i = 0; i < ncw; i++) {
        double u = rand() / (double)RAND_MAX;
        uint32_t q = 0;
        if (u < p_long) { q = 2; while (rand() % 2 && q < 40) q++; if (rand() % 2000 == 0) q += 30; }
        else q = rand() % 2;
        uint32_t uval;
        if (q == 0) { uint32_t low = rand() & ((2u << k2) - 1u); put(1, 1); put(low, k2 + 1); uval = low; }
        else { uint32_t low = rand() & ((1u << k2) - 1u); put(0, q); put(1, 1); put(low, k2); uval = low + (2u << k2) + ((q - 1u) << k2); }
        vals[i] = (int32_t)(uval >> 1) ^ -(int32_t)(uval & 1u);
    }
    bits_out = (uint32_t)bits.size();
    size_t nw = ((bits.size() + 31) / 32 + 64 + 3) / 4 * 4;
    words.assign(nw, 0);
    for (size_t i = 0; i < bits.size(); i++) if (bits[i]) words[i >> 5] |= 1u << (31 - (i & 31));
    for (auto &w : words) w = __builtin_bswap32(w);       // the kernel swaps back: payload bytes in stream order
}

template <int VAR> static void run(const char *name, uint32_t k2, double pl, uint32_t ncw, int blocks)
{
    std::vector<uint32_t> words; std::vector<int32_t> vals; uint32_t bits;
    gen(words, vals, k2, pl, bits, ncw);
    uint32_t *d; int32_t *dout; Result *dres;
    const size_t nwarps = (size_t)blocks * 4;
    cudaMalloc(&d, words.size() * 4); cudaMalloc(&dout, nwarps * ncw * 4); cudaMalloc(&dres, sizeof(Result));
    cudaMemcpy(d, words.data(), words.size() * 4, cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms = 0;
    for (int rep = 0; rep < 3; rep++) {
        cudaMemset(dout, 0xFF, nwarps * ncw * 4);
        cudaEventRecord(e0);
        wwalk<VAR><<<blocks, blocks == 1 ? 32 : 128>>>(d, (uint32_t)words.size(), k2, ncw, dout, dres);
        cudaEventRecord(e1);
        cudaError_t err = cudaDeviceSynchronize();
        if (err != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(err)); return; }
        cudaEventElapsedTime(&ms, e0, e1);
    }
    Result r; cudaMemcpy(&r, dres, sizeof(r), cudaMemcpyDeviceToHost);
    const size_t used_warps = blocks == 1 ? 1 : nwarps;
    std::vector<int32_t> got(ncw);
    size_t bad = 0;
    for (size_t w : {(size_t)0, used_warps - 1}) {
        cudaMemcpy(got.data(), dout + w * ncw, ncw * 4, cudaMemcpyDeviceToHost);
        for (uint32_t i = 0; i < ncw; i++) if (got[i] != vals[i]) { if (bad < 3) printf("  warp %zu cw %u: got %d want %d\n", w, i, got[i], vals[i]); bad++; }
    }
    printf("%-22s k2=%2u p_long=%.2f blocks=%5d: %6.1f cycles/cw (warp 0), %.3f ms = %.3f ns/cw/machine, rounds %u iters %u serial %u, endpos %u (want %u) %s\n",
           name, k2, pl, blocks, (double)r.cycles / ncw, ms, ms * 1e6 / ((double)used_warps * ncw), r.rounds, r.iters, r.serial, r.endpos, bits,
           (r.endpos == bits && bad == 0) ? "OK" : "MISMATCH");
    cudaFree(d); cudaFree(dout); cudaFree(dres);
}

int main()
{
    const uint32_t ncw = 20480;
    for (uint32_t k2 : {4u, 8u, 14u}) {
        for (double pl : {0.13, 0.25}) {
            run<0>("redux", k2, pl, ncw, 1);
            run<1>("ballot+shfl", k2, pl, ncw, 1);
        }
    }
    for (int blocks : {148, 148 * 4, 148 * 12, 148 * 24}) {
        run<0>("redux", 8, 0.25, ncw, blocks);
        run<1>("ballot+shfl", 8, 0.25, ncw, blocks);
    }
    return 0;
}
