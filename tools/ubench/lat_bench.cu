// lat_bench.cu -- dependent-issue latencies of the integer instructions the code-word walk is made of, one warp,
// one lane's worth of work (all lanes do the same).  Prints cycles per dependent operation.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t bfind(uint32_t x) { uint32_t r; asm volatile("bfind.u32 %0, %1;" : "=r"(r) : "r"(x)); return r; }

template <int OP>
__global__ void lat(uint32_t a, uint32_t b, uint32_t n, long long *out, uint32_t *sink)
{
    __shared__ uint32_t sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = ((i * 7 + 3) & 1023) * 4;   // pointer chase table (byte offsets)
    __syncthreads();
    uint32_t x = (OP == 4) ? 0u : a + threadIdx.x * 0, y = b, z = 0;
    uint32_t r0 = a, r1 = a + 1, r2 = a + 2, r3 = a + 3, r4 = a + 4, r5 = a + 5, r6 = a + 6, r7 = a + 7;
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(sm);
    long long t0 = clock64();
    for (uint32_t i = 0; i < n; i++) {
#pragma unroll
        for (int u = 0; u < 16; u++) {
            if (OP == 0) { asm volatile("add.u32 %0, %0, %1;" : "+r"(x) : "r"(y)); }
            else if (OP == 1) { asm volatile("shf.l.clamp.b32 %0, %1, %0, %2;" : "+r"(x) : "r"(y), "r"(b)); }
            else if (OP == 2) { x = bfind(x) | 0x10000u; }                                        // FLO + LOP
            else if (OP == 3) { asm volatile("min.u32 %0, %0, %1;" : "+r"(x) : "r"(y)); asm volatile("add.u32 %0, %0, 5;" : "+r"(x)); }  // MNMX + ADD
            else if (OP == 4) { asm volatile("ld.shared.u32 %0, [%1];" : "=r"(x) : "r"(base + x)); }
            else if (OP == 5) { asm volatile("xor.b32 %0, %0, %1;" : "+r"(x) : "r"(y)); }
            else if (OP == 6) { // never-taken branch in a dependent add chain
                asm volatile("add.u32 %0, %0, %1;" : "+r"(x) : "r"(y));
                if (x == 0xdeadbeefu) { z += x * 3; x ^= z; }
            }
            else if (OP == 7) { // taken forward branch (skips 4 instructions) in a dependent add chain
                asm volatile("add.u32 %0, %0, %1;" : "+r"(x) : "r"(y));
                if (x == 0xdeadbeefu || (y & 0x100u)) { asm volatile("add.u32 %0, %0, 7; xor.b32 %0, %0, 9; add.u32 %0, %0, 3; xor.b32 %0,%0, 5;" : "+r"(z)); }
            }
            else if (OP == 8) { // popc chain
                asm volatile("popc.b32 %0, %0;" : "+r"(x)); asm volatile("add.u32 %0, %0, %1;" : "+r"(x) : "r"(y));
            }
            else if (OP == 9) { // imad chain
                asm volatile("mad.lo.u32 %0, %0, %1, %1;" : "+r"(x) : "r"(y));
            }
            else if (OP == 10) { // two independent add chains (ILP 2)
                asm volatile("add.u32 %0, %0, %1;" : "+r"(x) : "r"(y)); asm volatile("add.u32 %0, %0, %1;" : "+r"(z) : "r"(y));
            }
            else if (OP == 11) { // four independent add chains
                uint32_t w = z;
                asm volatile("add.u32 %0, %0, %1;" : "+r"(x) : "r"(y)); asm volatile("add.u32 %0, %0, %1;" : "+r"(z) : "r"(y));
                asm volatile("xor.b32 %0, %0, %1;" : "+r"(y) : "r"(b)); asm volatile("add.u32 %0, %0, %1;" : "+r"(w) : "r"(b));
                z ^= w;
            }
            else if (OP == 15) { // data-dependent branch taken ~1/3 of the time (x mod 3), body of 6 instructions
                asm volatile("add.u32 %0, %0, %1;" : "+r"(x) : "r"(y));
                if ((x % 3u) == 0u) { asm volatile("add.u32 %0, %0, 7; xor.b32 %0, %0, 9; add.u32 %0, %0, 3; xor.b32 %0,%0, 5; add.u32 %0, %0, 11; xor.b32 %0, %0, 1;" : "+r"(z)); }
            }
            else if (OP == 16) { // the same as predicated selects (no branch)
                asm volatile("add.u32 %0, %0, %1;" : "+r"(x) : "r"(y));
                uint32_t w = z;
                asm volatile("add.u32 %0, %0, 7; xor.b32 %0, %0, 9; add.u32 %0, %0, 3; xor.b32 %0,%0, 5; add.u32 %0, %0, 11; xor.b32 %0, %0, 1;" : "+r"(w));
                z = ((x % 3u) == 0u) ? w : z;
            }
            else if (OP == 17) { // the walk's chain: FLO -> MNMX -> SUB -> SHF
                uint32_t f = bfind(x); uint32_t m = f < 30u ? f : 30u; uint32_t L = 40u - m;
                asm volatile("shf.l.clamp.b32 %0, %1, %0, %2;" : "+r"(x) : "r"(y), "r"(L));
            }
            else if (OP == 18) { // FLO -> SHF
                uint32_t f = bfind(x);
                asm volatile("shf.l.clamp.b32 %0, %1, %0, %2;" : "+r"(x) : "r"(y), "r"(f));
            }
            else if (OP == 19) { // MNMX -> SHF
                uint32_t m = x < 30u ? x : 30u;
                asm volatile("shf.l.clamp.b32 %0, %1, %0, %2;" : "+r"(x) : "r"(y), "r"(m));
            }
            else if (OP == 20) { // SUB(imad) -> SHF
                uint32_t L = 40u - x;
                asm volatile("shf.l.clamp.b32 %0, %1, %0, %2;" : "+r"(x) : "r"(y), "r"(L));
            }
            else if (OP == 21) { // FLO -> FLO
                x = bfind(x) + 0u; x = bfind(x | 0x100u);
            }
            else if (OP == 22) { // 8 independent ALU chains (shifts by a runtime amount: no folding)
                asm volatile("shf.l.clamp.b32 %0, %8, %0, %9; shf.l.clamp.b32 %1, %8, %1, %9; shf.l.clamp.b32 %2, %8, %2, %9; shf.l.clamp.b32 %3, %8, %3, %9;"
                             "shf.l.clamp.b32 %4, %8, %4, %9; shf.l.clamp.b32 %5, %8, %5, %9; shf.l.clamp.b32 %6, %8, %6, %9; shf.l.clamp.b32 %7, %8, %7, %9;"
                             : "+r"(r0), "+r"(r1), "+r"(r2), "+r"(r3), "+r"(r4), "+r"(r5), "+r"(r6), "+r"(r7) : "r"(y), "r"(b));
            }
            else if (OP == 23) { // 4 independent ALU chains + 4 independent IMAD chains
                asm volatile("shf.l.clamp.b32 %0, %8, %0, %9; mad.lo.u32 %4, %4, %8, %9; shf.l.clamp.b32 %1, %8, %1, %9; mad.lo.u32 %5, %5, %8, %9;"
                             "shf.l.clamp.b32 %2, %8, %2, %9; mad.lo.u32 %6, %6, %8, %9; shf.l.clamp.b32 %3, %8, %3, %9; mad.lo.u32 %7, %7, %8, %9;"
                             : "+r"(r0), "+r"(r1), "+r"(r2), "+r"(r3), "+r"(r4), "+r"(r5), "+r"(r6), "+r"(r7) : "r"(y), "r"(b));
            }
            else if (OP == 24) { // 8 independent IMAD chains
                asm volatile("mad.lo.u32 %0, %0, %8, %9; mad.lo.u32 %1, %1, %8, %9; mad.lo.u32 %2, %2, %8, %9; mad.lo.u32 %3, %3, %8, %9;"
                             "mad.lo.u32 %4, %4, %8, %9; mad.lo.u32 %5, %5, %8, %9; mad.lo.u32 %6, %6, %8, %9; mad.lo.u32 %7, %7, %8, %9;"
                             : "+r"(r0), "+r"(r1), "+r"(r2), "+r"(r3), "+r"(r4), "+r"(r5), "+r"(r6), "+r"(r7) : "r"(y), "r"(b));
            }
            else if (OP == 25) { // FLO -> IMAD (FMA pipe) -> SHF
                uint32_t f = bfind(x); uint32_t L;
                asm volatile("mad.lo.u32 %0, %1, 3, %2;" : "=r"(L) : "r"(f), "r"(y));
                asm volatile("shf.l.clamp.b32 %0, %1, %0, %2;" : "+r"(x) : "r"(y), "r"(L));
            }
            else if (OP == 26) { // FLO -> LOP3 (ALU pipe) -> SHF
                uint32_t f = bfind(x); uint32_t L;
                asm volatile("xor.b32 %0, %1, %2;" : "=r"(L) : "r"(f), "r"(y));
                asm volatile("shf.l.clamp.b32 %0, %1, %0, %2;" : "+r"(x) : "r"(y), "r"(L));
            }
            else if (OP == 27) { // SHF -> IMAD -> SHF (ALU -> FMA -> ALU)
                uint32_t L;
                asm volatile("mad.lo.u32 %0, %1, 3, %2;" : "=r"(L) : "r"(x), "r"(y));
                asm volatile("shf.l.clamp.b32 %0, %1, %0, %2;" : "+r"(x) : "r"(y), "r"(L));
            }
            else if (OP == 28) { // the walk's group step: FLO, LEA.HI, sub, 4 SHF
                const uint32_t f = bfind(r0);
                const uint32_t t = r0 >> 31;
                const uint32_t L = b + 32u - f + t;
                r0 = __funnelshift_lc(r1, r0, L); r1 = __funnelshift_lc(r2, r1, L); r2 = __funnelshift_lc(r3, r2, L); r3 = __funnelshift_lc(y, r3, L);
            }
            else if (OP == 29) { // the same with the length forced through the ALU pipe (lop3-based subtract: K - f == K + (f ^ 31) - 31 ... as add.u32 in asm)
                const uint32_t f = bfind(r0);
                uint32_t L;
                asm volatile("{ .reg .u32 t; shr.u32 t, %1, 31; sub.u32 t, t, %2; add.u32 %0, t, %3; }" : "=r"(L) : "r"(r0), "r"(f), "r"(b + 32u));
                r0 = __funnelshift_lc(r1, r0, L); r1 = __funnelshift_lc(r2, r1, L); r2 = __funnelshift_lc(r3, r2, L); r3 = __funnelshift_lc(y, r3, L);
            }
            else if (OP == 12) { // shfl chain
                x = __shfl_sync(0xffffffffu, x, (x + 1) & 31);
            }
            else if (OP == 13) { // prmt chain
                asm volatile("prmt.b32 %0, %0, %1, 0x0123;" : "+r"(x) : "r"(y));
            }
            else if (OP == 14) { // select on compare: setp + selp chain
                asm volatile("{ .reg .pred p; setp.lt.u32 p, %0, %1; selp.u32 %0, %2, %0, p; add.u32 %0, %0, 1; }" : "+r"(x) : "r"(y), "r"(b));
            }
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = t1 - t0; sink[0] = x + z + y + r0 + r1 + r2 + r3 + r4 + r5 + r6 + r7; }
}

template <int OP> static void run(const char *name, int per, int warps, long long *d, uint32_t *s)
{
    const uint32_t n = 2000;
    lat<OP><<<1, 32 * warps>>>(12345u, 7u, n, d, s);
    lat<OP><<<1, 32 * warps>>>(12345u, 7u, n, d, s);
    long long c; cudaDeviceSynchronize(); cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
    printf("%-44s warps=%d: %6.2f cycles per step (%d dependent ops per step)\n", name, warps, (double)c / (n * 16.0), per);
}

int main()
{
    long long *d; uint32_t *s; cudaMalloc(&d, 8); cudaMalloc(&s, 64);
    for (int warps : {1}) {
        run<0>("IADD chain", 1, warps, d, s);
        run<5>("LOP chain", 1, warps, d, s);
        run<1>("SHF (funnel, variable) chain", 1, warps, d, s);
        run<2>("FLO + LOP chain", 2, warps, d, s);
        run<3>("MNMX + ADD chain", 2, warps, d, s);
        run<8>("POPC + ADD chain", 2, warps, d, s);
        run<9>("IMAD chain", 1, warps, d, s);
        run<13>("PRMT chain", 1, warps, d, s);
        run<14>("SETP + SELP + ADD chain", 3, warps, d, s);
        run<4>("LDS pointer chase", 1, warps, d, s);
        run<12>("SHFL chain (+2 alu)", 3, warps, d, s);
        run<6>("ADD + never-taken branch", 1, warps, d, s);
        run<7>("ADD + taken forward branch", 1, warps, d, s);
        run<15>("ADD + 1/3-taken branch (6-instr body)", 1, warps, d, s);
        run<16>("ADD + the same body as a select", 1, warps, d, s);
        run<17>("FLO -> MNMX -> SUB -> SHF chain", 4, warps, d, s);
        run<18>("FLO -> SHF chain", 2, warps, d, s);
        run<19>("MNMX -> SHF chain", 2, warps, d, s);
        run<20>("SUB -> SHF chain", 2, warps, d, s);
        run<21>("FLO -> LOP -> FLO chain", 3, warps, d, s);
        run<25>("FLO -> IMAD -> SHF chain", 3, warps, d, s);
        run<26>("FLO -> LOP -> SHF chain", 3, warps, d, s);
        run<27>("IMAD -> SHF chain", 2, warps, d, s);
        run<28>("walk step (FLO, len, 4 SHF)", 3, warps, d, s);
        run<29>("walk step, asm length", 3, warps, d, s);
        run<22>("8 independent SHF (ALU pipe)", 8, warps, d, s);
        run<23>("4 SHF + 4 IMAD independent", 8, warps, d, s);
        run<24>("8 independent IMAD (FMA pipe)", 8, warps, d, s);
        run<10>("2 independent ADD chains", 1, warps, d, s);
        run<11>("4 independent chains", 1, warps, d, s);
    }
    return 0;
}
