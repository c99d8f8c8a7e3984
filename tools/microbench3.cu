// microbench3.cu -- per-instruction throughput of the integer ops behind 64-bit modular arithmetic (inline PTX so the
// SASS opcode is known): IMAD (lo), IMAD.HI, IMAD.WIDE, IADD3, LOP3/SHF, and mixes, at 16 warps/SMSP with 8 chains.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
typedef unsigned long long u64;
typedef unsigned int u32;
#define ITERS 2048
#define ILP 8

template <int MODE>
__global__ void probe(u32 *out, u32 seed, long long *clk) {
    u32 a[ILP], b[ILP], c[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { a[i] = seed + threadIdx.x * 7919u + i; b[i] = seed * 31 + threadIdx.x + i * 3; c[i] = seed ^ (i * 77u); }
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (MODE == 0) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b[i]), "r"(c[i]));
            if (MODE == 1) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b[i]), "r"(c[i]));
            if (MODE == 2) { u64 t = ((u64)c[i] << 32) | a[i]; asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(t) : "r"(a[i]), "r"(b[i])); a[i] = (u32)t; c[i] = (u32)(t >> 32); }
            if (MODE == 3) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b[i]));
            if (MODE == 4) { asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3;" : "+r"(a[i]), "+r"(c[i]) : "r"(b[i]), "r"(seed)); }
            if (MODE == 5) { asm volatile("mad.lo.u32 %0, %0, %2, %1; add.u32 %1, %1, %0;" : "+r"(a[i]), "+r"(c[i]) : "r"(b[i])); }  // imad + iadd pair
            if (MODE == 6) { u64 t = ((u64)c[i] << 32) | a[i]; asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(t) : "r"(a[i]), "r"(b[i])); a[i] = (u32)t; c[i] = (u32)(t >> 32) + b[i]; }  // wide + iadd
            if (MODE == 7) { asm volatile("mad.lo.u32 %0, %0, %2, %1; mad.lo.u32 %1, %1, %2, %0;" : "+r"(a[i]), "+r"(c[i]) : "r"(b[i])); }  // 2 imad
            if (MODE == 8) { u32 lo, hi; asm volatile("mul.lo.u32 %0, %2, %3; mul.hi.u32 %1, %2, %3;" : "=r"(lo), "=r"(hi) : "r"(a[i]), "r"(b[i])); a[i] = lo + c[i]; c[i] = hi; }  // lo+hi pair instead of wide
            if (MODE == 9) asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(a[i]) : "r"(b[i]));
            if (MODE == 10) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b[i]), "r"(c[i]));
        }
    }
    long long t1 = clock64();
    u32 acc = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc ^= a[i] ^ b[i] ^ c[i];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char *name, int ops_per_iter, int sms, int warps_per_smsp, u32 *out, long long *clk, long long *hclk) {
    const int threads = 128, bps = warps_per_smsp, blocks = sms * bps;
    probe<MODE><<<blocks, threads>>>(out, 12345, clk);
    probe<MODE><<<blocks, threads>>>(out, 12345, clk);
    cudaDeviceSynchronize();
    cudaMemcpy(hclk, clk, sizeof(long long) * blocks, cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < blocks; ++i) avg += (double)hclk[i];
    avg /= blocks;
    const double warp_instr_per_clk_smsp = (double)bps * (threads / 32) * ITERS * ILP * ops_per_iter / avg / 4.0;
    printf("%-28s warps/SMSP=%2d : %5.3f warp-instr/clk/SMSP  (%5.1f lanes/clk/SM)\n", name, warps_per_smsp,
           warp_instr_per_clk_smsp, warp_instr_per_clk_smsp * 128);
}

int main() {
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    u32 *out; long long *clk, *hclk;
    cudaMalloc(&out, sizeof(u32) * sms * 16 * 128);
    cudaMalloc(&clk, sizeof(long long) * sms * 16);
    hclk = (long long *)malloc(sizeof(long long) * sms * 16);
    for (int w : {8, 16}) {
        run<0>("IMAD (mad.lo.u32)", 1, sms, w, out, clk, hclk);
        run<1>("IMAD.HI (mad.hi.u32)", 1, sms, w, out, clk, hclk);
        run<2>("IMAD.WIDE (mad.wide.u32)", 1, sms, w, out, clk, hclk);
        run<3>("IADD (add.u32)", 1, sms, w, out, clk, hclk);
        run<4>("IADD3+IADD3.X (64-bit add)", 2, sms, w, out, clk, hclk);
        run<5>("IMAD + IADD mix", 2, sms, w, out, clk, hclk);
        run<6>("IMAD.WIDE + IADD mix", 2, sms, w, out, clk, hclk);
        run<7>("IMAD x2 dependent", 2, sms, w, out, clk, hclk);
        run<8>("mul.lo + mul.hi pair", 2, sms, w, out, clk, hclk);
        run<9>("SHF", 1, sms, w, out, clk, hclk);
        run<10>("LOP3", 1, sms, w, out, clk, hclk);
    }
    return 0;
}
