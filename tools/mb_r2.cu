// mb_r2.cu -- round-2 microbenchmarks: (1) issue rate of the integer instructions behind a 64-bit Shoup butterfly,
// alone and mixed, with residency verified; (2) complete butterfly variants with per-thread twiddles.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mb_r2 mb_r2.cu ; run: ./mb_r2
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
typedef unsigned long long u64;
typedef unsigned int u32;
#define ITERS 512

__device__ __forceinline__ u64 mk(u32 lo, u32 hi) { return ((u64)hi << 32) | lo; }
__device__ __forceinline__ u64 madwide(u32 a, u32 b, u64 c) { u64 d; asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(d) : "r"(a), "r"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 mulwide(u32 a, u32 b) { u64 d; asm("mul.wide.u32 %0, %1, %2;" : "=l"(d) : "r"(a), "r"(b)); return d; }

// ---------------------------------------------------------------- (1) instruction streams
template <int MODE, int ILP>
__global__ void __launch_bounds__(128) istream(u32 *out, u32 seed, long long *clk) {
    u32 a[ILP], b[ILP], c[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { a[i] = seed + threadIdx.x * 7919u + i; b[i] = seed * 31 + threadIdx.x + i * 3; c[i] = seed ^ (i * 77u); }
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (MODE == 0) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b[i]), "r"(c[i]));
            if (MODE == 1) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b[i]), "r"(c[i]));
            if (MODE == 2) { u64 t = mk(a[i], c[i]); asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(t) : "r"(a[i]), "r"(b[i])); a[i] = (u32)t; c[i] = (u32)(t >> 32); }
            if (MODE == 3) { asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3;" : "+r"(a[i]), "+r"(c[i]) : "r"(b[i]), "r"(seed)); }
            if (MODE == 4) {  // 1 wide + 1 imad (independent pipes?)
                u64 t = mk(a[i], c[i]); asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(t) : "r"(a[i]), "r"(b[i])); a[i] = (u32)t; c[i] = (u32)(t >> 32);
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(b[i]) : "r"(a[i]), "r"(c[i]));
            }
            if (MODE == 5) {  // 1 wide + 2 iadd3
                u64 t = mk(a[i], c[i]); asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(t) : "r"(a[i]), "r"(b[i])); a[i] = (u32)t; c[i] = (u32)(t >> 32);
                asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3;" : "+r"(b[i]), "+r"(c[i]) : "r"(a[i]), "r"(seed));
            }
            if (MODE == 6) {  // 1 imad + 1 iadd
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b[i]), "r"(c[i]));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(c[i]) : "r"(a[i]));
            }
            if (MODE == 7) {  // butterfly-like mix: 3 wide + 2 hi + 4 imad + 6 iadd
                u64 t = mk(a[i], c[i]);
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(t) : "r"(a[i]), "r"(b[i]));
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(t) : "r"(b[i]), "r"(c[i]));
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(t) : "r"(c[i]), "r"(a[i]));
                u32 h = (u32)(t >> 32), l = (u32)t;
                asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(h) : "r"(l), "r"(b[i]));
                asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(l) : "r"(h), "r"(c[i]));
                asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(h) : "r"(l), "r"(b[i]));
                asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(l) : "r"(h), "r"(c[i]));
                asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(h) : "r"(l), "r"(a[i]));
                asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(l) : "r"(h), "r"(b[i]));
                asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3;" : "+r"(a[i]), "+r"(c[i]) : "r"(l), "r"(h));
                asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3;" : "+r"(b[i]), "+r"(c[i]) : "r"(h), "r"(l));
                asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3;" : "+r"(a[i]), "+r"(b[i]) : "r"(l), "r"(h));
            }
        }
    }
    long long t1 = clock64();
    u32 acc = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc ^= a[i] ^ b[i] ^ c[i];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

// ---------------------------------------------------------------- (2) butterflies
// exact Shoup product in [0,2p): the compiler's mulhi
__device__ __forceinline__ u64 shoup2(u64 y, u64 w, u64 wp, u64 np) { return y * w + __umul64hi(y, wp) * np; }
// under-estimated quotient (never above, at most 2 below): product in [0,4p)
__device__ __forceinline__ u64 shoup4(u64 y, u64 w, u64 wp, u64 np) {
    const u32 y0 = (u32)y, y1 = (u32)(y >> 32), w0 = (u32)w, w1 = (u32)(w >> 32), wp0 = (u32)wp, wp1 = (u32)(wp >> 32),
              np0 = (u32)np, np1 = (u32)(np >> 32);
    const u64 q = mulwide(y1, wp1) + (u64)__umulhi(y1, wp0) + (u64)__umulhi(y0, wp1);
    const u32 q0 = (u32)q, q1 = (u32)(q >> 32);
    u64 V = mulwide(y0, w0);
    V = madwide(q0, np0, V);
    u32 vh = (u32)(V >> 32);
    vh = y1 * w0 + vh; vh = y0 * w1 + vh; vh = q1 * np0 + vh; vh = q0 * np1 + vh;
    return mk((u32)V, vh);
}
__device__ __forceinline__ u64 csubp(u64 x, u64 m) {  // predicated conditional subtract
    asm("{ .reg .pred q; setp.ge.u64 q, %0, %1; @q sub.u64 %0, %0, %1; }" : "+l"(x) : "l"(m));
    return x;
}

template <int VAR, int ILP, int MINB>
__global__ void __launch_bounds__(128, MINB) bfly(u64 *out, const ulonglong2 *tw, u64 p, long long *clk) {
    u64 a[ILP], b[ILP];
    const ulonglong2 w = tw[threadIdx.x + 128 * (blockIdx.x & 7)];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { a[i] = w.x * (i + 3) + threadIdx.x; b[i] = w.y * (i + 5) + blockIdx.x; }
    const u64 np = 0 - p, p2 = 2 * p, p4 = 4 * p;
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (VAR == 0) { const u64 v = shoup2(b[i], w.x, w.y, np), x = a[i]; a[i] = x + v; b[i] = x - v + p2; }
            if (VAR == 1) { const u64 v = shoup4(b[i], w.x, w.y, np), x = a[i]; a[i] = x + v; b[i] = x - v + p4; }
            if (VAR == 2) { const u64 v = shoup4(b[i], w.x, w.y, np), x = csubp(a[i], p4); a[i] = x + v; b[i] = x - v + p4; }
            if (VAR == 3) { const u64 v = shoup4(b[i], w.x, w.y, np); u64 x = a[i]; x = x >= p4 ? x - p4 : x; a[i] = x + v; b[i] = x - v + p4; }
            if (VAR == 4) { const u64 v = shoup2(b[i], w.x, w.y, np); u64 x = a[i]; x = x >= p2 ? x - p2 : x; a[i] = x + v; b[i] = x - v + p2; }
            if (VAR == 5) { const u64 s = a[i] + b[i]; b[i] = shoup4(a[i] - b[i] + p4, w.x, w.y, np); a[i] = s; }                   // GS narrow
            if (VAR == 6) { const u64 s = csubp(a[i] + b[i], p4); b[i] = shoup4(a[i] - b[i] + p4, w.x, w.y, np); a[i] = s; }       // GS mid
        }
    }
    long long t1 = clock64();
    u64 acc = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc ^= a[i] ^ b[i];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

static int g_sms;
static long long *g_clk, *g_hclk;
static void *g_out;
static ulonglong2 *g_tw;

template <typename K, typename... A>
static void timeit(const char *name, K kern, int bps, double ops_per_thread, const char *unit, A... args) {
    int resident = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, kern, 128, 0);
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, kern);
    if (resident < bps) { printf("%-34s warps/SMSP=%2d : SKIP (only %d blocks resident, %d regs)\n", name, bps, resident, fa.numRegs); return; }
    const int blocks = g_sms * bps;
    kern<<<blocks, 128>>>(args...);
    kern<<<blocks, 128>>>(args...);
    cudaDeviceSynchronize();
    cudaMemcpy(g_hclk, g_clk, sizeof(long long) * blocks, cudaMemcpyDeviceToHost);
    double avg = 0, mx = 0;
    for (int i = 0; i < blocks; ++i) { avg += (double)g_hclk[i]; if (g_hclk[i] > mx) mx = (double)g_hclk[i]; }
    avg /= blocks;
    const double per_clk_sm = (double)bps * 128 * ops_per_thread / avg;
    printf("%-34s warps/SMSP=%2d regs=%3d : %7.2f %s/clk/SM  (%5.3f warp-%s/clk/SMSP; max/avg clk %.2f)\n", name, bps, fa.numRegs,
           per_clk_sm, unit, per_clk_sm / 128, unit, mx / avg);
}

int main() {
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    g_sms = prop.multiProcessorCount;
    printf("device %s, %d SMs\n", prop.name, g_sms);
    cudaMalloc(&g_out, 8 * g_sms * 16 * 128);
    cudaMalloc(&g_clk, sizeof(long long) * g_sms * 16);
    g_hclk = (long long *)malloc(sizeof(long long) * g_sms * 16);
    const u64 p = 36028797018652673ull;
    ulonglong2 *htw = (ulonglong2 *)malloc(sizeof(ulonglong2) * 1024);
    for (int i = 0; i < 1024; ++i) { u64 w = (0x9E3779B97F4A7C15ull * (i + 1)) % p; htw[i].x = w; htw[i].y = (u64)(((unsigned __int128)w << 64) / p); }
    cudaMalloc(&g_tw, sizeof(ulonglong2) * 1024);
    cudaMemcpy(g_tw, htw, sizeof(ulonglong2) * 1024, cudaMemcpyHostToDevice);
    u32 *o32 = (u32 *)g_out; u64 *o64 = (u64 *)g_out;
    for (int w : {4, 8, 16}) {
        timeit("IMAD (mad.lo.u32)", istream<0, 8>, w, (double)ITERS * 8, "inst", o32, 12345u, g_clk);
        timeit("IMAD.HI (mad.hi.u32)", istream<1, 8>, w, (double)ITERS * 8, "inst", o32, 12345u, g_clk);
        timeit("IMAD.WIDE (mad.wide.u32 acc)", istream<2, 8>, w, (double)ITERS * 8, "inst", o32, 12345u, g_clk);
        timeit("IADD3+IADD3.X", istream<3, 8>, w, (double)ITERS * 8 * 2, "inst", o32, 12345u, g_clk);
        timeit("1 WIDE + 1 IMAD", istream<4, 8>, w, (double)ITERS * 8 * 2, "inst", o32, 12345u, g_clk);
        timeit("1 WIDE + 2 IADD3", istream<5, 8>, w, (double)ITERS * 8 * 3, "inst", o32, 12345u, g_clk);
        timeit("1 IMAD + 1 IADD", istream<6, 8>, w, (double)ITERS * 8 * 2, "inst", o32, 12345u, g_clk);
        timeit("mix 3W+2HI+4IMAD+6IADD", istream<7, 4>, w, (double)ITERS * 4 * 15, "inst", o32, 12345u, g_clk);
    }
#define BF(NAME, VAR) \
    timeit(NAME " ILP8", bfly<VAR, 8, 8>, 4, (double)ITERS * 8, "bfly", o64, g_tw, p, g_clk); \
    timeit(NAME " ILP8", bfly<VAR, 8, 8>, 8, (double)ITERS * 8, "bfly", o64, g_tw, p, g_clk); \
    timeit(NAME " ILP4", bfly<VAR, 4, 12>, 8, (double)ITERS * 4, "bfly", o64, g_tw, p, g_clk); \
    timeit(NAME " ILP4", bfly<VAR, 4, 12>, 12, (double)ITERS * 4, "bfly", o64, g_tw, p, g_clk); \
    timeit(NAME " ILP2", bfly<VAR, 2, 16>, 16, (double)ITERS * 2, "bfly", o64, g_tw, p, g_clk);
    BF("CT exact (compiler mulhi) <2p", 0)
    BF("CT approx q (3 heavy) <4p", 1)
    BF("CT approx + pred csub 4p", 2)
    BF("CT approx + ?: csub 4p", 3)
    BF("CT exact + csub 2p (r1 wide)", 4)
    BF("GS approx narrow", 5)
    BF("GS approx + pred csub", 6)
    return 0;
}
