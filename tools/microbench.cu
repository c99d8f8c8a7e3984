// microbench.cu -- integer-pipe throughput probes for the modular-arithmetic kernels (sm_100a).
// Prints lane-ops per clock per SM for the instruction mixes the NTT butterflies are made of.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o microbench microbench.cu && ./microbench
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

typedef unsigned long long u64;
typedef unsigned int u32;

#define ITERS 2048
#define ILP 4

__device__ __forceinline__ u64 shoup_lazy(u64 x, u64 w, u64 wp, u64 p) { return x * w - __umul64hi(x, wp) * p; }
// approximate high product: drops the lo*lo partial product (quotient off by at most 1 -> result in [0, 3p))
__device__ __forceinline__ u64 mulhi_approx(u64 a, u64 b) {
    u32 al = (u32)a, ah = (u32)(a >> 32), bl = (u32)b, bh = (u32)(b >> 32);
    u64 mid = (u64)ah * bl + (((u64)al * bh) >> 32);  // cannot overflow
    return (u64)ah * bh + (mid >> 32) + 0;            // ignoring low 32 bits of al*bh as well would lose more
}
__device__ __forceinline__ u64 shoup_lazy_approx(u64 x, u64 w, u64 wp, u64 p) { return x * w - mulhi_approx(x, wp) * p; }

template <int MODE>
__global__ void probe(u64 *out, u64 seed, u64 p, u64 w, u64 wp, long long *clk) {
    u64 a[ILP], b[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
        a[i] = seed + threadIdx.x * 7919u + i * 104729u + blockIdx.x;
        b[i] = seed * 31 + threadIdx.x + i;
    }
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (MODE == 0) a[i] = a[i] * b[i] + it;                       // mul.lo.u64 (+add)
            if (MODE == 1) a[i] = __umul64hi(a[i], b[i]) + b[i];          // mul.hi.u64
            if (MODE == 2) a[i] = (u64)(u32)a[i] * (u32)b[i] + a[i];      // mad.wide.u32
            if (MODE == 3) { u32 x = (u32)a[i] * (u32)b[i] + (u32)it; a[i] = x; }  // 32-bit imad
            if (MODE == 4) a[i] = a[i] + b[i] + (a[i] >> 3);              // 64-bit adds/shift
            if (MODE == 5) {                                              // Harvey CT butterfly, exact mulhi, no csub
                u64 t = shoup_lazy(b[i], w, wp, p);
                u64 x = a[i];
                a[i] = x + t;
                b[i] = x - t + 2 * p;
            }
            if (MODE == 6) {                                              // butterfly with csub on x (generic p < 2^62)
                u64 x = a[i] >= 2 * p ? a[i] - 2 * p : a[i];
                u64 t = shoup_lazy(b[i], w, wp, p);
                a[i] = x + t;
                b[i] = x - t + 2 * p;
            }
            if (MODE == 7) {                                              // butterfly with approximate mulhi
                u64 t = shoup_lazy_approx(b[i], w, wp, p);
                u64 x = a[i];
                a[i] = x + t;
                b[i] = x - t + 3 * p;
            }
            if (MODE == 8) {                                              // 128-bit MAC
                u64 lo = a[i] * b[i], hi = __umul64hi(a[i], b[i]);
                a[i] += lo;
                b[i] += hi + (a[i] < lo);
            }
            if (MODE == 9) {                                              // GS butterfly with csub
                u64 x = a[i], y = b[i];
                u64 s = x + y;
                a[i] = s >= 2 * p ? s - 2 * p : s;
                b[i] = shoup_lazy(x - y + 2 * p, w, wp, p);
            }
        }
    }
    long long t1 = clock64();
    u64 acc = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc ^= a[i] ^ b[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char *name, int sms, int blocks_per_sm, int threads, u64 *out, long long *clk, long long *hclk) {
    const int blocks = sms * blocks_per_sm;
    const u64 p = 36028797018652673ull, w = 15372713853695ull;
    const u64 wp = (u64)(((unsigned __int128)w << 64) / p);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    probe<MODE><<<blocks, threads>>>(out, 12345, p, w, wp, clk);
    cudaEventRecord(e0);
    probe<MODE><<<blocks, threads>>>(out, 12345, p, w, wp, clk);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaMemcpy(hclk, clk, sizeof(long long) * blocks, cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < blocks; ++i) avg += (double)hclk[i];
    avg /= blocks;
    const double ops = (double)blocks * threads * ITERS * ILP;
    // resident threads per SM = blocks_per_sm * threads, all co-resident: per-SM lane-ops per clock
    const double per_clk_sm = (double)blocks_per_sm * threads * ITERS * ILP / avg;
    printf("%-34s %8.3f ms  %9.1f Gop/s  %7.2f ops/clk/SM  (avg %.0f clk -> %.0f MHz)\n", name, ms, ops / ms / 1e6,
           per_clk_sm, avg, avg / ms / 1e3);
}

int main() {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    printf("device %s, %d SMs\n", prop.name, sms);
    u64 *out;
    long long *clk, *hclk;
    const int bps = 2, threads = 1024;
    cudaMalloc(&out, sizeof(u64) * sms * bps * threads);
    cudaMalloc(&clk, sizeof(long long) * sms * bps);
    hclk = (long long *)malloc(sizeof(long long) * sms * bps);
    run<0>("mul.lo.u64 + add", sms, bps, threads, out, clk, hclk);
    run<1>("mul.hi.u64 + add", sms, bps, threads, out, clk, hclk);
    run<2>("mad.wide.u32", sms, bps, threads, out, clk, hclk);
    run<3>("imad.u32", sms, bps, threads, out, clk, hclk);
    run<4>("add64 x2 + shift", sms, bps, threads, out, clk, hclk);
    run<5>("CT butterfly (exact mulhi, lazy)", sms, bps, threads, out, clk, hclk);
    run<6>("CT butterfly (+csub 2p)", sms, bps, threads, out, clk, hclk);
    run<7>("CT butterfly (approx mulhi)", sms, bps, threads, out, clk, hclk);
    run<8>("mac 64x64->128", sms, bps, threads, out, clk, hclk);
    run<9>("GS butterfly (+csub 2p)", sms, bps, threads, out, clk, hclk);
    return 0;
}
