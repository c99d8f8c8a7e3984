// microbench2.cu -- butterfly throughput vs occupancy and ILP (issue-bound or latency-bound?)
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
typedef unsigned long long u64;
#define ITERS 1024

template <int ILP, int MODE>
__global__ void probe(u64 *out, u64 seed, u64 p, u64 np, u64 w, u64 wp, long long *clk) {
    u64 a[ILP], b[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { a[i] = seed + threadIdx.x * 7919u + i * 104729u + blockIdx.x; b[i] = seed * 31 + threadIdx.x + i; }
    const u64 two_p = 2 * p;
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (MODE == 0) {  // narrow CT butterfly, np form
                u64 t = b[i] * w + __umul64hi(b[i], wp) * np;
                u64 x = a[i];
                a[i] = x + t; b[i] = x - t + two_p;
            } else if (MODE == 1) {  // wide CT butterfly (csub)
                u64 x = a[i] >= two_p ? a[i] - two_p : a[i];
                u64 t = b[i] * w + __umul64hi(b[i], wp) * np;
                a[i] = x + t; b[i] = x - t + two_p;
            } else if (MODE == 2) {  // wide CT butterfly, csub via min trick
                u64 d = a[i] - two_p; u64 x = d < a[i] ? d : a[i];   // unsigned wrap: if a<2p, d is huge
                u64 t = b[i] * w + __umul64hi(b[i], wp) * np;
                a[i] = x + t; b[i] = x - t + two_p;
            }
        }
    }
    long long t1 = clock64();
    u64 acc = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc ^= a[i] ^ b[i];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int ILP, int MODE>
void run(const char *name, int sms, int warps_per_smsp, u64 *out, long long *clk, long long *hclk) {
    const int threads = 128;                      // 4 warps per block -> 1 per SMSP
    const int bps = warps_per_smsp;               // blocks per SM
    const int blocks = sms * bps;
    const u64 p = 36028797018652673ull, w = 15372713853695ull;
    const u64 wp = (u64)(((unsigned __int128)w << 64) / p);
    probe<ILP, MODE><<<blocks, threads>>>(out, 12345, p, 0 - p, w, wp, clk);
    probe<ILP, MODE><<<blocks, threads>>>(out, 12345, p, 0 - p, w, wp, clk);
    cudaDeviceSynchronize();
    cudaMemcpy(hclk, clk, sizeof(long long) * blocks, cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < blocks; ++i) avg += (double)hclk[i];
    avg /= blocks;
    printf("%-22s ILP=%d warps/SMSP=%2d : %6.2f butterflies/clk/SM\n", name, ILP, warps_per_smsp,
           (double)bps * threads * ITERS * ILP / avg);
}

int main() {
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    u64 *out; long long *clk, *hclk;
    cudaMalloc(&out, sizeof(u64) * sms * 16 * 128);
    cudaMalloc(&clk, sizeof(long long) * sms * 16);
    hclk = (long long *)malloc(sizeof(long long) * sms * 16);
    for (int w : {2, 4, 6, 8, 12, 16}) run<4, 0>("narrow np", sms, w, out, clk, hclk);
    for (int w : {2, 4, 6, 8, 12, 16}) run<8, 0>("narrow np", sms, w, out, clk, hclk);
    for (int w : {4, 8, 16}) run<8, 1>("wide csub", sms, w, out, clk, hclk);
    for (int w : {4, 8, 16}) run<8, 2>("wide csub(min)", sms, w, out, clk, hclk);
    return 0;
}
