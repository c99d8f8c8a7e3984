// ntt_simple.cu -- generic shared-memory radix-2 negacyclic NTT (any power-of-two N >= 2 that fits in shared memory).
//
// This is the correctness baseline and the small-N path (the reference's KATs run at N = 2..32,
// Tests/HomomorphicEncryptionTests/NttTests.swift:73-180).  One CTA per row; the row lives in shared memory;
// one __syncthreads per stage.  The fast path for N = 4096/8192/16384 is in ntt_fast.cu.
//
// Forward (Cooley-Tukey, natural -> bit-reversed, _NttContext.forwardNtt PolyRq+Ntt.swift:237-319):
//   values are kept in [0, 4p) (Harvey), the last stage writes canonical residues.
// Inverse (Gentleman-Sande, bit-reversed -> natural, PolyRq+Ntt.swift:379-483): values in [0, 2p), the last
//   stage multiplies by N^-1 and N^-1 psi^-(N/2) (optionally times t) and writes canonical residues.
#include "kernels.cuh"

namespace hecuda {

template <bool INVERSE>
__global__ void ntt_simple_kernel(const u64 *__restrict__ in, u64 *__restrict__ out, const ModSlot *__restrict__ slots,
                                  NttRowMap map, int logn, int scale_mode) {
    extern __shared__ u64 sm[];
    const int64_t row = blockIdx.x;
    const int n = 1 << logn;
    const ModSlot &S = slots[map.slot[(row % map.rows_per_poly) / map.group]];
    const u64 p = S.p, two_p = 2 * p;
    const u64 *src = in + row * n;
    u64 *dst = out + row * n;
    if (!INVERSE && map.src_mod) {  // key-switch digit: [target row j]_{m_r}  (Bfv+Keys.swift:165-172)
        const int idx = (int)(row % map.rows_per_poly);
        src = in + (row / map.rows_per_poly) * map.src_poly_stride + (int64_t)(idx % map.src_mod) * n;
        for (int i = threadIdx.x; i < n; i += blockDim.x) sm[i] = barrett64(src[i], p, S.mu1);
    } else {
        for (int i = threadIdx.x; i < n; i += blockDim.x) sm[i] = src[i];
    }
    __syncthreads();
    const int half = n >> 1;
    if (!INVERSE) {
        const ulonglong2 *__restrict__ tw = S.tw;
        for (int s = 0; s < logn; ++s) {
            const int m = 1 << s, lt = logn - 1 - s, t = 1 << lt;
            const bool last = (s == logn - 1);
            for (int b = threadIdx.x; b < half; b += blockDim.x) {
                const int i = b >> lt, j = b & (t - 1);
                const int idx = (i << (lt + 1)) + j;
                const ulonglong2 w = tw[m + i];
                u64 x = csub(sm[idx], two_p);
                const u64 v = shoup_lazy(sm[idx + t], w.x, w.y, p);
                u64 xo = x + v, yo = x - v + two_p;  // [0, 4p)
                if (last) {
                    xo = csub(csub(xo, two_p), p);
                    yo = csub(csub(yo, two_p), p);
                }
                sm[idx] = xo;
                sm[idx + t] = yo;
            }
            __syncthreads();
        }
    } else {
        const ulonglong2 *__restrict__ tw = S.itw;
        for (int s = logn - 1; s >= 1; --s) {
            const int m = 1 << s, lt = logn - 1 - s, t = 1 << lt;
            for (int b = threadIdx.x; b < half; b += blockDim.x) {
                const int i = b >> lt, j = b & (t - 1);
                const int idx = (i << (lt + 1)) + j;
                const ulonglong2 w = tw[m + i];
                const u64 x = sm[idx], y = sm[idx + t];  // [0, 2p)
                sm[idx] = csub(x + y, two_p);
                sm[idx + t] = shoup_lazy(x - y + two_p, w.x, w.y, p);
            }
            __syncthreads();
        }
        const u64 c0 = S.inv_scale[scale_mode].c0, c0p = S.inv_scale[scale_mode].c0p;
        const u64 c1 = S.inv_scale[scale_mode].c1, c1p = S.inv_scale[scale_mode].c1p;
        for (int b = threadIdx.x; b < half; b += blockDim.x) {
            const u64 x = sm[b], y = sm[b + half];
            sm[b] = shoup_mul(x + y, c0, c0p, p);
            sm[b + half] = shoup_mul(x - y + two_p, c1, c1p, p);
        }
        __syncthreads();
    }
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = sm[i];
}

template <bool INVERSE>
static cudaError_t launch_simple(const Context &ctx, const NttRowMap &map, const u64 *in, u64 *out, int64_t rows,
                                 int scale_mode, cudaStream_t stream) {
    if (rows == 0) return cudaSuccess;
    const size_t smem = sizeof(u64) * (size_t)ctx.n;
    if (smem > 200 * 1024) return cudaErrorInvalidValue;
    if (smem > 48 * 1024) {  // per-device attribute; cheap enough to set on every launch
        cudaError_t e = cudaFuncSetAttribute(ntt_simple_kernel<INVERSE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)smem);
        if (e != cudaSuccess) return e;
    }
    int threads = (int)(ctx.n / 2);
    if (threads > 512) threads = 512;
    if (threads < 32) threads = 32;
    ++g_kernel_launches;
    ntt_simple_kernel<INVERSE><<<(unsigned)rows, threads, smem, stream>>>(in, out, ctx.d_slots, map, ctx.logn,
                                                                       scale_mode);
    return cudaGetLastError();
}

cudaError_t launch_ntt_forward_simple(const Context &ctx, const NttRowMap &map, const u64 *in, u64 *out, int64_t rows,
                                      cudaStream_t stream) {
    return launch_simple<false>(ctx, map, in, out, rows, 0, stream);
}
cudaError_t launch_ntt_inverse_simple(const Context &ctx, const NttRowMap &map, const u64 *in, u64 *out, int64_t rows,
                                      int scale_mode, cudaStream_t stream) {
    return launch_simple<true>(ctx, map, in, out, rows, scale_mode, stream);
}

}  // namespace hecuda
