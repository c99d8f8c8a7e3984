// codec.cu -- the reference's wire format for RNS polynomials on the device.
//
//   PolyRq.serialize(skipLSBs:) / load(from:skipLSBs:)         PolyRq/PolyRq+Serialize.swift:28-84
//   CoefficientPacking.coefficientsToBytes / bytesToCoefficients CoefficientPacking.swift:59-217
//
// Row i of a polynomial is a big-endian bit stream of N fields of ceil(log2 q_i) - skipLSBs bits, padded with zero
// bits to a whole byte; the rows follow each other.  Unpacking on the device lets query ciphertexts cross PCIe at
// ceil(log2 q) bits per coefficient instead of 64, and responses leave the same way.
#include <algorithm>
#include <cstdint>

#include "kernels.cuh"
#include "modarith.cuh"

namespace hecuda {


// bytes -> coefficients: one thread per coefficient
__global__ void __launch_bounds__(256) poly_load_kernel(const unsigned char *__restrict__ bytes, u64 *__restrict__ out,
                                                       const __grid_constant__ CodecConsts c, int n, int skip) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int row = blockIdx.y;
    const int64_t poly = blockIdx.z;
    const int w = c.width[row];
    const long long row_bytes = c.byte_offset[row + 1] - c.byte_offset[row];
    const unsigned char *src = bytes + poly * c.byte_offset[c.rows] + c.byte_offset[row];
    const long long bit = (long long)i * w;
    const long long first = bit >> 3;
    const int shift = (int)(bit & 7);
    u128 acc = 0;  // 9 bytes cover shift + w <= 7 + 64 bits
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        const long long at = first + k;
        acc = (acc << 8) | (u128)(at < row_bytes ? src[at] : 0);
    }
    const u64 mask = w >= 64 ? ~0ull : ((1ull << w) - 1);
    const u64 v = (u64)(acc >> (72 - shift - w)) & mask;
    out[(poly * c.rows + row) * n + i] = v << skip;
}

// coefficients -> bytes: one thread per output byte
__global__ void __launch_bounds__(256) poly_serialize_kernel(const u64 *__restrict__ in, unsigned char *__restrict__ bytes,
                                                            const __grid_constant__ CodecConsts c, int n, int skip) {
    const int row = blockIdx.y;
    const int64_t poly = blockIdx.z;
    const long long row_bytes = c.byte_offset[row + 1] - c.byte_offset[row];
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= row_bytes) return;
    const int w = c.width[row];
    const u64 mask = w >= 64 ? ~0ull : ((1ull << w) - 1);
    const u64 *src = in + (poly * c.rows + row) * n;
    const long long lo_bit = 8 * j, hi_bit = lo_bit + 8;
    unsigned value = 0;
    for (long long coeff = lo_bit / w; coeff < n && coeff * w < hi_bit; ++coeff) {
        const long long begin = coeff * w, end = begin + w;
        const long long lo = begin > lo_bit ? begin : lo_bit, hi = end < hi_bit ? end : hi_bit;
        const u64 v = (src[coeff] >> skip) & mask;
        const unsigned field = (unsigned)((v >> (end - hi)) & ((1ull << (hi - lo)) - 1));
        value |= field << (hi_bit - hi);
    }
    bytes[poly * c.byte_offset[c.rows] + c.byte_offset[row] + j] = (unsigned char)value;
}

static int ceil_log2_u64(u64 q) {  // T.ceilLog2
    int bits = 0;
    while (bits < 64 && (q - 1) >> bits) ++bits;
    return q <= 1 ? 0 : bits;
}

bool codec_consts(const Context &ctx, const NttRowMap &map, int skip, CodecConsts &c, std::string &err) {
    c.rows = map.rows_per_poly;
    c.byte_offset[0] = 0;
    for (int r = 0; r < c.rows; ++r) {
        const int bits = ceil_log2_u64(ctx.slots[map.slot[r]].dev.p);
        if (!(bits > 0 && bits > skip && skip >= 0)) {  // CoefficientPacking.validate (CoefficientPacking.swift:26-30)
            err = "invalidCoefficientPacking(bitsPerCoeff: " + std::to_string(bits) + ", skipLSBs: " + std::to_string(skip) + ")";
            return false;
        }
        c.width[r] = bits - skip;
        c.byte_offset[r + 1] = c.byte_offset[r] + ((long long)ctx.n * c.width[r] + 7) / 8;
    }
    return true;
}

long long serialized_poly_bytes(const CodecConsts &c) { return c.byte_offset[c.rows]; }

cudaError_t launch_poly_load(const Context &ctx, const CodecConsts &c, int skip, const unsigned char *bytes, u64 *out,
                             int64_t polys, cudaStream_t stream) {
    const int threads = ctx.n >= 256 ? 256 : (ctx.n < 32 ? 32 : (int)ctx.n);
    for (int64_t done = 0; done < polys;) {
        const int64_t chunk = (polys - done) > 65535 ? 65535 : (polys - done);
        dim3 grid((unsigned)((ctx.n + threads - 1) / threads), (unsigned)c.rows, (unsigned)chunk);
        ++g_kernel_launches;
        poly_load_kernel<<<grid, threads, 0, stream>>>(bytes + done * c.byte_offset[c.rows], out + done * c.rows * ctx.n, c,
                                                       (int)ctx.n, skip);
        done += chunk;
    }
    return cudaGetLastError();
}

// coefficients -> bytes, 8 output bytes per thread (rows whose byte count and offset are multiples of 8: N >= 64)
__global__ void __launch_bounds__(256) poly_serialize_words_kernel(const u64 *__restrict__ in, unsigned char *__restrict__ bytes,
                                                                  const __grid_constant__ CodecConsts c, int n, int skip) {
    const int row = blockIdx.y;
    const int64_t poly = blockIdx.z;
    const long long row_words = (c.byte_offset[row + 1] - c.byte_offset[row]) >> 3;
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= row_words) return;
    const int w = c.width[row];
    const u64 mask = w >= 64 ? ~0ull : ((1ull << w) - 1);
    const u64 *src = in + (poly * c.rows + row) * n;
    const long long lo_bit = 64 * j, hi_bit = lo_bit + 64;
    u64 value = 0;  // big-endian bit stream: stream bit lo_bit is the MSB of `value`
    for (long long coeff = lo_bit / w; coeff < n && coeff * w < hi_bit; ++coeff) {
        const long long begin = coeff * w, end = begin + w;
        const long long lo = begin > lo_bit ? begin : lo_bit, hi = end < hi_bit ? end : hi_bit;
        const int bits = (int)(hi - lo);
        const u64 v = (src[coeff] >> skip) & mask;
        const u64 field = (v >> (end - hi)) & (bits >= 64 ? ~0ull : ((1ull << bits) - 1));
        value |= bits >= 64 ? field : field << (hi_bit - hi);
    }
    // store most significant byte first
    const u64 swapped = __byte_perm((unsigned)(value >> 32), 0, 0x0123) | ((u64)__byte_perm((unsigned)value, 0, 0x0123) << 32);
    *reinterpret_cast<u64 *>(bytes + poly * c.byte_offset[c.rows] + c.byte_offset[row] + 8 * j) = swapped;
}

cudaError_t launch_poly_serialize(const Context &ctx, const CodecConsts &c, int skip, const u64 *in, unsigned char *bytes,
                                  int64_t polys, cudaStream_t stream) {
    long long widest = 0;
    bool words = (reinterpret_cast<uintptr_t>(bytes) & 7) == 0 && (c.byte_offset[c.rows] & 7) == 0;
    for (int r = 0; r < c.rows; ++r) {
        widest = std::max(widest, c.byte_offset[r + 1] - c.byte_offset[r]);
        words = words && (c.byte_offset[r] & 7) == 0 && (c.byte_offset[r + 1] & 7) == 0;
    }
    for (int64_t done = 0; done < polys;) {
        const int64_t chunk = (polys - done) > 65535 ? 65535 : (polys - done);
        ++g_kernel_launches;
        if (words) {
            dim3 grid((unsigned)((widest / 8 + 255) / 256), (unsigned)c.rows, (unsigned)chunk);
            poly_serialize_words_kernel<<<grid, 256, 0, stream>>>(in + done * c.rows * ctx.n, bytes + done * c.byte_offset[c.rows],
                                                                  c, (int)ctx.n, skip);
        } else {
            dim3 grid((unsigned)((widest + 255) / 256), (unsigned)c.rows, (unsigned)chunk);
            poly_serialize_kernel<<<grid, 256, 0, stream>>>(in + done * c.rows * ctx.n, bytes + done * c.byte_offset[c.rows], c,
                                                            (int)ctx.n, skip);
        }
        done += chunk;
    }
    return cudaGetLastError();
}

}  // namespace hecuda

// ------------------------------------------------------------------------------------------------ C ABI
#include "capi_internal.hpp"

using namespace hecuda;
using namespace hecuda::api;

namespace {

int32_t codec_setup(const hecuda_context *h, int32_t base, int32_t rows, int32_t skip, const void *a, const void *b,
                    int64_t polys, CodecConsts &c) {
    int32_t rc = check_ctx(h);
    if (rc) return rc;
    if (polys < 0 || (polys && (!a || !b))) return fail(HECUDA_ERR_INVALID_ARGUMENT, "null buffer");
    NttRowMap map;
    std::string err;
    if (!make_map(*h->ctx, base, rows, map, err)) return fail(HECUDA_ERR_INVALID_ARGUMENT, err);
    if (!codec_consts(*h->ctx, map, skip, c, err)) return fail(HECUDA_ERR_INVALID_ARGUMENT, err);
    return HECUDA_OK;
}

}  // namespace

extern "C" {

int32_t hecuda_poly_serialized_byte_count(const hecuda_context *h, int32_t base, int32_t rows, int32_t skip_lsbs,
                                          uint64_t *bytes) {
    CodecConsts c;
    if (!bytes) return fail(HECUDA_ERR_INVALID_ARGUMENT, "null argument");
    int32_t rc = codec_setup(h, base, rows, skip_lsbs, nullptr, nullptr, 0, c);
    if (rc) return rc;
    *bytes = (uint64_t)serialized_poly_bytes(c);
    return HECUDA_OK;
}

int32_t hecuda_poly_load_device(const hecuda_context *h, int32_t base, const uint8_t *serialized, int32_t skip_lsbs,
                                uint64_t *out, int32_t rows, int64_t polys, void *stream) {
    CodecConsts c;
    int32_t rc = codec_setup(h, base, rows, skip_lsbs, serialized, out, polys, c);
    if (rc || polys == 0) return rc;
    cudaError_t e = launch_poly_load(*h->ctx, c, skip_lsbs, serialized, (u64 *)out, polys, (cudaStream_t)stream);
    return e == cudaSuccess ? HECUDA_OK : cuda_fail(e, "poly_load");
}

int32_t hecuda_poly_serialize_device(const hecuda_context *h, int32_t base, const uint64_t *in, int32_t skip_lsbs,
                                     uint8_t *serialized, int32_t rows, int64_t polys, void *stream) {
    CodecConsts c;
    int32_t rc = codec_setup(h, base, rows, skip_lsbs, in, serialized, polys, c);
    if (rc || polys == 0) return rc;
    cudaError_t e = launch_poly_serialize(*h->ctx, c, skip_lsbs, (const u64 *)in, serialized, polys, (cudaStream_t)stream);
    return e == cudaSuccess ? HECUDA_OK : cuda_fail(e, "poly_serialize");
}

int32_t hecuda_poly_load(const hecuda_context *h, int32_t base, const uint8_t *serialized, int32_t skip_lsbs, uint64_t *out,
                         int32_t rows, int64_t polys) {
    CodecConsts c;
    int32_t rc = codec_setup(h, base, rows, skip_lsbs, serialized, out, polys, c);
    if (rc || polys == 0) return rc;
    WsGuard g(h);
    if (!g.w) return fail(HECUDA_ERR_CUDA, "could not create a CUDA stream / workspace");
    const size_t in_bytes = (size_t)serialized_poly_bytes(c) * polys, out_words = (size_t)rows * h->ctx->n * polys;
    cudaStream_t s = g.w->stream;
    unsigned char *d_in = nullptr;
    u64 *d_out = nullptr;
    cudaError_t e = cudaMallocAsync((void **)&d_in, in_bytes, s);
    if (e == cudaSuccess) e = cudaMallocAsync((void **)&d_out, out_words * sizeof(u64), s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_in, serialized, in_bytes, cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = launch_poly_load(*h->ctx, c, skip_lsbs, d_in, d_out, polys, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_out, out_words * sizeof(u64), cudaMemcpyDeviceToHost, s);
    if (d_in) cudaFreeAsync(d_in, s);
    if (d_out) cudaFreeAsync(d_out, s);
    const cudaError_t e2 = cudaStreamSynchronize(s);
    if (e == cudaSuccess) e = e2;
    return e == cudaSuccess ? HECUDA_OK : cuda_fail(e, "poly_load");
}

int32_t hecuda_poly_serialize(const hecuda_context *h, int32_t base, const uint64_t *in, int32_t skip_lsbs,
                              uint8_t *serialized, int32_t rows, int64_t polys) {
    CodecConsts c;
    int32_t rc = codec_setup(h, base, rows, skip_lsbs, in, serialized, polys, c);
    if (rc || polys == 0) return rc;
    WsGuard g(h);
    if (!g.w) return fail(HECUDA_ERR_CUDA, "could not create a CUDA stream / workspace");
    const size_t out_bytes = (size_t)serialized_poly_bytes(c) * polys, in_words = (size_t)rows * h->ctx->n * polys;
    cudaStream_t s = g.w->stream;
    unsigned char *d_out = nullptr;
    u64 *d_in = nullptr;
    cudaError_t e = cudaMallocAsync((void **)&d_in, in_words * sizeof(u64), s);
    if (e == cudaSuccess) e = cudaMallocAsync((void **)&d_out, out_bytes, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_in, in, in_words * sizeof(u64), cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = launch_poly_serialize(*h->ctx, c, skip_lsbs, d_in, d_out, polys, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(serialized, d_out, out_bytes, cudaMemcpyDeviceToHost, s);
    if (d_in) cudaFreeAsync(d_in, s);
    if (d_out) cudaFreeAsync(d_out, s);
    const cudaError_t e2 = cudaStreamSynchronize(s);
    if (e == cudaSuccess) e = e2;
    return e == cudaSuccess ? HECUDA_OK : cuda_fail(e, "poly_serialize");
}

}  // extern "C"
