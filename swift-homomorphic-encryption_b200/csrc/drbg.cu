// drbg.cu -- seeded-ciphertext expansion on the device.
//
//   NistCtrDrbg (CTR_DRBG, AES-128, no derivation function)   Random/NistCtrDrbg.swift:25-84  (NIST SP 800-90A)
//   NistAes128Ctr = BufferedRng<NistCtrDrbg>, 4096-byte buffer Random/NistAes128Ctr.swift:17-40, BufferedRng.swift:17-67
//   PolyRq.randomizeUniform(using:)                            PolyRq/PolyRq+Randomize.swift:49-81
//   Ciphertext(deserialize: .seeded(poly0:seed:))              SerializedCiphertext.swift:41-60
//
// The byte stream a seed produces is a chain of 4096-byte segments: segment s is AES-128-CTR under (key_s, V_s), and
// (key_{s+1}, V_{s+1}) come from two more blocks of the same keystream.  One thread per seed walks that chain (it is
// inherently sequential, 2 block encryptions + 1 key schedule per segment) and leaves the expanded round keys; then
// every 16-byte block of every segment is independent: one CTA per segment, one thread per block = per coefficient
// (a coefficient consumes exactly one little-endian 128-bit word, reduced modulo its row modulus).
// AES is FIPS-197 written from the specification (S-box, ShiftRows, MixColumns over GF(2^8)); the reference gets it
// from swift-crypto.  Pinned by the reference's NIST vectors through the oracle (tests/test_oracle_drbg.py).
#include <algorithm>

#include "capi_internal.hpp"
#include "drbg.cuh"
#include "modarith.cuh"

using namespace hecuda;
using namespace hecuda::api;
using namespace hecuda::drbg;

namespace {

__constant__ unsigned char c_sbox[256];
__constant__ u32w c_te0[256];

__device__ __forceinline__ void load_tables(unsigned char *sbox, u32w *te0) {
    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
        sbox[i] = c_sbox[i];
        te0[i] = c_te0[i];
    }
    __syncthreads();
}

// two lanes per seed walk its chain of segments: both expand the current key, lane p encrypts counter block V + 1 + p,
// the pair exchanges the blocks by shuffle and absorbs them (ctrDrbgUpdate); lane 0 leaves (round keys, V) of each segment
__global__ void __launch_bounds__(64) drbg_chain_kernel(const unsigned char *__restrict__ seeds, u32w *__restrict__ round_keys,
                                                        u64 *__restrict__ counters, int segments, long long batch) {
    __shared__ unsigned char sbox[256];
    __shared__ u32w te0[256];
    load_tables(sbox, te0);
    const long long b = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 1;
    const int lane = threadIdx.x & 1;
    const bool live = b < batch;  // dead pairs still take part in the shuffles
    const long long seed = live ? b : 0;
    u32w key[4] = {0, 0, 0, 0}, rk[kRoundKeyWords], blk[4], b0[4], b1[4], provided[8];
    u64 hi = 0, lo = 0;
    for (int i = 0; i < 8; ++i) {
        const unsigned char *p = seeds + 32 * seed + 4 * i;
        provided[i] = ((u32w)p[0] << 24) | ((u32w)p[1] << 16) | ((u32w)p[2] << 8) | p[3];
    }
    for (int s = -1; s < segments; ++s) {  // s = -1: init(entropy:) (NistCtrDrbg.swift:52-59)
        expand_key(key, rk, sbox);
        if (s >= 0) {
            if (live && lane == 0) {
                u32w *dst = round_keys + ((size_t)b * segments + s) * kRoundKeyWords;
                for (int i = 0; i < kRoundKeyWords; ++i) dst[i] = rk[i];
                counters[2 * ((size_t)b * segments + s)] = hi;
                counters[2 * ((size_t)b * segments + s) + 1] = lo;
            }
            const u64 l = lo + kSegmentBlocks;  // ctrDrbgGenerate(count: 4096) advances V by 256 blocks (:71-84)
            hi += l < lo ? 1 : 0;
            lo = l;
        }
        counter_block(hi, lo, 1 + (u64)lane, blk);
        encrypt_block(blk, rk, te0, sbox);
        for (int i = 0; i < 4; ++i) {
            b0[i] = __shfl_sync(0xffffffffu, blk[i], (threadIdx.x & 30), 32);
            b1[i] = __shfl_sync(0xffffffffu, blk[i], (threadIdx.x & 30) | 1, 32);
        }
        drbg_absorb(key, hi, lo, b0, b1, s < 0 ? provided : nullptr);
    }
}

struct FillConsts {
    int rows;
    u64 p[kMaxRows];
};

// one CTA per segment, one thread per 16-byte block = per coefficient (randomizeUniform, PolyRq+Randomize.swift:58-80)
__global__ void __launch_bounds__(kSegmentBlocks) drbg_fill_kernel(const u32w *__restrict__ round_keys,
                                                                   const u64 *__restrict__ counters, u64 *__restrict__ out,
                                                                   const __grid_constant__ FillConsts c, int n, int segments) {
    __shared__ unsigned char sbox[256];
    __shared__ u32w te0[256];
    __shared__ u32w rk[kRoundKeyWords];
    const long long b = blockIdx.y;
    const int s = blockIdx.x;
    if (threadIdx.x < kRoundKeyWords) rk[threadIdx.x] = round_keys[((size_t)b * segments + s) * kRoundKeyWords + threadIdx.x];
    load_tables(sbox, te0);
    const long long k = (long long)s * kSegmentBlocks + threadIdx.x;
    if (k >= (long long)c.rows * n) return;
    u32w blk[4];
    counter_block(counters[2 * ((size_t)b * segments + s)], counters[2 * ((size_t)b * segments + s) + 1], 1 + (u64)threadIdx.x, blk);
    encrypt_block(blk, rk, te0, sbox);
    // UInt128(littleEndianBytes:) of the 16 output bytes: byte i of the block is bits 8i.. of the value
    const u64 lo = (u64)__byte_perm(blk[0], 0, 0x0123) | ((u64)__byte_perm(blk[1], 0, 0x0123) << 32);
    const u64 hi = (u64)__byte_perm(blk[2], 0, 0x0123) | ((u64)__byte_perm(blk[3], 0, 0x0123) << 32);
    const int row = (int)(k / n);
    out[(size_t)b * c.rows * n + k] = (u64)((((u128)hi << 64) | lo) % c.p[row]);
}

cudaError_t random_polys_device(const Context &c, int l, const unsigned char *d_seeds, u64 *d_out, int64_t batch,
                                cudaStream_t s) {
    int device = 0;
    cudaGetDevice(&device);
    unsigned char sbox[256];
    u32w te0[256];
    make_tables(sbox, te0);
    cudaError_t e = cudaMemcpyToSymbolAsync(c_sbox, sbox, sizeof(sbox), 0, cudaMemcpyHostToDevice, s);  // per device, cheap
    if (e == cudaSuccess) e = cudaMemcpyToSymbolAsync(c_te0, te0, sizeof(te0), 0, cudaMemcpyHostToDevice, s);
    if (e != cudaSuccess) return e;
    e = cudaStreamSynchronize(s);  // the tables live on this stack frame
    if (e != cudaSuccess) return e;
    const int segments = (int)(((size_t)l * c.n * 16 + kSegmentBytes - 1) / kSegmentBytes);
    u32w *d_rk = nullptr;
    u64 *d_ctr = nullptr;
    e = cudaMallocAsync((void **)&d_rk, (size_t)batch * segments * kRoundKeyWords * sizeof(u32w), s);
    if (e == cudaSuccess) e = cudaMallocAsync((void **)&d_ctr, (size_t)batch * segments * 2 * sizeof(u64), s);
    if (e == cudaSuccess) {
        ++g_kernel_launches;
        drbg_chain_kernel<<<(unsigned)((batch + 31) / 32), 64, 0, s>>>(d_seeds, d_rk, d_ctr, segments, batch);
        e = cudaGetLastError();
    }
    FillConsts fc;
    fc.rows = l;
    for (int r = 0; r < l; ++r) fc.p[r] = c.slots[c.slot_q(r)].dev.p;
    for (int64_t done = 0; e == cudaSuccess && done < batch;) {
        const int64_t part = std::min<int64_t>(batch - done, 65535);
        ++g_kernel_launches;
        drbg_fill_kernel<<<dim3((unsigned)segments, (unsigned)part), kSegmentBlocks, 0, s>>>(
            d_rk + (size_t)done * segments * kRoundKeyWords, d_ctr + (size_t)done * segments * 2, d_out + (size_t)done * l * c.n, fc,
            (int)c.n, segments);
        e = cudaGetLastError();
        done += part;
    }
    if (d_rk) {
        cudaMemsetAsync(d_rk, 0, (size_t)batch * segments * kRoundKeyWords * sizeof(u32w), s);  // key material
        cudaFreeAsync(d_rk, s);
    }
    if (d_ctr) cudaFreeAsync(d_ctr, s);
    return e;
}

int32_t check(const hecuda_context *h, const void *seeds, int32_t l, const void *out, int64_t batch) {
    int32_t rc = check_ctx(h);
    if (rc) return rc;
    if (l < 1 || l > h->ctx->L) return fail(HECUDA_ERR_INVALID_ARGUMENT, "invalidPolyContext: moduli_count out of range");
    if (batch < 0 || (batch && (!seeds || !out))) return fail(HECUDA_ERR_INVALID_ARGUMENT, "null buffer");
    return HECUDA_OK;
}

}  // namespace

namespace hecuda {
namespace api {

// Ciphertext(deserialize: .seeded) for `batch` ciphertexts, all buffers on the device; d_out: batch x 2 x l x N (Coeff)
cudaError_t expand_seeded_device(const Context &c, int l, const unsigned char *d_poly0, const unsigned char *d_seeds, u64 *d_out,
                                 int64_t batch, cudaStream_t s) {
    const NttRowMap map = c.map_q(l);
    CodecConsts cc;
    std::string err;
    if (!codec_consts(c, map, 0, cc, err)) return cudaErrorInvalidValue;
    const size_t poly_words = (size_t)l * c.n, pw = poly_words * sizeof(u64);
    u64 *d_a = nullptr, *d_p0 = nullptr;
    cudaError_t e = cudaMallocAsync((void **)&d_a, pw * batch, s);
    if (e == cudaSuccess) e = cudaMallocAsync((void **)&d_p0, pw * batch, s);
    // poly0: PolyRq(deserialize:) ; poly1: random Eval polynomial converted to Coeff (SerializedCiphertext.swift:44-49)
    if (e == cudaSuccess) e = launch_poly_load(c, cc, 0, d_poly0, d_p0, batch, s);
    if (e == cudaSuccess) e = random_polys_device(c, l, d_seeds, d_a, batch, s);
    if (e == cudaSuccess) e = launch_ntt_inverse(c, map, d_a, d_a, batch * l, kScalePlain, s);
    if (e == cudaSuccess) e = cudaMemcpy2DAsync(d_out, 2 * pw, d_p0, pw, pw, (size_t)batch, cudaMemcpyDeviceToDevice, s);
    if (e == cudaSuccess) e = cudaMemcpy2DAsync(d_out + poly_words, 2 * pw, d_a, pw, pw, (size_t)batch, cudaMemcpyDeviceToDevice, s);
    if (d_a) cudaFreeAsync(d_a, s);
    if (d_p0) cudaFreeAsync(d_p0, s);
    return e;
}

}  // namespace api
}  // namespace hecuda

extern "C" {

int32_t hecuda_poly_random_from_seed(const hecuda_context *h, const uint8_t *seeds, int32_t l, uint64_t *out, int64_t batch) {
    int32_t rc = check(h, seeds, l, out, batch);
    if (rc || batch == 0) return rc;
    const Context &c = *h->ctx;
    WsGuard g(h);
    if (!g.w) return fail(HECUDA_ERR_CUDA, "could not create a CUDA stream / workspace");
    cudaStream_t s = g.w->stream;
    unsigned char *d_seeds = nullptr;
    u64 *d_out = nullptr;
    const size_t words = (size_t)l * c.n * batch;
    cudaError_t e = cudaMallocAsync((void **)&d_seeds, (size_t)32 * batch, s);
    if (e == cudaSuccess) e = cudaMallocAsync((void **)&d_out, words * sizeof(u64), s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_seeds, seeds, (size_t)32 * batch, cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = random_polys_device(c, l, d_seeds, d_out, batch, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_out, words * sizeof(u64), cudaMemcpyDeviceToHost, s);
    if (d_seeds) cudaFreeAsync(d_seeds, s);
    if (d_out) cudaFreeAsync(d_out, s);
    cudaError_t e2 = cudaStreamSynchronize(s);
    if (e == cudaSuccess) e = e2;
    return e == cudaSuccess ? HECUDA_OK : cuda_fail(e, "poly_random_from_seed");
}

int32_t hecuda_ciphertext_expand_seeded(const hecuda_context *h, const uint8_t *poly0, const uint8_t *seeds, int32_t l,
                                        uint64_t *out, int64_t batch) {
    int32_t rc = check(h, seeds, l, out, batch);
    if (rc) return rc;
    if (batch && !poly0) return fail(HECUDA_ERR_INVALID_ARGUMENT, "null buffer");
    if (batch == 0) return HECUDA_OK;
    const Context &c = *h->ctx;
    const NttRowMap map = c.map_q(l);
    CodecConsts cc;
    std::string err;
    if (!codec_consts(c, map, 0, cc, err)) return fail(HECUDA_ERR_INVALID_ARGUMENT, err);
    const size_t poly_bytes = (size_t)serialized_poly_bytes(cc), poly_words = (size_t)l * c.n;
    WsGuard g(h);
    if (!g.w) return fail(HECUDA_ERR_CUDA, "could not create a CUDA stream / workspace");
    cudaStream_t s = g.w->stream;
    unsigned char *d_seeds = nullptr, *d_poly0 = nullptr;
    u64 *d_out = nullptr;
    cudaError_t e = cudaMallocAsync((void **)&d_seeds, (size_t)32 * batch, s);
    if (e == cudaSuccess) e = cudaMallocAsync((void **)&d_poly0, poly_bytes * batch, s);
    if (e == cudaSuccess) e = cudaMallocAsync((void **)&d_out, 2 * poly_words * batch * sizeof(u64), s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_seeds, seeds, (size_t)32 * batch, cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_poly0, poly0, poly_bytes * batch, cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = expand_seeded_device(c, l, d_poly0, d_seeds, d_out, batch, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_out, 2 * poly_words * sizeof(u64) * batch, cudaMemcpyDeviceToHost, s);
    for (void *p : {(void *)d_seeds, (void *)d_poly0, (void *)d_out})
        if (p) cudaFreeAsync(p, s);
    cudaError_t e2 = cudaStreamSynchronize(s);
    if (e == cudaSuccess) e = e2;
    return e == cudaSuccess ? HECUDA_OK : cuda_fail(e, "expand_seeded");
}

}  // extern "C"
