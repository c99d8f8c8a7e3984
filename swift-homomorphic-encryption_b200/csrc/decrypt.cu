// decrypt.cu -- BFV decryption on the device.
//
//   Bfv.decryptCoeff / decryptEval     Bfv/Bfv+Decrypt.swift:21-41   (dotProduct(ciphertext:with:) :188-204)
//   RnsTool.scaleAndRound              RnsTool.swift:272-302         (BEHZ Algorithm 2, eprint 2016/510)
//
// c_0 + c_1 s + c_2 s^2 in Eval format, inverse NTT, then per coefficient: scale by gamma*t, exact base conversion to
// {t, gamma} (the sums are reduced modulo the small moduli, so any summation order gives the reference's residues),
// the gamma-centred correction, and the final multiplication by gamma^-1 * scalingFactor mod t.
#include "capi_internal.hpp"
#include "hostmath.hpp"
#include "modarith.cuh"

using namespace hecuda;
using namespace hecuda::api;

namespace {

// gamma = T.rnsCorrectionFactor comes from the context: 2^62 - 40797 (UInt64) or 2^30 - 20405 (UInt32), Scalar.swift:503-520

struct DecryptConsts {
    int l;
    u64 t;
    u64 gamma;                 // T.rnsCorrectionFactor
    u64 q[kMaxL];
    u64 gamma_t[kMaxL];        // gamma * t mod q_i                       (prodGammaTModQ, RnsTool.swift:145-146)
    u64 inv_punctured[kMaxL];  // (q / q_i)^-1 mod q_i                    (RnsBaseConverter)
    u64 punctured_t[kMaxL];    // q / q_i mod t
    u64 punctured_g[kMaxL];    // q / q_i mod gamma
    u64 neg_inv_q_t, neg_inv_q_g;  // -q^-1 mod t, mod gamma             (negInverseQModTGamma, :154-157)
    u64 inv_gamma_scaled;      // gamma^-1 * scalingFactor mod t          (:147-150, :298-299)
};

__device__ __forceinline__ u64 mulmod_dev(u64 a, u64 b, u64 m) { return (u64)(((u128)a * b) % m); }

// d = c_0 + sum_k c_k s^k (Eval), one thread per (item, row, coefficient)
__global__ void __launch_bounds__(256) dot_secret_kernel(const u64 *__restrict__ ct, const u64 *__restrict__ sk,
                                                        u64 *__restrict__ out, const __grid_constant__ DecryptConsts c,
                                                        int n, int polys) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const int row = blockIdx.y;
    const int64_t item = blockIdx.z;
    const u64 q = c.q[row];
    const u64 s = sk[(int64_t)row * n + e];
    const u64 *src = ct + (item * polys * c.l + row) * (int64_t)n + e;
    u64 acc = src[0], power = s;
    for (int k = 1; k < polys; ++k) {
        const u64 term = mulmod_dev(src[(int64_t)k * c.l * n], power, q);
        acc = acc + term >= q ? acc + term - q : acc + term;
        power = mulmod_dev(power, s, q);
    }
    out[(item * c.l + row) * (int64_t)n + e] = acc;
}

__global__ void __launch_bounds__(256) scale_and_round_kernel(const u64 *__restrict__ dot, u64 *__restrict__ out,
                                                             const __grid_constant__ DecryptConsts c, int n) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const int64_t item = blockIdx.y;
    u64 mod_t = 0, mod_g = 0;
    for (int i = 0; i < c.l; ++i) {
        const u64 x = mulmod_dev(dot[(item * c.l + i) * (int64_t)n + e], c.gamma_t[i], c.q[i]);
        const u64 y = mulmod_dev(x, c.inv_punctured[i], c.q[i]);
        mod_t = (mod_t + mulmod_dev(y % c.t, c.punctured_t[i], c.t)) % c.t;
        mod_g = (u64)(((u128)mod_g + mulmod_dev(y % c.gamma, c.punctured_g[i], c.gamma)) % c.gamma);
    }
    mod_t = mulmod_dev(mod_t, c.neg_inv_q_t, c.t);
    mod_g = mulmod_dev(mod_g, c.neg_inv_q_g, c.gamma);
    const u64 s_greater = (c.t - (c.gamma - mod_g) % c.t) % c.t;
    const u64 s_less = mod_g % c.t;
    const u64 s = mod_g > c.gamma / 2 ? s_greater : s_less;
    const u64 m = mod_t >= s ? mod_t - s : mod_t + c.t - s;
    out[item * (int64_t)n + e] = mulmod_dev(m, c.inv_gamma_scaled, c.t);
}

DecryptConsts make_consts(const Context &ctx, int l, u64 scaling_factor) {
    DecryptConsts c;
    c.l = l;
    c.t = ctx.t;
    const u64 kGamma = c.gamma = ctx.gamma;
    u64 q[kMaxL];
    for (int i = 0; i < l; ++i) q[i] = c.q[i] = ctx.slots[ctx.slot_q(i)].dev.p;
    for (int i = 0; i < l; ++i) {
        c.gamma_t[i] = host::mulmod(kGamma % q[i], ctx.t % q[i], q[i]);
        c.inv_punctured[i] = host::invmod(host::punctured_mod(q, l, i, q[i]), q[i]);
        c.punctured_t[i] = host::punctured_mod(q, l, i, ctx.t);
        c.punctured_g[i] = host::punctured_mod(q, l, i, kGamma);
    }
    const u64 q_t = host::prod_mod(q, l, ctx.t), q_g = host::prod_mod(q, l, kGamma);
    c.neg_inv_q_t = (ctx.t - host::invmod(q_t, ctx.t)) % ctx.t;
    c.neg_inv_q_g = (kGamma - host::invmod(q_g, kGamma)) % kGamma;
    c.inv_gamma_scaled = host::mulmod(host::invmod(kGamma % ctx.t, ctx.t), scaling_factor % ctx.t, ctx.t);
    return c;
}

cudaError_t decrypt_chunk(const Context &c, const DecryptConsts &dc, u64 *scratch, const u64 *sk, const u64 *ct, int polys,
                          u64 *out, int64_t items, cudaStream_t s) {
    const int l = dc.l;
    const int64_t n = c.n;
    u64 *ev = scratch, *dot = scratch + (size_t)items * polys * l * n;
    const NttRowMap map = c.map_q(l);
    cudaError_t e;
    if ((e = launch_ntt_forward(c, map, ct, ev, items * polys * l, s)) != cudaSuccess) return e;
    const int threads = n >= 256 ? 256 : (n < 32 ? 32 : (int)n);
    const unsigned gx = (unsigned)((n + threads - 1) / threads);
    for (int64_t done = 0; done < items;) {
        const int64_t part = std::min<int64_t>(items - done, 65535);
        ++g_kernel_launches;
        dot_secret_kernel<<<dim3(gx, (unsigned)l, (unsigned)part), threads, 0, s>>>(ev + done * polys * l * n, sk,
                                                                                   dot + done * l * n, dc, (int)n, polys);
        done += part;
    }
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    if ((e = launch_ntt_inverse(c, map, dot, dot, items * l, kScalePlain, s)) != cudaSuccess) return e;
    for (int64_t done = 0; done < items;) {
        const int64_t part = std::min<int64_t>(items - done, 65535);
        ++g_kernel_launches;
        scale_and_round_kernel<<<dim3(gx, (unsigned)part), threads, 0, s>>>(dot + done * l * n, out + done * n, dc, (int)n);
        done += part;
    }
    return cudaGetLastError();
}

}  // namespace

extern "C" {

int32_t hecuda_bfv_decrypt(const hecuda_context *h, const uint64_t *secret_key, const uint64_t *ciphertexts, int32_t polys,
                           int32_t l, uint64_t scaling_factor, uint64_t *plaintexts, int64_t batch) {
    int32_t rc = check_ctx(h);
    if (rc) return rc;
    const Context &c = *h->ctx;
    if (!secret_key) return fail(HECUDA_ERR_MISSING_KEY, "null secret key");
    if (polys < 2 || polys > 3) return fail(HECUDA_ERR_INVALID_ARGUMENT, "invalidCiphertext: poly_count must be 2 or 3");
    if (l < 1 || l > c.L) return fail(HECUDA_ERR_INVALID_ARGUMENT, "invalidCiphertext: moduli_count out of range");
    if (batch < 0 || (batch && (!ciphertexts || !plaintexts))) return fail(HECUDA_ERR_INVALID_ARGUMENT, "invalidCiphertext: null buffer");
    if (c.t >= c.gamma) return fail(HECUDA_ERR_UNSUPPORTED, "plaintext modulus too large");
    if (batch == 0) return HECUDA_OK;
    const DecryptConsts dc = make_consts(c, l, scaling_factor);
    // SecretKey.poly has K = L + 1 rows (Eval); rows 0..l-1 are the ones a level-l ciphertext uses
    u64 *d_sk = nullptr;
    CK(cudaMalloc(&d_sk, (size_t)l * c.n * sizeof(u64)));
    cudaError_t e = upload(d_sk, secret_key, (size_t)l * c.n * sizeof(u64));
    if (e != cudaSuccess) {
        cudaFree(d_sk);
        return cuda_fail(e, "secret key upload");
    }
    const size_t in_words = (size_t)polys * l * c.n;
    const int64_t chunk = std::max<int64_t>(1, (int64_t)((size_t)64 * 1024 * 1024 / in_words));
    WsGuard g(h);
    if (!g.w) {
        cudaFree(d_sk);
        return fail(HECUDA_ERR_CUDA, "could not create a CUDA stream / workspace");
    }
    cudaStream_t s = g.w->stream;
    u64 *d_in = nullptr, *d_scratch = nullptr, *d_out = nullptr;
    const int64_t cap = std::min<int64_t>(chunk, batch);
    e = cudaMallocAsync((void **)&d_in, in_words * cap * sizeof(u64), s);
    if (e == cudaSuccess) e = cudaMallocAsync((void **)&d_scratch, (in_words + (size_t)l * c.n) * cap * sizeof(u64), s);
    if (e == cudaSuccess) e = cudaMallocAsync((void **)&d_out, (size_t)c.n * cap * sizeof(u64), s);
    for (int64_t done = 0; e == cudaSuccess && done < batch; done += cap) {
        const int64_t items = std::min<int64_t>(cap, batch - done);
        e = cudaMemcpyAsync(d_in, ciphertexts + in_words * done, in_words * items * sizeof(u64), cudaMemcpyHostToDevice, s);
        if (e == cudaSuccess) e = decrypt_chunk(c, dc, d_scratch, d_sk, d_in, polys, d_out, items, s);
        if (e == cudaSuccess)
            e = cudaMemcpyAsync(plaintexts + (size_t)c.n * done, d_out, (size_t)c.n * items * sizeof(u64), cudaMemcpyDeviceToHost, s);
    }
    if (d_in) cudaFreeAsync(d_in, s);
    if (d_scratch) cudaFreeAsync(d_scratch, s);
    if (d_out) cudaFreeAsync(d_out, s);
    cudaError_t e2 = cudaStreamSynchronize(s);
    fill(d_sk, 0, (size_t)l * c.n * sizeof(u64));  // zeroize the key copy (the reference zeroizes SecretKey storage)
    cudaFree(d_sk);
    if (e == cudaSuccess) e = e2;
    return e == cudaSuccess ? HECUDA_OK : cuda_fail(e, "decrypt");
}

}  // extern "C"
