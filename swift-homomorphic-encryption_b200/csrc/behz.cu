// behz.cu -- the coefficient-wise BEHZ steps of BFV ct x ct multiply (eprint 2016/510), one thread per coefficient
// column, rows strided by N so every load/store is coalesced along the coefficient axis.
//
//   lift   = _RnsTool.liftQToQBsk            RnsTool.swift:324-368  (+ RnsBaseConverter.swift:97-143)
//   tensor = Bfv.multiplyWithoutScaling      Bfv+Multiply.swift:80-82
//   floor  = _RnsTool.floorQBskToQ           RnsTool.swift:378-456
//
// Each step evaluates the reference's chain of exact modular operations with pre-multiplied constants
// (context.hpp) as one 128-bit multiply-accumulate pass per output residue followed by one Barrett reduction;
// the stored residues are the same canonical values.
#include "kernels.cuh"

namespace hecuda {

// Bounds (checked for the actual moduli by Context::create): every 128-bit accumulator below stays < 2^127 and the
// Montgomery-reduced sums are < 2p (lift, f_j, out_i) or < 4p (alpha), so the conditional subtractions shown suffice.
template <int L>
__global__ void __launch_bounds__(256) lift_kernel(const u64 *__restrict__ in, int polys_in, u64 *__restrict__ ext,
                                                  int ext_polys, int out_poly_offset, const __grid_constant__ LiftConsts c,
                                                  int64_t n, int64_t total) {
    constexpr int R = 2 * L + 1;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int64_t poly = idx / n, coeff = idx - poly * n;
    const int64_t item = poly / polys_in, pin = poly - item * polys_in;
    const u64 *src = in + poly * L * n + coeff;
    u64 *dst = ext + ((item * ext_polys + out_poly_offset + pin) * R) * n + coeff;
    u64 z[L];
    u32 acc_mt = 0;
#pragma unroll
    for (int i = 0; i < L; ++i) {
        const u64 x = src[(int64_t)i * n];
        dst[(int64_t)i * n] = x;
        z[i] = shoup_mul(x, c.in_w[i], c.in_wp[i], c.q[i]);  // canonical: reinterpreted mod b_j and mod m~ below
        acc_mt += (u32)z[i] * c.punct_mt[i];
    }
    const u32 r = acc_mt * c.neg_inv_q_mt;        // [-x' Q^-1]_{m~}, RnsTool.swift:343-348
    const bool neg = r >= 0x80000000u;            // centered representative r - m~ (:357-360)
#pragma unroll
    for (int j = 0; j <= L; ++j) {
        const u64 rc = neg ? (u64)r + c.b[j] - 0x100000000ull : (u64)r;
        u128 acc = (u128)rc * c.qr[j];
#pragma unroll
        for (int i = 0; i < L; ++i) mac128(acc, z[i], c.mat[j][i]);
        dst[(int64_t)(L + j) * n] = csub(mont_reduce(acc, c.b[j], c.b_ninv[j]), c.b[j]);
    }
}

// Tensor product in Montgomery form: out = a b 2^-64 mod p (canonical).  The missing 2^64 is restored by the
// kScaleTMont scaling of the inverse NTT that always follows (Bfv+Multiply.swift:80-82 then :40-41).
__global__ void __launch_bounds__(256) tensor_kernel(const u64 *__restrict__ ext, u64 *__restrict__ ten,
                                                    const ModSlot *__restrict__ slots, NttRowMap map, int64_t n) {
    const int R = map.rows_per_poly;
    const int row = blockIdx.y;
    const int64_t item = blockIdx.z;
    const int64_t coeff = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (coeff >= n) return;
    const ModSlot &S = slots[map.slot[row]];
    const u64 p = S.p, ninv = S.ninv;
    const u64 *e = ext + (item * 4 * R + row) * n + coeff;
    const int64_t ps = (int64_t)R * n;
    const u64 a0 = e[0], a1 = e[ps], b0 = e[2 * ps], b1 = e[3 * ps];
    u64 *o = ten + (item * 3 * R + row) * n + coeff;
    o[0] = csub(mont_reduce((u128)a0 * b0, p, ninv), p);
    u128 mid = (u128)a0 * b1;
    mac128(mid, a1, b0);
    o[ps] = csub(mont_reduce(mid, p, ninv), p);
    o[2 * ps] = csub(mont_reduce((u128)a1 * b1, p, ninv), p);
}

template <int L>
__global__ void __launch_bounds__(256) floor_kernel(const u64 *__restrict__ in, u64 *__restrict__ out,
                                                   const __grid_constant__ FloorConsts c, int64_t n, int64_t total) {
    constexpr int R = 2 * L + 1;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int64_t poly = idx / n, coeff = idx - poly * n;
    const u64 *src = in + poly * R * n + coeff;
    u64 *dst = out + poly * L * n + coeff;
    u64 y[L];
#pragma unroll
    for (int i = 0; i < L; ++i) y[i] = shoup_mul(src[(int64_t)i * n], c.inq_w[i], c.inq_wp[i], c.q[i]);
    // approximateFloor, RnsTool.swift:378-398: f_j = (x_bj - FBC(x_Q)_j) Q^-1 mod b_j   (kept lazy, < 2 b_j)
    u64 f[L + 1];
#pragma unroll
    for (int j = 0; j <= L; ++j) {
        u128 acc = (u128)src[(int64_t)(L + j) * n] * c.fq[j];
#pragma unroll
        for (int i = 0; i < L; ++i) mac128(acc, y[i], c.fmat[j][i]);
        f[j] = mont_reduce(acc, c.b[j], c.b_ninv[j]);
    }
    // convertApproximateBskToQ, RnsTool.swift:402-450
    const u64 msk = c.b[L];
    u64 w[L];
    u128 acc = (u128)f[L] * c.a_msk;
#pragma unroll
    for (int k = 0; k < L; ++k) {
        w[k] = shoup_mul(f[k], c.inb_w[k], c.inb_wp[k], c.b[k]);  // canonical: reinterpreted mod m_sk and mod q_i
        mac128(acc, w[k], c.amat[k]);
    }
    u64 alpha = mont_reduce(acc, msk, c.b_ninv[L]);
    alpha = csub(csub(csub(alpha, 4 * msk), 2 * msk), msk);
    const bool exceeds = alpha > (msk >> 1);
    const u64 alpha_c = exceeds ? msk - alpha : alpha;
#pragma unroll
    for (int i = 0; i < L; ++i) {
        u128 o = (u128)alpha_c * (exceeds ? c.b_mod_q[i] : c.neg_b_mod_q[i]);
#pragma unroll
        for (int k = 0; k < L; ++k) mac128(o, w[k], c.omat[i][k]);
        dst[(int64_t)i * n] = csub(csub(mont_reduce(o, c.q[i], c.q_ninv[i]), 2 * c.q[i]), c.q[i]);
    }
}

#define HE_DISPATCH_L(L_, CALL)                                                                                       \
    switch (L_) {                                                                                                     \
        case 1: { constexpr int LL = 1; CALL; } break;                                                                \
        case 2: { constexpr int LL = 2; CALL; } break;                                                                \
        case 3: { constexpr int LL = 3; CALL; } break;                                                                \
        case 4: { constexpr int LL = 4; CALL; } break;                                                                \
        case 5: { constexpr int LL = 5; CALL; } break;                                                                \
        case 6: { constexpr int LL = 6; CALL; } break;                                                                \
        case 7: { constexpr int LL = 7; CALL; } break;                                                                \
        case 8: { constexpr int LL = 8; CALL; } break;                                                                \
        case 9: { constexpr int LL = 9; CALL; } break;                                                                \
        case 10: { constexpr int LL = 10; CALL; } break;                                                              \
        case 11: { constexpr int LL = 11; CALL; } break;                                                              \
        case 12: { constexpr int LL = 12; CALL; } break;                                                              \
        case 13: { constexpr int LL = 13; CALL; } break;                                                              \
        case 14: { constexpr int LL = 14; CALL; } break;                                                              \
        case 15: { constexpr int LL = 15; CALL; } break;                                                              \
        case 16: { constexpr int LL = 16; CALL; } break;                                                              \
        default: return cudaErrorInvalidValue;                                                                        \
    }

cudaError_t launch_lift(const Context &ctx, const u64 *in, int polys_in, u64 *ext, int ext_polys, int out_poly_offset,
                        int64_t items, cudaStream_t stream) {
    const int64_t total = items * polys_in * ctx.n;
    if (total == 0) return cudaSuccess;
    const unsigned blocks = (unsigned)((total + 255) / 256);
    g_kernel_launches++;
    HE_DISPATCH_L(ctx.L, (lift_kernel<LL><<<blocks, 256, 0, stream>>>(in, polys_in, ext, ext_polys, out_poly_offset,
                                                                     ctx.lift, ctx.n, total)));
    return cudaGetLastError();
}

cudaError_t launch_tensor(const Context &ctx, const u64 *ext, u64 *ten, int64_t items, cudaStream_t stream) {
    if (items == 0) return cudaSuccess;
    const NttRowMap map = ctx.map_qbsk();
    const int threads = ctx.n >= 256 ? 256 : (int)ctx.n < 32 ? 32 : (int)ctx.n;
    for (int64_t done = 0; done < items;) {  // gridDim.z <= 65535
        const int64_t chunk = (items - done) > 65535 ? 65535 : (items - done);
        dim3 grid((unsigned)((ctx.n + threads - 1) / threads), (unsigned)map.rows_per_poly, (unsigned)chunk);
        ++g_kernel_launches;
        tensor_kernel<<<grid, threads, 0, stream>>>(ext + done * 4 * map.rows_per_poly * ctx.n,
                                                    ten + done * 3 * map.rows_per_poly * ctx.n, ctx.d_slots, map, ctx.n);
        done += chunk;
    }
    return cudaGetLastError();
}

cudaError_t launch_floor(const Context &ctx, const u64 *in, u64 *out, int64_t polys, cudaStream_t stream) {
    const int64_t total = polys * ctx.n;
    if (total == 0) return cudaSuccess;
    const unsigned blocks = (unsigned)((total + 255) / 256);
    g_kernel_launches++;
    HE_DISPATCH_L(ctx.L, (floor_kernel<LL><<<blocks, 256, 0, stream>>>(in, out, ctx.floor, ctx.n, total)));
    return cudaGetLastError();
}

}  // namespace hecuda
