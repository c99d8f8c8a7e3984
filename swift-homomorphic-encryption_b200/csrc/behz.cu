// behz.cu -- the coefficient-wise BEHZ steps of BFV ct x ct multiply (eprint 2016/510).
//
//   lift   = _RnsTool.liftQToQBsk            RnsTool.swift:324-368  (+ RnsBaseConverter.swift:97-143)
//   tensor = Bfv.multiplyWithoutScaling      Bfv+Multiply.swift:80-82
//   floor  = _RnsTool.floorQBskToQ           RnsTool.swift:378-456
//
// Layout/launch: a thread owns two adjacent coefficient columns (16-byte loads/stores, coalesced along the
// coefficient axis; rows are strided by N); blockIdx.y/z select the polynomial, so there is no integer division.
// Arithmetic: each step evaluates the reference's chain of exact modular operations with pre-multiplied constants
// (context.hpp) as one 128-bit multiply-accumulate pass per output residue followed by ONE Montgomery reduction
// (the 2^64 factor lives in the constants); the stored residues are the same canonical values.  All kernels are
// instruction-issue bound (64-bit integer multiplies), not HBM bound -- see DESIGN.md.
#include "kernels.cuh"

namespace hecuda {

constexpr int kColsPerThread = 2;
constexpr int kThreads = 128;

__device__ __forceinline__ ulonglong2 ld2(const u64 *p) { return *reinterpret_cast<const ulonglong2 *>(p); }
__device__ __forceinline__ void st2(u64 *p, u64 a, u64 b) { *reinterpret_cast<ulonglong2 *>(p) = make_ulonglong2(a, b); }

// Bounds (checked for the actual moduli by Context::create): every 128-bit accumulator below stays < 2^127 and the
// Montgomery-reduced sums are < 2p (lift, f_j, out_i: one or two conditional subtractions) or < 4p (alpha).
template <int L>
__global__ void __launch_bounds__(kThreads) lift_kernel(const u64 *__restrict__ in, int polys_in, u64 *__restrict__ ext,
                                                       int ext_polys, int out_poly_offset,
                                                       const __grid_constant__ LiftConsts c, int n) {
    constexpr int R = 2 * L + 1;
    const int coeff = (blockIdx.x * kThreads + threadIdx.x) * kColsPerThread;
    if (coeff >= n) return;
    const int64_t poly = (int64_t)blockIdx.z * gridDim.y + blockIdx.y;  // index among items * polys_in
    const int64_t item = polys_in == 2 ? (poly >> 1) : poly / polys_in;  // (no 64-bit division on the hot path)
    const int pin = (int)(poly - item * polys_in);
    const u64 *src = in + poly * L * n + coeff;
    u64 *dst = ext + ((item * ext_polys + out_poly_offset + pin) * R) * n + coeff;
    u64 z[kColsPerThread][L];
    u32 acc_mt[kColsPerThread] = {0, 0};
#pragma unroll
    for (int i = 0; i < L; ++i) {
        const ulonglong2 x = ld2(src + (int64_t)i * n);
        st2(dst + (int64_t)i * n, x.x, x.y);
        // canonical: z is reinterpreted mod b_j and mod m~ below
        z[0][i] = shoup_mul(x.x, c.in_w[i], c.in_wp[i], c.q[i]);
        z[1][i] = shoup_mul(x.y, c.in_w[i], c.in_wp[i], c.q[i]);
        acc_mt[0] += (u32)z[0][i] * c.punct_mt[i];
        acc_mt[1] += (u32)z[1][i] * c.punct_mt[i];
    }
    u32 r[kColsPerThread];
    bool neg[kColsPerThread];
#pragma unroll
    for (int k = 0; k < kColsPerThread; ++k) {
        r[k] = acc_mt[k] * c.neg_inv_q_mt;   // [-x' Q^-1]_{m~}, RnsTool.swift:343-348
        neg[k] = r[k] >= 0x80000000u;        // centered representative r - m~ (:357-360)
    }
#pragma unroll
    for (int j = 0; j <= L; ++j) {
        u64 o[kColsPerThread];
#pragma unroll
        for (int k = 0; k < kColsPerThread; ++k) {
            const u64 rc = neg[k] ? (u64)r[k] + c.b[j] - 0x100000000ull : (u64)r[k];
            u128 acc = (u128)rc * c.qr[j];
#pragma unroll
            for (int i = 0; i < L; ++i) mac128(acc, z[k][i], c.mat[j][i]);
            o[k] = csub(mont_reduce(acc, c.b[j], c.b_ninv[j]), c.b[j]);
        }
        st2(dst + (int64_t)(L + j) * n, o[0], o[1]);
    }
}

// Tensor product in Montgomery form: out = a b 2^-64 mod p (canonical).  The missing 2^64 is restored by the
// kScaleTMont scaling of the inverse NTT that always follows (Bfv+Multiply.swift:80-82 then :40-41).
struct TensorConsts {
    int R;
    u64 p[kMaxRows], ninv[kMaxRows];
};

__global__ void __launch_bounds__(kThreads) tensor_kernel(const u64 *__restrict__ ext, u64 *__restrict__ ten,
                                                         const __grid_constant__ TensorConsts c, int n) {
    const int R = c.R;
    const int row = blockIdx.y;
    const int64_t item = blockIdx.z;
    const int coeff = (blockIdx.x * kThreads + threadIdx.x) * kColsPerThread;
    if (coeff >= n) return;
    const u64 p = c.p[row], ninv = c.ninv[row];
    const u64 *e = ext + (item * 4 * R + row) * n + coeff;
    const int64_t ps = (int64_t)R * n;
    const ulonglong2 a0 = ld2(e), a1 = ld2(e + ps), b0 = ld2(e + 2 * ps), b1 = ld2(e + 3 * ps);
    u64 *o = ten + (item * 3 * R + row) * n + coeff;
    st2(o, csub(mont_reduce((u128)a0.x * b0.x, p, ninv), p), csub(mont_reduce((u128)a0.y * b0.y, p, ninv), p));
    u128 m0 = (u128)a0.x * b1.x, m1 = (u128)a0.y * b1.y;
    mac128(m0, a1.x, b0.x);
    mac128(m1, a1.y, b0.y);
    st2(o + ps, csub(mont_reduce(m0, p, ninv), p), csub(mont_reduce(m1, p, ninv), p));
    st2(o + 2 * ps, csub(mont_reduce((u128)a1.x * b1.x, p, ninv), p), csub(mont_reduce((u128)a1.y * b1.y, p, ninv), p));
}

template <int L>
__global__ void __launch_bounds__(kThreads) floor_kernel(const u64 *__restrict__ in, u64 *__restrict__ out,
                                                        const __grid_constant__ FloorConsts c, int n) {
    constexpr int R = 2 * L + 1;
    const int coeff = (blockIdx.x * kThreads + threadIdx.x) * kColsPerThread;
    if (coeff >= n) return;
    const int64_t poly = (int64_t)blockIdx.z * gridDim.y + blockIdx.y;
    const u64 *src = in + poly * R * n + coeff;
    u64 *dst = out + poly * L * n + coeff;
    u64 y[kColsPerThread][L];
#pragma unroll
    for (int i = 0; i < L; ++i) {
        const ulonglong2 x = ld2(src + (int64_t)i * n);
        y[0][i] = shoup_mul(x.x, c.inq_w[i], c.inq_wp[i], c.q[i]);
        y[1][i] = shoup_mul(x.y, c.inq_w[i], c.inq_wp[i], c.q[i]);
    }
    // approximateFloor, RnsTool.swift:378-398: f_j = (x_bj - FBC(x_Q)_j) Q^-1 mod b_j   (kept lazy, < 2 b_j)
    u64 f[kColsPerThread][L + 1];
#pragma unroll
    for (int j = 0; j <= L; ++j) {
        const ulonglong2 xb = ld2(src + (int64_t)(L + j) * n);
        u128 acc0 = (u128)xb.x * c.fq[j], acc1 = (u128)xb.y * c.fq[j];
#pragma unroll
        for (int i = 0; i < L; ++i) {
            mac128(acc0, y[0][i], c.fmat[j][i]);
            mac128(acc1, y[1][i], c.fmat[j][i]);
        }
        f[0][j] = mont_reduce(acc0, c.b[j], c.b_ninv[j]);
        f[1][j] = mont_reduce(acc1, c.b[j], c.b_ninv[j]);
    }
    // convertApproximateBskToQ, RnsTool.swift:402-450
    const u64 msk = c.b[L];
    u64 outv[kColsPerThread][L];
#pragma unroll
    for (int k = 0; k < kColsPerThread; ++k) {
        u64 w[L];
        u128 acc = (u128)f[k][L] * c.a_msk;
#pragma unroll
        for (int i = 0; i < L; ++i) {
            w[i] = shoup_mul(f[k][i], c.inb_w[i], c.inb_wp[i], c.b[i]);  // canonical: reinterpreted mod m_sk and q_i
            mac128(acc, w[i], c.amat[i]);
        }
        u64 alpha = mont_reduce(acc, msk, c.b_ninv[L]);
        alpha = csub(csub(csub(alpha, 4 * msk), 2 * msk), msk);
        const bool exceeds = alpha > (msk >> 1);
        const u64 alpha_c = exceeds ? msk - alpha : alpha;
#pragma unroll
        for (int i = 0; i < L; ++i) {
            u128 o = (u128)alpha_c * (exceeds ? c.b_mod_q[i] : c.neg_b_mod_q[i]);
#pragma unroll
            for (int kk = 0; kk < L; ++kk) mac128(o, w[kk], c.omat[i][kk]);
            outv[k][i] = csub(csub(mont_reduce(o, c.q[i], c.q_ninv[i]), 2 * c.q[i]), c.q[i]);
        }
    }
#pragma unroll
    for (int i = 0; i < L; ++i) st2(dst + (int64_t)i * n, outv[0][i], outv[1][i]);
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

// grid over (coefficient pairs, polys) with polys folded into y (<= 32768) and z
static inline dim3 poly_grid(int64_t n, int64_t polys) {
    const unsigned gx = (unsigned)((n / kColsPerThread + kThreads - 1) / kThreads);
    const int64_t gy = polys < 32768 ? polys : 32768;
    return dim3(gx ? gx : 1, (unsigned)gy, (unsigned)((polys + gy - 1) / gy));
}

cudaError_t launch_lift(const Context &ctx, const u64 *in, int polys_in, u64 *ext, int ext_polys, int out_poly_offset,
                        int64_t items, cudaStream_t stream) {
    int64_t polys = items * polys_in;
    if (polys == 0) return cudaSuccess;
    const int64_t pstride_in = (int64_t)ctx.L * ctx.n;
    // the z dimension must divide exactly: launch in slabs of y = 32768 polys, then the remainder
    while (polys > 0) {
        int64_t slab = polys >= 32768 ? (polys / 32768) * 32768 : polys;
        const dim3 grid = poly_grid(ctx.n, slab);
        ++g_kernel_launches;
        HE_DISPATCH_L(ctx.L, (lift_kernel<LL><<<grid, kThreads, 0, stream>>>(in, polys_in, ext, ext_polys, out_poly_offset,
                                                                          ctx.lift, (int)ctx.n)));
        // advance whole items only (32768 is even and polys_in is 1 or 2)
        in += slab * pstride_in;
        ext += (slab / polys_in) * (int64_t)ext_polys * (2 * ctx.L + 1) * ctx.n;
        polys -= slab;
    }
    return cudaGetLastError();
}

cudaError_t launch_tensor(const Context &ctx, const u64 *ext, u64 *ten, int64_t items, cudaStream_t stream) {
    if (items == 0) return cudaSuccess;
    TensorConsts tc;
    tc.R = 2 * ctx.L + 1;
    const NttRowMap map = ctx.map_qbsk();
    for (int r = 0; r < tc.R; ++r) {
        tc.p[r] = ctx.slots[map.slot[r]].dev.p;
        tc.ninv[r] = ctx.slots[map.slot[r]].dev.ninv;
    }
    const unsigned gx = (unsigned)((ctx.n / kColsPerThread + kThreads - 1) / kThreads);
    for (int64_t done = 0; done < items;) {  // gridDim.z <= 65535
        const int64_t chunk = (items - done) > 65535 ? 65535 : (items - done);
        dim3 grid(gx ? gx : 1, (unsigned)tc.R, (unsigned)chunk);
        ++g_kernel_launches;
        tensor_kernel<<<grid, kThreads, 0, stream>>>(ext + done * 4 * tc.R * ctx.n, ten + done * 3 * tc.R * ctx.n, tc,
                                                     (int)ctx.n);
        done += chunk;
    }
    return cudaGetLastError();
}

cudaError_t launch_floor(const Context &ctx, const u64 *in, u64 *out, int64_t polys, cudaStream_t stream) {
    if (polys == 0) return cudaSuccess;
    const int R = 2 * ctx.L + 1;
    while (polys > 0) {
        int64_t slab = polys >= 32768 ? (polys / 32768) * 32768 : polys;
        const dim3 grid = poly_grid(ctx.n, slab);
        ++g_kernel_launches;
        HE_DISPATCH_L(ctx.L, (floor_kernel<LL><<<grid, kThreads, 0, stream>>>(in, out, ctx.floor, (int)ctx.n)));
        in += slab * (int64_t)R * ctx.n;
        out += slab * (int64_t)ctx.L * ctx.n;
        polys -= slab;
    }
    return cudaGetLastError();
}

}  // namespace hecuda
