// ntt_fast.cu -- sm_100a kernels of the register-tiled NTT (see ntt_fast.cuh for the pass structure).
// One CTA per row, N/16 threads, row staged in padded shared memory between passes, 16-byte coalesced global
// accesses on the contiguous side.  Rows are dispatched per modulus class (NARROW / WIDE) so each launch runs
// one specialised instruction stream.
#include "kernels.cuh"
#include "ntt_fast.cuh"

namespace hecuda {
using namespace fast;

constexpr int kMaxRowList = (kMaxL + 1) * kMaxL;  // key-switch digit rows: (l + 1) * l
struct RowList {          // rows (within a polynomial) that one launch handles
    int rows_per_poly;
    int count;
    unsigned short row[kMaxRowList];
    unsigned char slot[kMaxRowList];
    // input side (forward only): source row and whether it must be re-reduced into this row's modulus
    unsigned char src_row[kMaxRowList];
    unsigned char reduce[kMaxRowList];
    long long src_poly_stride;  // words between consecutive input polynomials
};

template <int LOGN, bool NARROW>
__global__ void __launch_bounds__((1 << LOGN) / 16, (1024 / ((1 << LOGN) / 16))) ntt_fwd_fast_kernel(const u64 *__restrict__ in, u64 *__restrict__ out,
                                                                       const ModSlot *__restrict__ slots,
                                                                       const __grid_constant__ RowList rl) {
    extern __shared__ u64 sm[];
    constexpr int P = plan_passes(LOGN);
    const int tau = threadIdx.x;
    const int64_t poly = blockIdx.x / rl.count;
    const int which = blockIdx.x - poly * rl.count;
    const int64_t row = poly * rl.rows_per_poly + rl.row[which];
    const ModSlot &S = slots[rl.slot[which]];
    RowMod m;
    m.p = S.p;
    m.two_p = 2 * S.p;
    m.mu1 = S.mu1;
    m.np = 0 - S.p;
    m.four_p = 4 * S.p;
    m.red_shift = S.red_shift;
    m.red_recip = S.red_recip;
    m.tw = S.tw;
    const u64 *src = in + poly * rl.src_poly_stride + ((int64_t)rl.src_row[which] << LOGN);
    u64 *dst = out + (row << LOGN);
    u64 x[16];
    {
        constexpr int C = fwd_c(LOGN, 0), LB = fwd_lb(LOGN, 0);
        load_global<LOGN, LB, C>(x, src, tau);
        if (rl.reduce[which]) {  // gathered key-switch digit whose source modulus is too large for the lazy range
#pragma unroll
            for (int r = 0; r < 16; ++r) x[r] = barrett64(x[r], m.p, m.mu1);
        }
        fwd_pass<LOGN, LB, C, NARROW>(x, tau, m);
        store_smem<LOGN, LB, C>(x, sm, tau);
    }
    __syncthreads();
    {
        constexpr int C = fwd_c(LOGN, 1), LB = fwd_lb(LOGN, 1);
        load_smem<LOGN, LB, C>(x, sm, tau);
        fwd_pass<LOGN, LB, C, NARROW>(x, tau, m);
        store_smem<LOGN, LB, C>(x, sm, tau);
    }
    __syncthreads();
    if (P == 4) {
        constexpr int C = fwd_c(LOGN, 2), LB = fwd_lb(LOGN, 2);
        load_smem<LOGN, LB, C>(x, sm, tau);
        fwd_pass<LOGN, LB, C, NARROW>(x, tau, m);
        store_smem<LOGN, LB, C>(x, sm, tau);
        __syncthreads();
    }
    {
        constexpr int C = fwd_c(LOGN, P - 1), LB = fwd_lb(LOGN, P - 1);
        load_smem<LOGN, LB, C>(x, sm, tau);
        fwd_pass<LOGN, LB, C, NARROW>(x, tau, m);
        fwd_finish<LOGN, NARROW>(x, m);
        store_global<LOGN, LB, C>(x, dst, tau);
    }
}

template <int LOGN, bool NARROW>
__global__ void __launch_bounds__((1 << LOGN) / 16, (1024 / ((1 << LOGN) / 16))) ntt_inv_fast_kernel(const u64 *__restrict__ in, u64 *__restrict__ out,
                                                                       const ModSlot *__restrict__ slots,
                                                                       const __grid_constant__ RowList rl, int scale_mode) {
    extern __shared__ u64 sm[];
    constexpr int P = plan_passes(LOGN);
    const int tau = threadIdx.x;
    const int64_t poly = blockIdx.x / rl.count;
    const int which = blockIdx.x - poly * rl.count;
    const int64_t row = poly * rl.rows_per_poly + rl.row[which];
    const ModSlot &S = slots[rl.slot[which]];
    RowMod m;
    m.p = S.p;
    m.two_p = 2 * S.p;
    m.mu1 = S.mu1;
    m.np = 0 - S.p;
    m.four_p = 4 * S.p;
    m.red_shift = S.red_shift;
    m.red_recip = S.red_recip;
    m.tw = S.itw;
    m.c0 = S.inv_scale[scale_mode].c0;
    m.c0p = S.inv_scale[scale_mode].c0p;
    m.c1 = S.inv_scale[scale_mode].c1;
    m.c1p = S.inv_scale[scale_mode].c1p;
    const u64 *src = in + (row << LOGN);
    u64 *dst = out + (row << LOGN);
    u64 x[16];
    {
        constexpr int C = inv_c(LOGN, 0), LB = inv_lb(LOGN, 0);
        load_global<LOGN, LB, C>(x, src, tau);
        inv_pass<LOGN, LB, C, NARROW, inv_bound_in(LOGN, 0)>(x, tau, m);
        store_smem<LOGN, LB, C>(x, sm, tau);
    }
    __syncthreads();
    {
        constexpr int C = inv_c(LOGN, 1), LB = inv_lb(LOGN, 1);
        load_smem<LOGN, LB, C>(x, sm, tau);
        if (NARROW && inv_reduce_at(LOGN, 1)) inv_reduce<0>(x, m);
        inv_pass<LOGN, LB, C, NARROW, inv_bound_in(LOGN, 1)>(x, tau, m);
        store_smem<LOGN, LB, C>(x, sm, tau);
    }
    __syncthreads();
    if (P == 4) {
        constexpr int C = inv_c(LOGN, 2), LB = inv_lb(LOGN, 2);
        load_smem<LOGN, LB, C>(x, sm, tau);
        if (NARROW && inv_reduce_at(LOGN, 2)) inv_reduce<0>(x, m);
        inv_pass<LOGN, LB, C, NARROW, inv_bound_in(LOGN, 2)>(x, tau, m);
        store_smem<LOGN, LB, C>(x, sm, tau);
        __syncthreads();
    }
    {
        constexpr int C = inv_c(LOGN, P - 1), LB = inv_lb(LOGN, P - 1);
        load_smem<LOGN, LB, C>(x, sm, tau);
        if (NARROW && inv_reduce_at(LOGN, P - 1)) inv_reduce<0>(x, m);
        inv_pass<LOGN, LB, C, NARROW, inv_bound_in(LOGN, P - 1)>(x, tau, m);
        store_global<LOGN, LB, C>(x, dst, tau);
    }
}

// ---------------------------------------------------------------------------------------------- launch
static void build_row_lists(const Context &ctx, const NttRowMap &map, RowList &narrow, RowList &wide) {
    narrow.rows_per_poly = wide.rows_per_poly = map.rows_per_poly;
    narrow.count = wide.count = 0;
    narrow.src_poly_stride = wide.src_poly_stride =
        map.src_mod ? map.src_poly_stride : (long long)map.rows_per_poly * ctx.n;
    for (int r = 0; r < map.rows_per_poly; ++r) {
        const int slot = map.slot[r / map.group];
        const bool is_narrow = ctx.slots[slot].dev.bits <= kNarrowBits;
        RowList &l = is_narrow ? narrow : wide;
        l.row[l.count] = (unsigned short)r;
        l.slot[l.count] = (unsigned char)slot;
        l.src_row[l.count] = (unsigned char)(map.src_mod ? r % map.src_mod : r);
        l.reduce[l.count] = 0;
        if (map.src_mod) {
            // inputs are residues mod the source modulus: fine as they are while they stay inside the lazy input range
            // of the butterflies (< 2p NARROW, < 4p WIDE); otherwise re-reduce on load (Bfv+Keys.swift:168-172)
            const u64 p = ctx.slots[slot].dev.p, src_p = ctx.slots[map.src_slot[r % map.src_mod]].dev.p;
            l.reduce[l.count] = (src_p > p && (src_p - 1) / p >= (is_narrow ? 2u : 4u)) ? 1 : 0;
        }
        ++l.count;
    }
}

template <int LOGN, bool NARROW, bool INVERSE>
static cudaError_t launch_class(const Context &ctx, const RowList &rl, const u64 *in, u64 *out, int64_t polys,
                                int scale_mode, cudaStream_t stream) {
    if (rl.count == 0 || polys == 0) return cudaSuccess;
    constexpr int threads = (1 << LOGN) / 16;
    constexpr size_t smem = sizeof(u64) * smem_words(LOGN);
    const int64_t blocks = polys * rl.count;
    if (blocks > 0x7fffffffLL) return cudaErrorInvalidValue;
    cudaError_t e;
    ++g_kernel_launches;
    if (INVERSE) {
        auto k = ntt_inv_fast_kernel<LOGN, NARROW>;
        if (smem > 48 * 1024 && (e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
        k<<<(unsigned)blocks, threads, smem, stream>>>(in, out, ctx.d_slots, rl, scale_mode);
    } else {
        auto k = ntt_fwd_fast_kernel<LOGN, NARROW>;
        if (smem > 48 * 1024 && (e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
        k<<<(unsigned)blocks, threads, smem, stream>>>(in, out, ctx.d_slots, rl);
    }
    return cudaGetLastError();
}

template <int LOGN, bool INVERSE>
static cudaError_t launch_logn(const Context &ctx, const NttRowMap &map, const u64 *in, u64 *out, int64_t rows,
                               int scale_mode, cudaStream_t stream) {
    if (rows % map.rows_per_poly) return cudaErrorInvalidValue;
    RowList narrow, wide;
    build_row_lists(ctx, map, narrow, wide);
    const int64_t polys = rows / map.rows_per_poly;
    cudaError_t e = launch_class<LOGN, true, INVERSE>(ctx, narrow, in, out, polys, scale_mode, stream);
    if (e != cudaSuccess) return e;
    return launch_class<LOGN, false, INVERSE>(ctx, wide, in, out, polys, scale_mode, stream);
}

bool ntt_fast_supported(const Context &ctx) { return ctx.logn >= kMinLogN && ctx.logn <= kMaxLogN; }

template <bool INVERSE>
static cudaError_t launch_fast(const Context &ctx, const NttRowMap &map, const u64 *in, u64 *out, int64_t rows,
                               int scale_mode, cudaStream_t stream) {
    switch (ctx.logn) {
        case 10: return launch_logn<10, INVERSE>(ctx, map, in, out, rows, scale_mode, stream);
        case 11: return launch_logn<11, INVERSE>(ctx, map, in, out, rows, scale_mode, stream);
        case 12: return launch_logn<12, INVERSE>(ctx, map, in, out, rows, scale_mode, stream);
        case 13: return launch_logn<13, INVERSE>(ctx, map, in, out, rows, scale_mode, stream);
        case 14: return launch_logn<14, INVERSE>(ctx, map, in, out, rows, scale_mode, stream);
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_ntt_forward_fast(const Context &ctx, const NttRowMap &map, const u64 *in, u64 *out, int64_t rows,
                                    cudaStream_t stream) {
    return launch_fast<false>(ctx, map, in, out, rows, 0, stream);
}
cudaError_t launch_ntt_inverse_fast(const Context &ctx, const NttRowMap &map, const u64 *in, u64 *out, int64_t rows,
                                    int scale_mode, cudaStream_t stream) {
    return launch_fast<true>(ctx, map, in, out, rows, scale_mode, stream);
}

}  // namespace hecuda
