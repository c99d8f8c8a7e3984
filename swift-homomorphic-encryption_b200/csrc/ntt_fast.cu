// ntt_fast.cu -- sm_100a kernels of the register-tiled NTT (see ntt_fast.cuh for the pass structure).
//
// Persistent CTAs: each CTA (T = N/16 threads, as many CTAs per SM as 64-register threads allow) walks over rows
// task = blockIdx.x, blockIdx.x + gridDim.x, ...  For every row
//   1. TMA in: one thread issues bulk-tensor copies (cp.async.bulk.tensor.2d, 256 lines = 32 KB each, 128-byte
//      swizzle) that land the row in shared memory and complete on an mbarrier; it also prefetches the CTA's next row
//      into L2 (cp.async.bulk.prefetch.tensor) and, when the row's modulus differs from the previous row's, bulk-copies
//      the first N/16 twiddles (all the LB > 0 passes need) into the CTA's shared-memory twiddle cache;
//   2. 3-4 register passes over the row in shared memory (ntt_fast.cuh), one __syncthreads between passes;
//   3. TMA out: bulk-tensor copies shared -> global of the finished row (same swizzle, undone by the copy engine);
//      the next row's TMA-in is issued by the same thread once the copy engine has read the buffer.
// NARROW / MID / WIDE rows (different lazy-reduction schedules) are mixed in one launch: the row's class selects the
// instruction stream, rows are ordered class-major so all CTAs walk the classes in step, and there is one tail per
// NTT call instead of one per class.
#include <cuda.h>

#include <mutex>

#include <algorithm>
#include <cstdlib>
#include "kernels.cuh"
#include "ntt_fast.cuh"

namespace hecuda {
using namespace fast;

constexpr int kMaxRowList = 2 * (kMaxL + 1) * kMaxL;  // key-switch digit rows: (l + 1) * l, twice that as half rows of N = 2^15
struct RowList {          // rows (within a polynomial) that one launch handles
    int rows_per_poly;
    int count;
    unsigned short row[kMaxRowList];
    unsigned short slot[kMaxRowList];  // (virtual slots of N = 2^15 exceed a byte at 32 moduli)
    // 3 bits class | bit 3: re-reduce the gathered input into this row's modulus (forward only) | bit 4: half of a
    // 2^15 row (inverse only: no N^-1 scaling)
    unsigned char flags[kMaxRowList];
    unsigned short src_row[kMaxRowList];  // input side (forward only): source row
    long long src_poly_stride;           // words between consecutive input polynomials
};

// ------------------------------------------------------------------------------------------------ TMA / mbarrier
__device__ __forceinline__ u32 smem_u32(const void *p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(u64 *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(u64 *bar, u32 bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(u64 *bar, u32 parity) {
    asm volatile(
        "{\n\t.reg .pred done;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 done, [%0], %1;\n\t"
        "@!done bra WAIT_%=;\n\t}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// 1-D bulk copy global -> shared (twiddle cache), completes `bytes` on the mbarrier
__device__ __forceinline__ void tma_load_1d(void *smem_dst, const void *gmem_src, u32 bytes, u64 *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// one box (16 words x kBoxLines lines) of the 2-D view {word in line, line} of a buffer, global -> shared
__device__ __forceinline__ void tma_load_box(void *smem_dst, const CUtensorMap *map, int line, u64 *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(map), "r"(0), "r"(line), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tma_store_box(const CUtensorMap *map, int line, const void *smem_src) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];" ::"l"(map), "r"(0), "r"(line),
                 "r"(smem_u32(smem_src))
                 : "memory");
}
__device__ __forceinline__ void tma_prefetch_box(const CUtensorMap *map, int line) {
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(map), "r"(0), "r"(line) : "memory");
}
__device__ __forceinline__ void tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------ row bodies
template <int LOGN, int CLS, int K>
__device__ __forceinline__ void fwd_row_pass(u64 (&x)[16], u64 *sm, int tau, const RowMod &m, bool reduce_in) {
    constexpr int P = plan_passes(LOGN), C = fwd_c(LOGN, K), LB = fwd_lb(LOGN, K);
    load_smem<LOGN, LB, C>(x, sm, tau);
    if (K == 0 && reduce_in) {  // gathered key-switch digit whose source modulus is too large for the lazy range
#pragma unroll
        for (int r = 0; r < 16; ++r) x[r] = barrett64(x[r], 0 - m.np, m.slot->mu1);
    }
    fwd_pass<LOGN, LB, C, CLS>(x, tau, m);
    if (K == P - 1) fwd_finish<CLS>(x, m);
    store_smem<LOGN, LB, C>(x, sm, tau);
}
template <int LOGN, int CLS>
__device__ __forceinline__ void fwd_row(u64 *sm, int tau, const RowMod &m, bool reduce_in) {
    constexpr int P = plan_passes(LOGN);
    u64 x[16];
    fwd_row_pass<LOGN, CLS, 0>(x, sm, tau, m, reduce_in);
    __syncthreads();
    fwd_row_pass<LOGN, CLS, 1>(x, sm, tau, m, false);
    __syncthreads();
    if (P == 4) {
        fwd_row_pass<LOGN, CLS, (P == 4 ? 2 : 1)>(x, sm, tau, m, false);
        __syncthreads();
    }
    fwd_row_pass<LOGN, CLS, P - 1>(x, sm, tau, m, false);
}

template <int LOGN, int CLS, int K>
__device__ __forceinline__ void inv_row_pass(u64 (&x)[16], u64 *sm, int tau, const RowMod &m) {
    constexpr int C = inv_c(LOGN, K), LB = inv_lb(LOGN, K);
    load_smem<LOGN, LB, C>(x, sm, tau);
    if (narrow_like(CLS) && K > 0 && inv_reduce_at(LOGN, K)) inv_reduce(x, m);
    inv_pass<LOGN, LB, C, CLS, inv_bound_in(LOGN, K)>(x, tau, m);
    if (K == plan_passes(LOGN) - 1 && m.partial) {  // half of a 2^15 row: hand canonical residues to the merge kernel
        if (CLS == kSmall) {
#pragma unroll
            for (int r = 0; r < 16; ++r) x[r] = csub32((u32)x[r], (u32)(0 - m.np));
        } else {
            reduce_small16(x, m);  // NARROW < 512 p, MID < 4p, WIDE < 2p
        }
    }
    store_smem<LOGN, LB, C>(x, sm, tau);
}
template <int LOGN, int CLS>
__device__ __forceinline__ void inv_row(u64 *sm, int tau, const RowMod &m) {
    constexpr int P = plan_passes(LOGN);
    u64 x[16];
    inv_row_pass<LOGN, CLS, 0>(x, sm, tau, m);
    __syncthreads();
    inv_row_pass<LOGN, CLS, 1>(x, sm, tau, m);
    __syncthreads();
    if (P == 4) {
        inv_row_pass<LOGN, CLS, (P == 4 ? 2 : 1)>(x, sm, tau, m);
        __syncthreads();
    }
    inv_row_pass<LOGN, CLS, P - 1>(x, sm, tau, m);
}

// Shared memory of a CTA: [row: N words][twiddle cache: N/16 entries][2 mbarriers].
template <int LOGN>
constexpr size_t ntt_smem_bytes() {
    return sizeof(u64) * ((size_t)1 << LOGN) + sizeof(ulonglong2) * ((size_t)1 << (LOGN - 4)) + 16;
}

template <int LOGN, bool INVERSE>
__global__ void __launch_bounds__((1 << LOGN) / 16, (1024 / ((1 << LOGN) / 16)))
    ntt_rows_kernel(const __grid_constant__ CUtensorMap map_in, const __grid_constant__ CUtensorMap map_out,
                    const ModSlot *__restrict__ slots, const __grid_constant__ RowList rl, const int polys,
                    const int scale_mode, const int debug_flags) {
    extern __shared__ __align__(1024) u64 sm[];  // row first: the 128-byte swizzle wants it 1024-byte aligned
    constexpr int kLines = (1 << LOGN) / kLineWords;
    constexpr int kBoxes = kLines > kBoxLines ? kLines / kBoxLines : 1;
    constexpr int kLinesPerBox = kLines / kBoxes;
    constexpr u32 kRowBytes = (u32)sizeof(u64) << LOGN, kTwBytes = (u32)sizeof(ulonglong2) << (LOGN - 4);
    ulonglong2 *tw_cache = reinterpret_cast<ulonglong2 *>(sm + (1 << LOGN));
    u64 *bar_row = reinterpret_cast<u64 *>(tw_cache + (1 << (LOGN - 4)));
    u64 *bar_tw = bar_row + 1;
    const int tau = threadIdx.x;
    const int tasks = polys * rl.count;  // < 2^31 (checked by the launcher)
    if (tau == 0) {
        if (smem_u32(sm) & 1023) __trap();  // dynamic shared memory starts at the window base when there is no static part
        mbar_init(bar_row, 1);
        mbar_init(bar_tw, 1);
    }
    __syncthreads();
    u32 phase_row = 0, phase_tw = 0;
    int cached_slot = -1;
    for (int task = blockIdx.x; task < tasks; task += gridDim.x) {
        const int which = task / polys;
        const int poly = task - which * polys;
        const int flags = rl.flags[which];
        const int slot = rl.slot[which];
        const ModSlot &S = slots[slot];
        // line index (16-word lines) of the row inside the buffer each tensor map describes
        const int64_t out_word = ((int64_t)poly * rl.rows_per_poly + rl.row[which]) << LOGN;
        const int64_t in_word = INVERSE ? out_word : (int64_t)poly * rl.src_poly_stride + ((int64_t)rl.src_row[which] << LOGN);
        const bool new_slot = slot != cached_slot;  // uniform over the CTA
        cached_slot = slot;
        if (tau == 0) {
            // the previous row's TMA-out has finished reading the buffer (wait_group.read below, same thread), and every
            // thread has passed the barrier that follows its last use of the twiddle cache
            if (new_slot) {
                mbar_arrive_expect_tx(bar_tw, kTwBytes);
                tma_load_1d(tw_cache, INVERSE ? S.itw : S.tw, kTwBytes, bar_tw);
            }
            mbar_arrive_expect_tx(bar_row, kRowBytes);
            const int line = (int)(in_word >> 4);
#pragma unroll
            for (int b = 0; b < kBoxes; ++b) tma_load_box(sm + b * kLinesPerBox * kLineWords, &map_in, line + b * kLinesPerBox, bar_row);
            const int next = task + gridDim.x;
            if (next < tasks && !(debug_flags & 1)) {
                const int nw = next / polys;
                const int np_ = next - nw * polys;
                const int64_t nword = INVERSE ? ((int64_t)np_ * rl.rows_per_poly + rl.row[nw]) << LOGN
                                              : (int64_t)np_ * rl.src_poly_stride + ((int64_t)rl.src_row[nw] << LOGN);
#pragma unroll
                for (int b = 0; b < kBoxes; ++b) tma_prefetch_box(&map_in, (int)(nword >> 4) + b * kLinesPerBox);
            }
        }
        const int cls = flags & 7;
        RowMod m;
        m.np = 0 - S.p;
        m.kp = (cls == kWide || cls == kSmall) ? 2 * S.p : 4 * S.p;  // (NARROW / NARROW-H / MID: 4p)
        m.tw = nullptr;
        m.tw_s = smem_u32(tw_cache);
        m.slot = &S;
        m.scale_mode = INVERSE ? scale_mode : -1;
        m.partial = INVERSE && (flags & 16) != 0;
        if (new_slot) {
            mbar_wait(bar_tw, phase_tw);
            phase_tw ^= 1;
        }
        mbar_wait(bar_row, phase_row);
        phase_row ^= 1;
        if (debug_flags & 4) {  // experiments: data movement only
        } else
#ifndef HE_EXPERIMENT_ONLY_CLASS
        if (INVERSE) {
            if (kNarrowHEnabled && cls == kNarrowH) inv_row<LOGN, kNarrowHEnabled ? kNarrowH : kNarrow>(sm, tau, m);
            else if (cls == kNarrow) inv_row<LOGN, kNarrow>(sm, tau, m);
            else if (cls == kSmall) inv_row<LOGN, kSmall>(sm, tau, m);
            else if (cls == kMid) inv_row<LOGN, kMid>(sm, tau, m);
            else inv_row<LOGN, kWide>(sm, tau, m);
        } else {
            const bool reduce_in = (flags & 8) != 0;
            if (kNarrowHEnabled && cls == kNarrowH) fwd_row<LOGN, kNarrowHEnabled ? kNarrowH : kNarrow>(sm, tau, m, reduce_in);
            else if (cls == kNarrow) fwd_row<LOGN, kNarrow>(sm, tau, m, reduce_in);
            else if (cls == kSmall) fwd_row<LOGN, kSmall>(sm, tau, m, reduce_in);
            else if (cls == kMid) fwd_row<LOGN, kMid>(sm, tau, m, reduce_in);
            else fwd_row<LOGN, kWide>(sm, tau, m, reduce_in);
        }
#else  // register-pressure experiments: one class only
        if (INVERSE) inv_row<LOGN, HE_EXPERIMENT_ONLY_CLASS>(sm, tau, m);
        else fwd_row<LOGN, HE_EXPERIMENT_ONLY_CLASS>(sm, tau, m, (flags & 8) != 0);
#endif
        // ---- TMA out: every thread makes its generic-proxy writes visible to the async proxy, then one thread copies
        fence_proxy_async_smem();
        __syncthreads();
        if (tau == 0 && !(debug_flags & 2)) {
            const int line = (int)(out_word >> 4);
#pragma unroll
            for (int b = 0; b < kBoxes; ++b) tma_store_box(&map_out, line + b * kLinesPerBox, sm + b * kLinesPerBox * kLineWords);
            tma_commit();
            tma_store_wait_read();  // the buffer may be refilled once the copy engine has read it
        }
    }
    if (tau == 0) tma_store_wait_all();
}

// ---------------------------------------------------------------------------------------------- launch
static void build_row_list(const Context &ctx, const NttRowMap &map, bool inverse, RowList &rl) {
    rl.rows_per_poly = map.rows_per_poly;
    rl.count = 0;
    rl.src_poly_stride = map.src_mod ? map.src_poly_stride : (long long)map.rows_per_poly * ctx.n;
    // class-major order (the cheapest rows last, so the tail of the launch is made of short rows)
    static const int order[5] = {kWide, kMid, kNarrow, kNarrowH, kSmall};
    for (int oi = 0; oi < 5; ++oi) {
        const int cls = order[oi];
        for (int r = 0; r < map.rows_per_poly; ++r) {
            const int slot = map.slot[r / map.group];
            if (class_of_modulus(ctx.slots[slot].dev.p, ctx.slots[slot].dev.bits) != cls) continue;
            const int i = rl.count++;
            rl.row[i] = (unsigned short)r;
            rl.slot[i] = (unsigned short)slot;
            rl.src_row[i] = (unsigned short)(map.src_mod && !inverse ? r % map.src_mod : r);
            int flags = cls;
            if (map.src_mod && !inverse) {
                // inputs are residues mod the source modulus: fine as they are while they stay inside the lazy input
                // range of the butterflies (< 2p NARROW, < 8p MID, < 4p WIDE); otherwise re-reduce on load
                // (Bfv+Keys.swift:168-172)
                const u64 p = ctx.slots[slot].dev.p, src_p = ctx.slots[map.src_slot[r % map.src_mod]].dev.p;
                const u64 room = narrow_like(cls) ? 2u : cls == kMid ? 8u : cls == kSmall ? 1u : 4u;
                if (src_p > p && (src_p - 1) / p >= room) flags |= 8;
            }
            rl.flags[i] = (unsigned char)flags;
        }
    }
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return (EncodeTiledFn)p;
    }();
    return fn;
}
// 2-D view of a u64 buffer for the row copies: dimension 0 = the 16 words of a 128-byte line, dimension 1 = lines
static bool make_line_map(CUtensorMap *map, const u64 *base, int logn) {
    EncodeTiledFn enc = encode_tiled();
    if (!enc || (reinterpret_cast<uintptr_t>(base) & 15)) return false;
    const int lines = (1 << logn) / kLineWords;
    const cuuint64_t dims[2] = {kLineWords, (cuuint64_t)1 << 31};  // extent only bounds the coordinates
    const cuuint64_t strides[1] = {kLineWords * sizeof(u64)};
    const cuuint32_t box[2] = {kLineWords, (cuuint32_t)(lines > kBoxLines ? kBoxLines : lines)};
    const cuuint32_t elem_strides[2] = {1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, const_cast<u64 *>(base), dims, strides, box, elem_strides,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int LOGN, bool INVERSE>
static cudaError_t launch_logn(const Context &ctx, const NttRowMap &map, const u64 *in, u64 *out, int64_t rows,
                               int scale_mode, cudaStream_t stream) {
    if (rows % map.rows_per_poly) return cudaErrorInvalidValue;
    if (rows == 0) return cudaSuccess;
    if (rows > 0x7fffffffLL) return cudaErrorInvalidValue;
    RowList rl;
    build_row_list(ctx, map, INVERSE, rl);
    if (rl.src_poly_stride % kLineWords) return cudaErrorInvalidValue;
    CUtensorMap map_in, map_out;
    if (!make_line_map(&map_in, in, LOGN) || !make_line_map(&map_out, out, LOGN)) return cudaErrorInvalidValue;
    const int polys = (int)(rows / map.rows_per_poly);
    constexpr int threads = (1 << LOGN) / 16;
    constexpr size_t smem = ntt_smem_bytes<LOGN>();
    auto k = ntt_rows_kernel<LOGN, INVERSE>;
    static int ctas_per_sm[64] = {0};  // per instantiation and device
    static std::mutex mu;
    int per_sm;
    {
        std::lock_guard<std::mutex> lock(mu);
        int &c = ctas_per_sm[ctx.device & 63];
        if (c == 0) {
            cudaError_t e;
            if ((e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
            int n = 0;
            if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k, threads, smem)) != cudaSuccess) return e;
            if (n < 1) return cudaErrorInvalidConfiguration;
            c = n;
        }
        per_sm = c;
    }
    const int64_t grid = std::min<int64_t>(rows, (int64_t)ctx.sm_count * per_sm);
    ++g_kernel_launches;
    static const int debug_flags = [] {  // experiments only: bit 0 = no L2 prefetch of the next row, bit 1 = no TMA-out, bit 2 = no compute
        const char *env = std::getenv("HECUDA_NTT_DEBUG");
        return env ? std::atoi(env) : 0;
    }();
    k<<<(unsigned)grid, threads, smem, stream>>>(map_in, map_out, ctx.d_slots, rl, polys, scale_mode, debug_flags);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------- N = 2^15
// A 2^15 row (256 KB) does not fit one CTA's shared memory.  Its first forward stage (pairs i, i + N/2, one twiddle) and
// last inverse stage are elementwise over the two halves; in between each half is an independent 2^14-point transform
// that uses the entries (2 + h) 2^s + g of the twiddle table, which context.cu lays out as a table of its own ("virtual
// slot").  So: forward = split kernel (global memory, also performs the key-switch gather) + the 2^14 kernel over twice
// as many rows; inverse = the 2^14 kernel with the `partial` flag + merge kernel (N^-1 scaling).
struct SplitRows {
    int rows_per_poly, count;
    unsigned short row[kMaxRowList];
    unsigned short src_row[kMaxRowList], slot[kMaxRowList];
    unsigned char reduce_in[kMaxRowList];
    long long src_poly_stride;
};
__global__ void __launch_bounds__(256) ntt_split_forward_kernel(const u64 *__restrict__ in, u64 *__restrict__ out,
                                                               const ModSlot *__restrict__ slots,
                                                               const __grid_constant__ SplitRows sr, int polys, int logn) {
    const int64_t n = (int64_t)1 << logn, half = n >> 1;
    const int which = blockIdx.y;
    const int64_t poly = blockIdx.z;
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 2;  // two adjacent coefficients per thread
    if (i >= half) return;
    const ModSlot &S = slots[sr.slot[which]];
    const u64 p = S.p;
    const ulonglong2 w = S.tw[1];
    const u64 *src = in + poly * sr.src_poly_stride + ((int64_t)sr.src_row[which] << logn) + i;
    u64 *dst = out + (((poly * sr.rows_per_poly) + sr.row[which]) << logn) + i;
    ulonglong2 x = *reinterpret_cast<const ulonglong2 *>(src), y = *reinterpret_cast<const ulonglong2 *>(src + half);
    if (sr.reduce_in[which]) {  // gathered key-switch digit: residues of another modulus (Bfv+Keys.swift:168-172)
        x.x = barrett64(x.x, p, S.mu1), x.y = barrett64(x.y, p, S.mu1);
        y.x = barrett64(y.x, p, S.mu1), y.y = barrett64(y.y, p, S.mu1);
    }
    const u64 v0 = shoup_mul(y.x, w.x, w.y, p), v1 = shoup_mul(y.y, w.x, w.y, p);
    *reinterpret_cast<ulonglong2 *>(dst) = make_ulonglong2(add_mod(x.x, v0, p), add_mod(x.y, v1, p));
    *reinterpret_cast<ulonglong2 *>(dst + half) = make_ulonglong2(sub_mod(x.x, v0, p), sub_mod(x.y, v1, p));
}
__global__ void __launch_bounds__(256) ntt_merge_inverse_kernel(u64 *__restrict__ data, const ModSlot *__restrict__ slots,
                                                               const __grid_constant__ SplitRows sr, int polys, int logn,
                                                               int scale_mode) {
    const int64_t n = (int64_t)1 << logn, half = n >> 1;
    const int which = blockIdx.y;
    const int64_t poly = blockIdx.z;
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 2;
    if (i >= half) return;
    const ModSlot &S = slots[sr.slot[which]];
    const u64 p = S.p;
    const ModSlot::InvScale sc = S.inv_scale[scale_mode];
    u64 *row = data + (((poly * sr.rows_per_poly) + sr.row[which]) << logn) + i;
    const ulonglong2 x = *reinterpret_cast<const ulonglong2 *>(row), y = *reinterpret_cast<const ulonglong2 *>(row + half);
    // (x + y) N^-1 and (x - y) N^-1 psi^-(N/2), each times the scaling of scale_mode (PolyRq+Ntt.swift:416-419)
    *reinterpret_cast<ulonglong2 *>(row) =
        make_ulonglong2(shoup_mul(x.x + y.x, sc.c0, sc.c0p, p), shoup_mul(x.y + y.y, sc.c0, sc.c0p, p));
    *reinterpret_cast<ulonglong2 *>(row + half) =
        make_ulonglong2(shoup_mul(x.x - y.x + p, sc.c1, sc.c1p, p), shoup_mul(x.y - y.y + p, sc.c1, sc.c1p, p));
}

template <bool INVERSE>
static cudaError_t launch_split(const Context &ctx, const NttRowMap &map, const u64 *in, u64 *out, int64_t rows, int scale_mode,
                                cudaStream_t stream) {
    constexpr int LOGN = kMaxLogN;  // the half transforms
    if (rows % map.rows_per_poly) return cudaErrorInvalidValue;
    if (rows == 0) return cudaSuccess;
    if (rows * 2 > 0x7fffffffLL || 2 * map.rows_per_poly > kMaxRowList) return cudaErrorInvalidValue;
    const int polys = (int)(rows / map.rows_per_poly);
    RowList full;
    build_row_list(ctx, map, INVERSE, full);
    SplitRows sr;
    sr.rows_per_poly = full.rows_per_poly;
    sr.count = full.count;
    sr.src_poly_stride = full.src_poly_stride;
    RowList rl;  // the half rows: row 2r + h of a polynomial with twice as many rows, on the virtual slot of (slot, h)
    rl.rows_per_poly = 2 * full.rows_per_poly;
    rl.count = 2 * full.count;
    rl.src_poly_stride = (long long)rl.rows_per_poly << LOGN;
    for (int i = 0; i < full.count; ++i) {
        sr.row[i] = full.row[i];
        sr.slot[i] = full.slot[i];
        sr.src_row[i] = full.src_row[i];
        // any gathered row is re-reduced here (cheap next to the two transforms), so the halves see canonical residues
        sr.reduce_in[i] = (unsigned char)((map.src_mod && !INVERSE) ? 1 : 0);
        for (int hh = 0; hh < 2; ++hh) {
            const int j = 2 * i + hh;
            rl.row[j] = (unsigned short)(2 * full.row[i] + hh);
            rl.src_row[j] = rl.row[j];  // the half transforms run in place: source row = the row itself
            rl.slot[j] = (unsigned short)(ctx.split_slot_base + 2 * full.slot[i] + hh);
            rl.flags[j] = (unsigned char)((full.flags[i] & 7) | (INVERSE ? 16 : 0));
        }
    }
    const dim3 egrid((unsigned)((ctx.n / 4 + 255) / 256), (unsigned)full.count, (unsigned)polys);
    if (!INVERSE) {
        ++g_kernel_launches;
        ntt_split_forward_kernel<<<egrid, 256, 0, stream>>>(in, out, ctx.d_slots, sr, polys, ctx.logn);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        in = out;
    }
    // the 2^14 kernel, in place over `out` (forward) / from `in` to `out` (inverse); rows addressed like inverse rows
    CUtensorMap map_in, map_out;
    if (!make_line_map(&map_in, in, LOGN) || !make_line_map(&map_out, out, LOGN)) return cudaErrorInvalidValue;
    {
        constexpr int threads = (1 << LOGN) / 16;
        constexpr size_t smem = ntt_smem_bytes<LOGN>();
        auto k = ntt_rows_kernel<LOGN, INVERSE>;
        cudaError_t e;
        if ((e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
        int per_sm = 0;
        if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, threads, smem)) != cudaSuccess) return e;
        if (per_sm < 1) return cudaErrorInvalidConfiguration;
        const int64_t grid = std::min<int64_t>(2 * rows, (int64_t)ctx.sm_count * per_sm);
        ++g_kernel_launches;
        k<<<(unsigned)grid, threads, smem, stream>>>(map_in, map_out, ctx.d_slots, rl, polys, scale_mode, 0);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
    }
    if (INVERSE) {
        ++g_kernel_launches;
        ntt_merge_inverse_kernel<<<egrid, 256, 0, stream>>>(out, ctx.d_slots, sr, polys, ctx.logn, scale_mode);
        return cudaGetLastError();
    }
    return cudaSuccess;
}

bool ntt_fast_supported(const Context &ctx) { return ctx.logn >= kMinLogN && ctx.logn <= kSplitLogN; }

template <bool INVERSE>
static cudaError_t launch_fast(const Context &ctx, const NttRowMap &map, const u64 *in, u64 *out, int64_t rows,
                               int scale_mode, cudaStream_t stream) {
    switch (ctx.logn) {
        case 10: return launch_logn<10, INVERSE>(ctx, map, in, out, rows, scale_mode, stream);
        case 11: return launch_logn<11, INVERSE>(ctx, map, in, out, rows, scale_mode, stream);
        case 12: return launch_logn<12, INVERSE>(ctx, map, in, out, rows, scale_mode, stream);
        case 13: return launch_logn<13, INVERSE>(ctx, map, in, out, rows, scale_mode, stream);
        case 14: return launch_logn<14, INVERSE>(ctx, map, in, out, rows, scale_mode, stream);
        case 15: return launch_split<INVERSE>(ctx, map, in, out, rows, scale_mode, stream);
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
