// ntt_fast.cuh -- register-tiled negacyclic NTT passes (shared by the sm_100a kernels in ntt_fast.cu and by the
// host-side SIMT emulation in tests/emu/ntt_emulate.cu, which replays the exact index maps on the CPU).
//
// One CTA transforms one row of N = 2^LOGN residues with T = N/16 threads; every thread keeps 16 residues in
// registers and runs up to 4 butterfly stages on them between two shared-memory exchanges.
//
// A pass is described by (LB, C): it covers C consecutive stages and works on "sub-blocks" of 2^C elements
//     e = (hi << (LB + C)) | (a << LB) | lo,      a = 0 .. 2^C - 1
// one sub-block per (hi, lo).  Thread tau owns the G = 16 >> C sub-blocks sb = tau + g*T (g < G), with
// hi = sb >> LB, lo = sb & (2^LB - 1); register index r = g * 2^C + a.
//   forward (Cooley-Tukey, reference _NttContext.forwardNtt, PolyRq+Ntt.swift:237-319): stages in order
//       s = S0 + j (S0 = LOGN - LB - C), pair distance 2^(C-1-j) in `a`, twiddle index 2^s + (hi << j) + (a >> (C-j))
//   inverse (Gentleman-Sande, PolyRq+Ntt.swift:379-483): stages with t = 2^(LB + j), pair distance 2^j in `a`,
//       m = 2^(LOGN-1-LB-j) groups, twiddle index m + (hi << (C-1-j)) + (a >> (j+1)); the stage with m = 1
//       multiplies by N^-1 and N^-1 psi^-(N/2) (optionally times t).
// The pass lists below make every shared-memory access conflict-free with the padding phys(e) = e + (e >> 4)
// and make the global side of the first/last pass fully coalesced (checked by tests/test_ntt_plan.py).
//
// Lazy ranges.  NARROW (p < 2^55): forward never reduces until the end (values < (2 + 2 LOGN) p; the bound analysis
// below allows up to 4p per twiddle product so a cheaper under-estimated Shoup quotient, shoup_lazy4, could be dropped
// in -- measured: no gain, the compiler's sequence for it issues as many instructions), inverse lets values double
// per stage and reduces once at the entry of the pass that could overflow; reductions use a small-quotient estimate
// (one 32-bit multiply) instead of a 64-bit Barrett because the FMA-heavy pipe (IMAD.WIDE) is the measured bottleneck.  WIDE (p < 2^62): Harvey's [0,4p) forward / [0,2p) inverse
// with exact quotients and one conditional subtraction per butterfly.
#pragma once
#include "context.hpp"
#include "modarith.cuh"

namespace hecuda {
namespace fast {

constexpr int kMinLogN = 10, kMaxLogN = 14;
constexpr int kNarrowBits = 55;  // moduli below 2^55 take the reduction-free butterflies (lazy values < 512 p < 2^64)

HE_HD constexpr int plan_passes(int logn) { return logn == 10 ? 3 : 4; }
// stage counts of the forward passes, in execution order; the inverse runs the mirrored list
HE_HD constexpr int plan_c(int logn, int k) {
    return logn == 10   ? (k == 0 ? 4 : k == 1 ? 4 : 2)
           : logn == 11 ? (k == 0 ? 1 : k == 1 ? 4 : k == 2 ? 4 : 2)
           : logn == 12 ? (k == 0 ? 4 : k == 1 ? 2 : k == 2 ? 4 : 2)
           : logn == 13 ? (k == 3 ? 1 : 4)
                        : (k == 3 ? 2 : 4);
}
HE_HD constexpr int fwd_c(int logn, int k) { return plan_c(logn, k); }
HE_HD constexpr int fwd_lb(int logn, int k) {
    int s0 = 0;
    for (int i = 0; i < k; ++i) s0 += plan_c(logn, i);
    return logn - s0 - plan_c(logn, k);
}
HE_HD constexpr int inv_c(int logn, int k) { return plan_c(logn, plan_passes(logn) - 1 - k); }
HE_HD constexpr int inv_lb(int logn, int k) {
    int u0 = 0;
    for (int i = 0; i < k; ++i) u0 += inv_c(logn, i);
    return u0;
}
// NARROW inverse bounds, in units of p.  A stage with inputs < b p outputs x' < 2 b p and y' < 4p, so the bound
// evolves as b -> max(2b, 4), and needs 2 b p < 2^64, i.e. b <= 256 for p < 2^55.
HE_HD constexpr int inv_stage_bound(int b) { return 2 * b > 4 ? 2 * b : 4; }
HE_HD constexpr int inv_bound_after(int b, int stages) {
    for (int i = 0; i < stages; ++i) b = inv_stage_bound(b);
    return b;
}
HE_HD constexpr bool inv_pass_fits(int b, int stages) {  // every stage of the pass sees inputs <= 256 p
    for (int i = 0; i < stages; ++i) {
        if (b > 256) return false;
        b = inv_stage_bound(b);
    }
    return true;
}
// whether pass k reduces its inputs on entry, and the bound on its inputs after that
HE_HD constexpr bool inv_reduce_at(int logn, int k) {
    int b = 1;
    bool red = false;
    for (int i = 0; i <= k; ++i) {
        const int c = inv_c(logn, i);
        red = !inv_pass_fits(b, c);
        if (red) b = 1;
        b = inv_bound_after(b, c);
    }
    return red;
}
HE_HD constexpr int inv_bound_in(int logn, int k) {
    int b = 1;
    for (int i = 0; i <= k; ++i) {
        const int c = inv_c(logn, i);
        if (!inv_pass_fits(b, c)) b = 1;
        if (i < k) b = inv_bound_after(b, c);
    }
    return b;
}

HE_HD int smem_phys(int e) { return e + (e >> 4); }
HE_HD constexpr int smem_words(int logn) { return (1 << logn) + (1 << (logn - 4)); }

HE_HD ulonglong2 ld_tw(const ulonglong2 *p) {
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}

template <int LOGN, int LB, int C>
HE_HD int elem_index(int tau, int g, int a) {
    constexpr int T = (1 << LOGN) / 16;
    const int sb = tau + g * T;
    const int hi = sb >> LB, lo = sb & ((1 << LB) - 1);
    return (hi << (LB + C)) | (a << LB) | lo;
}

// ---- register <-> memory moves.  Within a sub-block the bit fields of e are disjoint, so both the element index
// and its padded shared-memory position are (per-thread base) + (compile-time offset of a):
//     e = base_e + (a << LB),   phys(e) = phys(base_e) + (a << LB) + ((a << LB) >> 4)
template <int LB>
HE_HD constexpr int smem_off(int a) { return (a << LB) + ((a << LB) >> 4); }

template <int LOGN, int LB, int C>
HE_HD void load_smem(u64 (&x)[16], const u64 *sm, int tau) {
#pragma unroll
    for (int g = 0; g < (16 >> C); ++g) {
        const u64 *b = sm + smem_phys(elem_index<LOGN, LB, C>(tau, g, 0));
#pragma unroll
        for (int a = 0; a < (1 << C); ++a) x[g * (1 << C) + a] = b[smem_off<LB>(a)];
    }
}
template <int LOGN, int LB, int C>
HE_HD void store_smem(const u64 (&x)[16], u64 *sm, int tau) {
#pragma unroll
    for (int g = 0; g < (16 >> C); ++g) {
        u64 *b = sm + smem_phys(elem_index<LOGN, LB, C>(tau, g, 0));
#pragma unroll
        for (int a = 0; a < (1 << C); ++a) b[smem_off<LB>(a)] = x[g * (1 << C) + a];
    }
}
// global side: when LB == 0 a thread's 2^C elements of one sub-block are contiguous -> 16-byte vectors
template <int LOGN, int LB, int C>
HE_HD void load_global(u64 (&x)[16], const u64 *src, int tau) {
#pragma unroll
    for (int g = 0; g < (16 >> C); ++g) {
        const u64 *b = src + elem_index<LOGN, LB, C>(tau, g, 0);
        if (LB == 0) {
#pragma unroll
            for (int a = 0; a < (1 << C); a += 2) {
                const ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(b + a);
                x[g * (1 << C) + a] = v.x;
                x[g * (1 << C) + a + 1] = v.y;
            }
        } else {
#pragma unroll
            for (int a = 0; a < (1 << C); ++a) x[g * (1 << C) + a] = b[a << LB];
        }
    }
}
template <int LOGN, int LB, int C>
HE_HD void store_global(const u64 (&x)[16], u64 *dst, int tau) {
#pragma unroll
    for (int g = 0; g < (16 >> C); ++g) {
        u64 *b = dst + elem_index<LOGN, LB, C>(tau, g, 0);
        if (LB == 0) {
#pragma unroll
            for (int a = 0; a < (1 << C); a += 2)
                *reinterpret_cast<ulonglong2 *>(b + a) = make_ulonglong2(x[g * (1 << C) + a], x[g * (1 << C) + a + 1]);
        } else {
#pragma unroll
            for (int a = 0; a < (1 << C); ++a) b[a << LB] = x[g * (1 << C) + a];
        }
    }
}

struct RowMod {
    u64 p, two_p, mu1;
    u64 np;  // 2^64 - p: lets the Shoup product be all multiply-adds (x*w + q*np)
    u64 four_p;
    int red_shift;   // bits(p) - 7
    u32 red_recip;   // floor(2^(red_shift + 18) / p) < 2^12
    const ulonglong2 *tw;
    u64 c0, c0p, c1, c1p;  // inverse: final-stage scalings
};

// ------------------------------------------------------------------------------------------------ forward
// x*w mod p in [0, 2p) for any x (Shoup), written as multiply-adds only
HE_HD u64 shoup_lazy_np(u64 x, u64 w, u64 wp, u64 np) { return x * w + mulhi64(x, wp) * np; }

// floor(a b / 2^64) - {0,1,2}: three of the four partial products (drops lo*lo and the low halves of the cross terms)
HE_HD u64 mulhi64_lo3(u64 a, u64 b) {
    const u32 al = (u32)a, ah = (u32)(a >> 32), bl = (u32)b, bh = (u32)(b >> 32);
    const u64 t1 = (u64)ah * bl, t2 = (u64)al * bh;
    return (u64)ah * bh + (t1 >> 32) + (t2 >> 32);
}
// x*w mod p in [0, 4p) for any x: Shoup with the under-estimated quotient above
HE_HD u64 shoup_lazy4(u64 x, u64 w, u64 wp, u64 np) { return x * w + mulhi64_lo3(x, wp) * np; }

// x mod p for x < 512 p, p < 2^55: quotient estimated with one 32-bit multiply (never above, at most 2 below)
HE_HD u64 reduce_small(u64 x, const u64 p, const int shift, const u32 recip) {
    const u32 xs = (u32)(x >> shift);            // < 2^16
    const u32 qhat = (xs * recip) >> 18;         // < 2^10
    u64 r = x - (u64)qhat * p;                   // [0, 3p)
    r = csub(r, 2 * p);
    return csub(r, p);
}

template <bool NARROW>
HE_HD void ct_butterfly(u64 &x, u64 &y, const ulonglong2 w, const RowMod &m) {
    if (NARROW) {
        const u64 v = shoup_lazy_np(y, w.x, w.y, m.np);
        const u64 xo = x + v;
        y = x - v + m.two_p;
        x = xo;
    } else {
        const u64 xr = csub(x, m.two_p);
        const u64 v = shoup_lazy_np(y, w.x, w.y, m.np);
        x = xr + v;
        y = xr - v + m.two_p;
    }
}

template <int LOGN, int LB, int C, bool NARROW>
HE_HD void fwd_pass(u64 (&x)[16], int tau, const RowMod &m) {
    constexpr int T = (1 << LOGN) / 16, G = 16 >> C, S0 = LOGN - LB - C;
#pragma unroll
    for (int g = 0; g < G; ++g) {
        const int hi = (tau + g * T) >> LB;
#pragma unroll
        for (int j = 0; j < C; ++j) {
            const int h = 1 << (C - 1 - j);
#pragma unroll
            for (int grp = 0; grp < (1 << j); ++grp) {
                const ulonglong2 w = ld_tw(m.tw + ((1 << (S0 + j)) + (hi << j) + grp));
#pragma unroll
                for (int k = 0; k < h; ++k) {
                    const int a = grp * 2 * h + k;
                    ct_butterfly<NARROW>(x[g * (1 << C) + a], x[g * (1 << C) + a + h], w, m);
                }
            }
        }
    }
}

// reduce the outputs of the last forward stage to canonical residues
template <int LOGN, bool NARROW>
HE_HD void fwd_finish(u64 (&x)[16], const RowMod &m) {
#pragma unroll
    for (int r = 0; r < 16; ++r) {
        if (NARROW) x[r] = reduce_small(x[r], m.p, m.red_shift, m.red_recip);  // < (2 + 4 LOGN) p
        else x[r] = csub(csub(x[r], m.two_p), m.p);           // < 4p
    }
}

// ------------------------------------------------------------------------------------------------ inverse
template <bool NARROW>
HE_HD void gs_butterfly(u64 &x, u64 &y, const ulonglong2 w, const RowMod &m, u64 kp) {
    if (NARROW) {  // inputs < kp (a multiple of p), outputs x < 2 kp, y < 2p
        const u64 s = x + y;
        y = shoup_lazy_np(x - y + kp, w.x, w.y, m.np);
        x = s;
    } else {  // inputs < 2p, outputs < 2p
        const u64 s = csub(x + y, m.two_p);
        y = shoup_lazy_np(x - y + m.two_p, w.x, w.y, m.np);
        x = s;
    }
}

// one inverse stage (local index J) on the sub-block held in registers base .. base + 2^C - 1
template <int LOGN, int LB, int C, bool NARROW, int BIN, int J>
HE_HD void inv_stage(u64 (&x)[16], const int base, const int hi, const RowMod &m) {
    constexpr int HH = 1 << J;
    constexpr bool kLast = (LB + J == LOGN - 1);
    constexpr int kGroups = 1 << (LOGN - 1 - LB - J);
    const u64 kp = m.p * (u64)inv_bound_after(BIN, J);  // inputs of this stage are < kp
#pragma unroll
    for (int grp = 0; grp < (1 << (C - 1 - J)); ++grp) {
        if (!kLast) {
            const ulonglong2 w = ld_tw(m.tw + (kGroups + (hi << (C - 1 - J)) + grp));
#pragma unroll
            for (int k = 0; k < HH; ++k) {
                const int a = grp * 2 * HH + k;
                gs_butterfly<NARROW>(x[base + a], x[base + a + HH], w, m, kp);
            }
        } else {
#pragma unroll
            for (int k = 0; k < HH; ++k) {
                const int a = grp * 2 * HH + k;
                const u64 xa = x[base + a], ya = x[base + a + HH];
                const u64 s = xa + ya;
                const u64 d = xa - ya + (NARROW ? kp : m.two_p);
                x[base + a] = shoup_mul(s, m.c0, m.c0p, m.p);        // (x + y) N^-1      (PolyRq+Ntt.swift:416-419)
                x[base + a + HH] = shoup_mul(d, m.c1, m.c1p, m.p);   // (x - y) N^-1 psi^-(N/2)
            }
        }
    }
}

// BIN = bound (units of p) on the pass inputs (NARROW only; ignored for WIDE)
template <int LOGN, int LB, int C, bool NARROW, int BIN>
HE_HD void inv_pass(u64 (&x)[16], int tau, const RowMod &m) {
    constexpr int T = (1 << LOGN) / 16, G = 16 >> C;
#pragma unroll
    for (int g = 0; g < G; ++g) {
        const int hi = (tau + g * T) >> LB;
        const int base = g * (1 << C);
        inv_stage<LOGN, LB, C, NARROW, BIN, 0>(x, base, hi, m);
        if (C > 1) inv_stage<LOGN, LB, C, NARROW, BIN, (C > 1 ? 1 : 0)>(x, base, hi, m);
        if (C > 2) inv_stage<LOGN, LB, C, NARROW, BIN, (C > 2 ? 2 : 0)>(x, base, hi, m);
        if (C > 3) inv_stage<LOGN, LB, C, NARROW, BIN, (C > 3 ? 3 : 0)>(x, base, hi, m);
    }
}

template <int BIN>
HE_HD void inv_reduce(u64 (&x)[16], const RowMod &m) {
#pragma unroll
    for (int r = 0; r < 16; ++r) x[r] = reduce_small(x[r], m.p, m.red_shift, m.red_recip);  // < 512 p
}

}  // namespace fast
}  // namespace hecuda
