// ntt_emulate.cu -- host-side SIMT emulation of the register-tiled NTT (csrc/ntt_fast.cuh).
// Replays, thread by thread and pass by pass, exactly the index maps / butterflies / lazy-reduction schedule the
// sm_100a kernels use (same __host__ __device__ functions), and writes the result so the Python test can compare it
// with the oracle.  Built and run on the CPU by tests/test_ntt_emulation.py (there is no GPU in the build container).
//   usage: ntt_emulate <logn> <p> <fwd|inv> <scale_t> < input(u64 text)  > output
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../swift-homomorphic-encryption_b200/csrc/hostmath.hpp"
#include "../../swift-homomorphic-encryption_b200/csrc/ntt_fast.cuh"

using namespace hecuda;
using namespace hecuda::fast;

// the emulated "registers" of thread tau are x[tau*16 .. tau*16+15]; the row lives in the kernels' swizzled
// shared-memory layout for the whole transform, moved in and out line by line exactly like the TMA copies
template <int LOGN, int CLS, int K>
static void fwd_pass_k(std::vector<u64> &x, std::vector<u64> &sm, const RowMod &m) {
    constexpr int P = plan_passes(LOGN), T = (1 << LOGN) / 16;
    constexpr int C = fwd_c(LOGN, K), LB = fwd_lb(LOGN, K);
    for (int tau = 0; tau < T; ++tau) {  // phase 1: every thread reads
        u64(&xr)[16] = *reinterpret_cast<u64(*)[16]>(&x[tau * 16]);
        load_smem<LOGN, LB, C>(xr, sm.data(), tau);
    }
    for (int tau = 0; tau < T; ++tau) {
        u64(&xr)[16] = *reinterpret_cast<u64(*)[16]>(&x[tau * 16]);
        fwd_pass<LOGN, LB, C, CLS>(xr, tau, m);
        if (K == P - 1) fwd_finish<CLS>(xr, m);
    }
    for (int tau = 0; tau < T; ++tau) {
        u64(&xr)[16] = *reinterpret_cast<u64(*)[16]>(&x[tau * 16]);
        store_smem<LOGN, LB, C>(xr, sm.data(), tau);
    }
}
template <int LOGN, int CLS, int K>
static void inv_pass_k(std::vector<u64> &x, std::vector<u64> &sm, const RowMod &m) {
    constexpr int T = (1 << LOGN) / 16;
    constexpr int C = inv_c(LOGN, K), LB = inv_lb(LOGN, K);
    for (int tau = 0; tau < T; ++tau) {
        u64(&xr)[16] = *reinterpret_cast<u64(*)[16]>(&x[tau * 16]);
        load_smem<LOGN, LB, C>(xr, sm.data(), tau);
    }
    for (int tau = 0; tau < T; ++tau) {
        u64(&xr)[16] = *reinterpret_cast<u64(*)[16]>(&x[tau * 16]);
        if (narrow_like(CLS) && K > 0 && inv_reduce_at(LOGN, K)) inv_reduce(xr, m);
        inv_pass<LOGN, LB, C, CLS, inv_bound_in(LOGN, K)>(xr, tau, m);
    }
    for (int tau = 0; tau < T; ++tau) {
        u64(&xr)[16] = *reinterpret_cast<u64(*)[16]>(&x[tau * 16]);
        store_smem<LOGN, LB, C>(xr, sm.data(), tau);
    }
}

template <int LOGN, int CLS>
static void run(bool inverse, const u64 *src, u64 *dst, const RowMod &m) {
    constexpr int P = plan_passes(LOGN), T = (1 << LOGN) / 16;
    std::vector<u64> x(T * 16), sm(smem_words(LOGN), 0xDEADBEEFDEADBEEFull);
    // "TMA in" with CU_TENSOR_MAP_SWIZZLE_128B: 16-byte chunk c of line l lands at chunk c ^ (l & 7)
    for (int line = 0; line < T; ++line)
        for (int c = 0; c < 8; ++c)
            memcpy(&sm[kLineWords * line + 2 * (c ^ (line & 7))], src + kLineWords * line + 2 * c, 16);
    if (!inverse) {
        fwd_pass_k<LOGN, CLS, 0>(x, sm, m);
        fwd_pass_k<LOGN, CLS, 1>(x, sm, m);
        if (P == 4) fwd_pass_k<LOGN, CLS, (P == 4 ? 2 : 1)>(x, sm, m);
        fwd_pass_k<LOGN, CLS, P - 1>(x, sm, m);
    } else {
        inv_pass_k<LOGN, CLS, 0>(x, sm, m);
        inv_pass_k<LOGN, CLS, 1>(x, sm, m);
        if (P == 4) inv_pass_k<LOGN, CLS, (P == 4 ? 2 : 1)>(x, sm, m);
        inv_pass_k<LOGN, CLS, P - 1>(x, sm, m);
    }
    for (int line = 0; line < T; ++line)  // "TMA out": the same swizzle, undone by the copy engine
        for (int c = 0; c < 8; ++c)
            memcpy(dst + kLineWords * line + 2 * c, &sm[kLineWords * line + 2 * (c ^ (line & 7))], 16);
}

int main(int argc, char **argv) {
    if (argc < 5) return 2;
    const int logn = atoi(argv[1]);
    const u64 p = strtoull(argv[2], nullptr, 10);
    const bool inverse = !strcmp(argv[3], "inv");
    const u64 t = strtoull(argv[4], nullptr, 10);  // 0 = no t scaling
    const int n = 1 << logn;
    std::vector<u64> in(n), out(n);
    for (int i = 0; i < n; ++i)
        if (scanf("%llu", &in[i]) != 1) return 3;
    // tables exactly as context.cu builds them
    const u64 psi = host::min_primitive_root(2 * (u64)n, p), psi_inv = host::invmod(psi, p);
    std::vector<ulonglong2> tw(n), itw(n);
    u64 pw = 1, ipw = 1;
    for (int i = 0; i < n; ++i) {
        unsigned r = host::bitrev((unsigned)i, logn);
        tw[r] = make_ulonglong2(pw, host::shoup_factor(pw, p));
        itw[r] = make_ulonglong2(ipw, host::shoup_factor(ipw, p));
        pw = host::mulmod(pw, psi, p);
        ipw = host::mulmod(ipw, psi_inv, p);
    }
    // transposed tables of the line-owning pass, exactly as context.cu builds them
    const int threads = n / 16;
    std::vector<ulonglong2> tr(15 * (size_t)threads), itr(15 * (size_t)threads);
    for (int k = 0; k < 15; ++k)
        for (int tau = 0; tau < threads; ++tau) {
            tr[(size_t)k * threads + tau] = tw[fwd_last_source(logn, k, tau)];
            itr[(size_t)k * threads + tau] = itw[inv_first_source(logn, k, tau)];
        }
    ModSlot slot;
    memset(&slot, 0, sizeof(slot));
    slot.p = p;
    slot.mu1 = (u64)(((unsigned __int128)1 << 64) / p);
    slot.bits = host::bit_length(p);
    slot.red_shift = slot.bits > 12 ? slot.bits - 12 : 0;
    slot.red_recip = (u32)((((unsigned __int128)1) << (slot.red_shift + 32)) / p);
    u64 n_inv = host::invmod((u64)n % p, p);
    if (t) n_inv = host::mulmod(n_inv, t % p, p);
    const u64 n_inv_w = host::mulmod(n_inv, itw[1].x, p);
    slot.inv_scale[0].c0 = n_inv; slot.inv_scale[0].c0p = host::shoup_factor(n_inv, p);
    slot.inv_scale[0].c1 = n_inv_w; slot.inv_scale[0].c1p = host::shoup_factor(n_inv_w, p);
    const int cls = class_of_modulus(p, slot.bits);
    RowMod m;
    m.np = 0 - p;
    m.kp = (cls == kWide || cls == kSmall) ? 2 * p : 4 * p;
    m.tw = inverse ? itw.data() : tw.data();
    slot.tw_t = tr.data();
    slot.itw_t = itr.data();
    m.slot = &slot;
    m.scale_mode = inverse ? 0 : -1;
    m.partial = false;
#define RUN(L) case L: if (cls == kNarrowH) run<L, kNarrowH>(inverse, in.data(), out.data(), m); else if (cls == kNarrow) run<L, kNarrow>(inverse, in.data(), out.data(), m); else if (cls == kSmall) run<L, kSmall>(inverse, in.data(), out.data(), m); else if (cls == kMid) run<L, kMid>(inverse, in.data(), out.data(), m); else run<L, kWide>(inverse, in.data(), out.data(), m); break;
    switch (logn) { RUN(10) RUN(11) RUN(12) RUN(13) RUN(14) default: return 4; }
    for (int i = 0; i < n; ++i) printf("%llu\n", out[i]);
    return 0;
}
