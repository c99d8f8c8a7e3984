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

// the emulated "registers" of thread tau are x[tau*16 .. tau*16+15]; memory moves use the kernels' own helpers
template <int LOGN, int LB, int C>
static void emu_load(std::vector<u64> &x, const u64 *src, int tau, bool smem) {
    u64(&xr)[16] = *reinterpret_cast<u64(*)[16]>(&x[tau * 16]);
    if (smem) load_smem<LOGN, LB, C>(xr, src, tau);
    else load_global<LOGN, LB, C>(xr, src, tau);
}
template <int LOGN, int LB, int C>
static void emu_store(std::vector<u64> &x, u64 *dst, int tau, bool smem) {
    u64(&xr)[16] = *reinterpret_cast<u64(*)[16]>(&x[tau * 16]);
    if (smem) store_smem<LOGN, LB, C>(xr, dst, tau);
    else store_global<LOGN, LB, C>(xr, dst, tau);
}

template <int LOGN, bool NARROW, int K>
static void fwd_pass_k(std::vector<u64> &x, std::vector<u64> &sm, const u64 *src, u64 *dst, const RowMod &m) {
    constexpr int P = plan_passes(LOGN), T = (1 << LOGN) / 16;
    constexpr int C = fwd_c(LOGN, K), LB = fwd_lb(LOGN, K);
    for (int tau = 0; tau < T; ++tau) {  // phase 1: every thread reads
        if (K == 0) emu_load<LOGN, LB, C>(x, src, tau, false);
        else emu_load<LOGN, LB, C>(x, sm.data(), tau, true);
    }
    for (int tau = 0; tau < T; ++tau) {
        u64(&xr)[16] = *reinterpret_cast<u64(*)[16]>(&x[tau * 16]);
        fwd_pass<LOGN, LB, C, NARROW>(xr, tau, m);
        if (K == P - 1) fwd_finish<LOGN, NARROW>(xr, m);
    }
    for (int tau = 0; tau < T; ++tau) {
        if (K == P - 1) emu_store<LOGN, LB, C>(x, dst, tau, false);
        else emu_store<LOGN, LB, C>(x, sm.data(), tau, true);
    }
}
template <int LOGN, bool NARROW, int K>
static void inv_pass_k(std::vector<u64> &x, std::vector<u64> &sm, const u64 *src, u64 *dst, const RowMod &m) {
    constexpr int P = plan_passes(LOGN), T = (1 << LOGN) / 16;
    constexpr int C = inv_c(LOGN, K), LB = inv_lb(LOGN, K);
    for (int tau = 0; tau < T; ++tau) {
        if (K == 0) emu_load<LOGN, LB, C>(x, src, tau, false);
        else emu_load<LOGN, LB, C>(x, sm.data(), tau, true);
    }
    for (int tau = 0; tau < T; ++tau) {
        u64(&xr)[16] = *reinterpret_cast<u64(*)[16]>(&x[tau * 16]);
        if (NARROW && K > 0 && inv_reduce_at(LOGN, K)) inv_reduce<0>(xr, m);
        inv_pass<LOGN, LB, C, NARROW, inv_bound_in(LOGN, K)>(xr, tau, m);
    }
    for (int tau = 0; tau < T; ++tau) {
        if (K == P - 1) emu_store<LOGN, LB, C>(x, dst, tau, false);
        else emu_store<LOGN, LB, C>(x, sm.data(), tau, true);
    }
}

template <int LOGN, bool NARROW>
static void run(bool inverse, const u64 *src, u64 *dst, const RowMod &m) {
    constexpr int P = plan_passes(LOGN), T = (1 << LOGN) / 16;
    std::vector<u64> x(T * 16), sm(smem_words(LOGN), 0xDEADBEEFDEADBEEFull);
    if (!inverse) {
        fwd_pass_k<LOGN, NARROW, 0>(x, sm, src, dst, m);
        fwd_pass_k<LOGN, NARROW, 1>(x, sm, src, dst, m);
        if (P == 4) fwd_pass_k<LOGN, NARROW, (P == 4 ? 2 : 1)>(x, sm, src, dst, m);
        fwd_pass_k<LOGN, NARROW, P - 1>(x, sm, src, dst, m);
    } else {
        inv_pass_k<LOGN, NARROW, 0>(x, sm, src, dst, m);
        inv_pass_k<LOGN, NARROW, 1>(x, sm, src, dst, m);
        if (P == 4) inv_pass_k<LOGN, NARROW, (P == 4 ? 2 : 1)>(x, sm, src, dst, m);
        inv_pass_k<LOGN, NARROW, P - 1>(x, sm, src, dst, m);
    }
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
    RowMod m;
    m.p = p;
    m.two_p = 2 * p;
    m.mu1 = (u64)(((unsigned __int128)1 << 64) / p);
    m.np = 0 - p;
    m.four_p = 4 * p;
    m.red_shift = host::bit_length(p) > 7 ? host::bit_length(p) - 7 : 0;
    m.red_recip = (u32)((((unsigned __int128)1) << (m.red_shift + 18)) / p);
    m.tw = inverse ? itw.data() : tw.data();
    u64 n_inv = host::invmod((u64)n % p, p);
    if (t) n_inv = host::mulmod(n_inv, t % p, p);
    const u64 n_inv_w = host::mulmod(n_inv, itw[1].x, p);
    m.c0 = n_inv; m.c0p = host::shoup_factor(n_inv, p);
    m.c1 = n_inv_w; m.c1p = host::shoup_factor(n_inv_w, p);
    const bool narrow = host::bit_length(p) <= kNarrowBits;
#define RUN(L) case L: if (narrow) run<L, true>(inverse, in.data(), out.data(), m); else run<L, false>(inverse, in.data(), out.data(), m); break;
    switch (logn) { RUN(10) RUN(11) RUN(12) RUN(13) RUN(14) default: return 4; }
    for (int i = 0; i < n; ++i) printf("%llu\n", out[i]);
    return 0;
}
