/*
 * he_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE ONLY; see he_oracle.h).
 *
 * Restates, step for step, the algorithms of apple/swift-homomorphic-encryption's RNS-BFV hot path in
 * plain C (unsigned __int128).  Hot loops keep the reference's structure (Harvey lazy butterflies,
 * Barrett / Shoup reductions, 128-bit lazy accumulators) so that timing this file is a fair
 * "C restatement of the Swift reference" CPU baseline.  Citations are relative to /root/reference/.
 */
#include "he_oracle.h"

#include <malloc.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef unsigned __int128 u128;
typedef uint64_t u64;
typedef int64_t i64;

#define ORC_MAX_MODULI 80 /* 32 coefficient moduli (EncryptionParameters.swift:148) + Bsk + m~ */

/* =====================================================================================
 * Scalar arithmetic -- Sources/ModularArithmetic/Scalar.swift, Modulus.swift
 * ===================================================================================== */

static inline u64 mul_hi(u64 a, u64 b) { return (u64)(((u128)a * b) >> 64); }

/* subtractIfExceeds, Scalar.swift:160-166 (branch-free sign-mask form) */
static inline u64 csub(u64 x, u64 p) {
    u64 d = x - p;
    u64 mask = (u64)0 - (d >> 63);
    return d + (p & mask);
}
static inline u64 add_mod(u64 a, u64 b, u64 p) { return csub(a + b, p); }          /* Scalar.swift:146-152 */
static inline u64 sub_mod(u64 a, u64 b, u64 p) { return csub(a + p - b, p); }      /* Scalar.swift:187-192 */
static inline u64 neg_mod(u64 a, u64 p) { return csub(p - a, p); }                 /* Scalar.swift:174-177 */

static inline int sig_bits(u64 x) { return x ? 64 - __builtin_clzll(x) : 0; }
static inline int ilog2(u64 x) { return 63 - __builtin_clzll(x); }
static inline int is_pow2(u64 x) { return x && !(x & (x - 1)); }

/* Modulus<T> with its three ReduceModulus factors, Modulus.swift:18-111,169-242 */
typedef struct {
    u64 p;
    u64 single_factor; /* floor(2^64 / p)              (Modulus.swift:206-209) */
    u128 double_factor; /* floor(2^128 / p)            (Modulus.swift:224-232) */
    int nbits;          /* significantBitCount          */
    u64 prod_factor;    /* floor(2^(nbits+62) / p)     (Modulus.swift:235-240) */
} modulus_t;

static modulus_t modulus_make(u64 p) {
    modulus_t m;
    m.p = p;
    m.single_factor = (u64)((((u128)1) << 64) / p);
    if (is_pow2(p)) {
        m.double_factor = ((u128)1) << (128 - ilog2(p));
    } else {
        m.double_factor = (~(u128)0) / p;
    }
    m.nbits = sig_bits(p);
    m.prod_factor = (u64)((((u128)1) << (m.nbits + 62)) / p);
    return m;
}

/* ReduceModulus.reduce(_ x: T), Modulus.swift:258-263 */
static inline u64 reduce_single(const modulus_t *m, u64 x) {
    u64 q = mul_hi(x, m->single_factor);
    u64 z = x - q * m->p;
    return csub(z, m->p);
}

/* high 128 bits of a 128x128 product */
static inline u128 mul_hi128(u128 x, u128 f) {
    u64 xl = (u64)x, xh = (u64)(x >> 64), fl = (u64)f, fh = (u64)(f >> 64);
    u128 ll = (u128)xl * fl, lh = (u128)xl * fh, hl = (u128)xh * fl, hh = (u128)xh * fh;
    u128 mid = (ll >> 64) + (u64)lh + (u64)hl;
    return hh + (lh >> 64) + (hl >> 64) + (mid >> 64);
}

/* ReduceModulus.reduce(_ x: T.DoubleWidth), Modulus.swift:319-325 */
static inline u64 reduce_double(const modulus_t *m, u128 x) {
    u128 qhat = mul_hi128(x, m->double_factor);
    u128 qp = qhat * (u128)m->p;
    u128 z = x - qp;
    return csub((u64)z, m->p);
}

/* ReduceModulus.reduceProduct(_ x) for x < p^2, Modulus.swift:349-360 */
static inline u64 reduce_product(const modulus_t *m, u128 x) {
    u64 xs = (u64)(x >> (m->nbits - 2));
    u64 q = mul_hi(xs, m->prod_factor);
    u64 z = (u64)x - q * m->p;
    return csub(z, m->p);
}
static inline u64 mul_mod(const modulus_t *m, u64 a, u64 b) { return reduce_product(m, (u128)a * b); }

/* MultiplyConstantModulus, Modulus.swift:377-416; factor from HomomorphicEncryption/Modulus.swift:92-103 */
typedef struct {
    u64 w, wp, p;
} shoup_t;
static shoup_t shoup_make(u64 w, u64 p) {
    shoup_t s;
    s.w = w;
    s.p = p;
    s.wp = (u64)((((u128)w) << 64) / p);
    return s;
}
static inline u64 shoup_lazy(const shoup_t *s, u64 x) { /* multiplyModLazy -> [0, 2p) */
    u64 q = mul_hi(x, s->wp);
    return x * s->w - q * s->p;
}
static inline u64 shoup_mul(const shoup_t *s, u64 x) { return csub(shoup_lazy(s, x), s->p); }

/* slow exact helpers for setup-time constants */
static u64 slow_mul_mod(u64 a, u64 b, u64 p) { return (u64)(((u128)a * b) % p); }

uint64_t orc_pow_mod(uint64_t base, uint64_t exp, uint64_t p) { /* Scalar.swift:207-229 */
    if (p == 1) return 0;
    u64 result = 1 % p;
    base %= p;
    while (exp) {
        if (exp & 1) result = slow_mul_mod(result, base, p);
        base = slow_mul_mod(base, base, p);
        exp >>= 1;
    }
    return result;
}

/* inverseMod, HomomorphicEncryption/Scalar.swift:76-96 (signed extended Euclid) */
uint64_t orc_inverse_mod(uint64_t a_in, uint64_t modulus) {
    if (a_in == 0 || modulus == 0) return 0;
    i64 a = (i64)a_in, m = (i64)modulus, x0 = 0, inv = 1;
    while (a > 1) {
        if (m == 0) return 0;
        inv -= (a / m) * x0;
        a %= m;
        i64 tmp = a; a = m; m = tmp;
        tmp = x0; x0 = inv; inv = tmp;
    }
    if (inv < 0) inv += (i64)modulus;
    if (slow_mul_mod((u64)inv, a_in % modulus, modulus) != 1 % modulus) return 0;
    return (u64)inv;
}

/* isPrime, HomomorphicEncryption/Scalar.swift:160-202 (Miller-Rabin, 12 fixed bases) */
int orc_is_prime(uint64_t n) {
    static const u64 bases[12] = {2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37};
    if (n <= 1) return 0;
    for (int i = 0; i < 12; i++) {
        if (n == bases[i]) return 1;
        if (n % bases[i] == 0) return 0;
    }
    int r = 0;
    u64 d = n - 1;
    while ((d & 1) == 0) { d >>= 1; r++; }
    for (int i = 0; i < 12; i++) {
        u64 x = orc_pow_mod(bases[i], d, n);
        if (x == 1 || x == n - 1) continue;
        int witness = 1;
        for (int k = 0; k < r; k++) {
            x = slow_mul_mod(x, x, n);
            if (x == n - 1) { witness = 0; break; }
        }
        if (witness) return 0;
    }
    return 1;
}

/* generatePrimes, HomomorphicEncryption/Scalar.swift:113-154 */
int orc_generate_primes(const int32_t *bit_counts, int32_t n, int32_t prefer_small, int64_t ntt_degree,
                        uint64_t *out) {
    int found = 0;
    for (int k = 0; k < n; k++) {
        int b = bit_counts[k];
        if (b < 2 || b > 64) return found;
        u64 upper = (b == 64) ? ~(u64)0 : ((u64)1 << b); /* exclusive */
        u64 lower = (u64)1 << (b - 1);
        u64 step = (u64)(2 * ntt_degree);
        u64 cand = prefer_small ? lower + 1 : (upper - step) + 1;
        while (cand >= lower && cand < upper) {
            int dup = 0;
            for (int j = 0; j < found; j++) dup |= (out[j] == cand);
            /* isNttModulus: PolyRq+Ntt.swift:24-27 */
            if (!dup && orc_is_prime(cand) && (cand % (u64)(2 * ntt_degree) == 1) && cand != 1) {
                out[found++] = cand;
                break;
            }
            if (prefer_small) {
                if (cand > ~(u64)0 - step) break;
                cand += step;
            } else {
                if (cand < step) break;
                cand -= step;
            }
        }
    }
    return found;
}

/* reverseBits, ModularArithmetic/Scalar.swift:238-253 */
uint32_t orc_reverse_bits(uint32_t x, int32_t bit_count) {
    x = ((x & 0xAAAAAAAAu) >> 1) | ((x & 0x55555555u) << 1);
    x = ((x & 0xCCCCCCCCu) >> 2) | ((x & 0x33333333u) << 2);
    x = ((x & 0xF0F0F0F0u) >> 4) | ((x & 0x0F0F0F0Fu) << 4);
    x = ((x & 0xFF00FF00u) >> 8) | ((x & 0x00FF00FFu) << 8);
    x = (x >> 16) | (x << 16);
    x >>= (32 - bit_count);
    return x;
}

/* minPrimitiveRootOfUnity, PolyRq+Ntt.swift:87-105.  The reference finds *a* primitive root by random
 * trials (:45-79) and then takes the minimum over all odd powers, so the result does not depend on
 * which generator was found; we find one deterministically. */
uint64_t orc_min_primitive_root(int64_t degree, uint64_t p) {
    if (degree < 2 || (degree & (degree - 1)) || (p - 1) % (u64)degree != 0) return 0;
    u64 root = 0;
    for (u64 g = 2; g < p && g < 100000; g++) {
        u64 r = orc_pow_mod(g, (p - 1) / (u64)degree, p);
        if (orc_pow_mod(r, (u64)degree / 2, p) == p - 1) { root = r; break; } /* :30-37 */
    }
    if (!root) return 0;
    u64 smallest = root, cur = root, g2 = slow_mul_mod(root, root, p);
    for (i64 i = 0; i < degree / 2; i++) {
        if (cur < smallest) smallest = cur;
        cur = slow_mul_mod(cur, g2, p);
    }
    return smallest;
}

uint64_t orc_barrett_reduce_single(uint64_t x, uint64_t p) { modulus_t m = modulus_make(p); return reduce_single(&m, x); }
uint64_t orc_barrett_reduce_double(uint64_t hi, uint64_t lo, uint64_t p) {
    modulus_t m = modulus_make(p);
    return reduce_double(&m, (((u128)hi) << 64) | lo);
}
uint64_t orc_barrett_reduce_product(uint64_t x, uint64_t y, uint64_t p) { modulus_t m = modulus_make(p); return mul_mod(&m, x, y); }
uint64_t orc_shoup_mul(uint64_t x, uint64_t w, uint64_t p) { shoup_t s = shoup_make(w, p); return shoup_mul(&s, x); }
uint64_t orc_shoup_mul_lazy(uint64_t x, uint64_t w, uint64_t p) { shoup_t s = shoup_make(w, p); return shoup_lazy(&s, x); }

/* =====================================================================================
 * NTT -- Sources/HomomorphicEncryption/PolyRq/PolyRq+Ntt.swift
 * ===================================================================================== */

typedef struct ntt_tables {
    i64 n;
    u64 p;
    modulus_t mod;
    u64 *roots, *roots_p;         /* rootOfUnityPowers (bit-reversed order) + Shoup factors, :125-143 */
    u64 *inv_roots, *inv_roots_p; /* reordered inverse powers, :145-157 */
    shoup_t inv_degree;           /* :159-160 */
    shoup_t inv_degree_root;      /* :162-168 */
    struct ntt_tables *next;
} ntt_tables;

static ntt_tables *g_ntt_cache = NULL;

static ntt_tables *ntt_tables_build(i64 n, u64 p) {
    int logn = ilog2((u64)n);
    u64 root = orc_min_primitive_root(2 * n, p); /* :120 */
    if (!root) return NULL;
    u64 inv_root = orc_inverse_mod(root, p); /* :123 */
    ntt_tables *t = (ntt_tables *)calloc(1, sizeof(ntt_tables));
    t->n = n;
    t->p = p;
    t->mod = modulus_make(p);
    t->roots = (u64 *)malloc(sizeof(u64) * n);
    t->roots_p = (u64 *)malloc(sizeof(u64) * n);
    t->inv_roots = (u64 *)malloc(sizeof(u64) * n);
    t->inv_roots_p = (u64 *)malloc(sizeof(u64) * n);
    u64 *inv_tmp = (u64 *)malloc(sizeof(u64) * n);
    for (i64 i = 0; i < n; i++) { t->roots[i] = 1; inv_tmp[i] = 1; t->inv_roots[i] = 1; }
    i64 prev = 0;
    for (u64 idx = 1; idx < (u64)n; idx++) { /* :128-137 */
        i64 rev = logn ? (i64)orc_reverse_bits((uint32_t)idx, logn) : 0;
        t->roots[rev] = slow_mul_mod(root, t->roots[prev], p);
        inv_tmp[rev] = slow_mul_mod(inv_root, inv_tmp[prev], p);
        prev = rev;
    }
    i64 inv_idx = 1; /* :146-153 */
    for (int l = logn - 1; l >= 0; l--) {
        i64 m = (i64)1 << l;
        for (i64 i = 0; i < m; i++) t->inv_roots[inv_idx++] = inv_tmp[m + i];
    }
    for (i64 i = 0; i < n; i++) {
        t->roots_p[i] = (u64)((((u128)t->roots[i]) << 64) / p);
        t->inv_roots_p[i] = (u64)((((u128)t->inv_roots[i]) << 64) / p);
    }
    u64 inv_degree = orc_inverse_mod((u64)n % p, p);
    t->inv_degree = shoup_make(inv_degree, p);
    t->inv_degree_root = shoup_make(slow_mul_mod(inv_degree, t->inv_roots[n - 1], p), p);
    free(inv_tmp);
    return t;
}

static const ntt_tables *ntt_tables_get(i64 n, u64 p) {
    ntt_tables *found = NULL;
#pragma omp critical(orc_ntt_cache)
    {
        for (ntt_tables *t = g_ntt_cache; t; t = t->next)
            if (t->n == n && t->p == p) { found = t; break; }
        if (!found) {
            found = ntt_tables_build(n, p);
            if (found) { found->next = g_ntt_cache; g_ntt_cache = found; }
        }
    }
    return found;
}

int orc_ntt_tables(int64_t n, uint64_t p, uint64_t *roots, uint64_t *inv_roots_reordered, uint64_t *inv_degree,
                   uint64_t *inv_degree_root) {
    const ntt_tables *t = ntt_tables_get(n, p);
    if (!t) return -1;
    memcpy(roots, t->roots, sizeof(u64) * n);
    memcpy(inv_roots_reordered, t->inv_roots, sizeof(u64) * n);
    *inv_degree = t->inv_degree.w;
    *inv_degree_root = t->inv_degree_root.w;
    return 0;
}

/* _NttContext.forwardNtt, PolyRq+Ntt.swift:237-319 (Harvey CT butterflies with delayed reduction) */
static void ntt_forward_row(const ntt_tables *t, u64 *d) {
    const i64 n = t->n;
    const u64 p = t->p, twice = p << 1;
    const int logn = ilog2((u64)n);
    i64 lazy = -1;
    const u64 max_lazy = (~(u64)0) / (2 * p) - 1;
    for (int log2m = 0; log2m < logn; log2m++) {
        const i64 m = (i64)1 << log2m;
        const i64 tt = n >> (log2m + 1);
        lazy += 2;
        const int time_to_reduce = (u64)lazy > max_lazy;
        if (time_to_reduce) {
            if (tt == 1) lazy = (lazy - 2 > 2) ? lazy - 2 : 2;
            else lazy = 1;
        }
        if (tt == 1) { /* applyFinalStageOp :251-269 */
            for (i64 i = 0; i < m; i++) {
                shoup_t w = {t->roots[m + i], t->roots_p[m + i], p};
                u64 x = d[2 * i], y = d[2 * i + 1];
                if (time_to_reduce) x = csub(x, twice);
                u64 tw = shoup_lazy(&w, y); /* forwardButterfly :182-201 */
                u64 yo = x + twice - tw;
                u64 xo = x + tw;
                d[2 * i] = reduce_single(&t->mod, xo);
                d[2 * i + 1] = reduce_single(&t->mod, yo);
            }
        } else { /* applyNonFinalStageOp :271-287 */
            for (i64 i = 0; i < m; i++) {
                shoup_t w = {t->roots[m + i], t->roots_p[m + i], p};
                i64 j1 = 2 * i * tt;
                for (i64 j = j1; j < j1 + tt; j++) {
                    u64 x = d[j], y = d[j + tt];
                    if (time_to_reduce) x = reduce_single(&t->mod, x);
                    u64 tw = shoup_lazy(&w, y);
                    d[j + tt] = x + twice - tw;
                    d[j] = x + tw;
                }
            }
        }
    }
}

/* _NttContext.inverseNtt, PolyRq+Ntt.swift:379-483 (GS butterflies, N^{-1} folded into the last stage) */
static void ntt_inverse_row(const ntt_tables *t, u64 *d) {
    const i64 n = t->n;
    const u64 p = t->p;
    const int logn = ilog2((u64)n);
    const int clz = __builtin_clzll(p);
    const int multiples = (logn + 1 < clz) ? logn + 1 : clz;
    i64 root_idx = 1;
    int lazy = -1;
    const i64 ndiv2 = n >> 1;
    for (int log2m = logn - 1; log2m >= 0; log2m--) {
        const i64 m = (i64)1 << log2m;
        const i64 tt = n >> (log2m + 1);
        lazy += 1;
        const int time_to_reduce = (lazy == multiples);
        if (time_to_reduce) {
            if (m == 1) lazy -= 1;
            else lazy = 0;
        }
        const u64 kp = p << lazy;
        if (m == 1) { /* :407-430 */
            for (i64 xi = 0; xi < ndiv2; xi++) {
                i64 yi = xi + ndiv2;
                u64 x = d[xi], y = d[yi];
                if (time_to_reduce) { x = csub(x, kp); y = csub(y, kp); }
                u64 tx = x + y;
                u64 ty = x + kp - y;
                d[xi] = shoup_mul(&t->inv_degree, tx);
                d[yi] = shoup_mul(&t->inv_degree_root, ty);
            }
        } else { /* :431-480; inverseButterfly :359-375 */
            for (i64 i = 0; i < m; i++) {
                shoup_t w = {t->inv_roots[root_idx + i], t->inv_roots_p[root_idx + i], p};
                i64 j1 = 2 * i * tt;
                for (i64 j = j1; j < j1 + tt; j++) {
                    u64 x = d[j], y = d[j + tt];
                    if (time_to_reduce) { x = reduce_single(&t->mod, x); y = reduce_single(&t->mod, y); }
                    u64 tv = x + kp - y;
                    d[j] = x + y;
                    d[j + tt] = shoup_lazy(&w, tv);
                }
            }
        }
        root_idx += m;
    }
    if (logn == 0) { /* degree 1: nothing to do */ }
}

int orc_ntt_forward(int64_t n, const uint64_t *moduli, int32_t nmod, uint64_t *data, int64_t rows) {
    for (i64 r = 0; r < rows; r++) {
        const ntt_tables *t = ntt_tables_get(n, moduli[r % nmod]);
        if (!t) return -1;
        ntt_forward_row(t, data + r * n);
    }
    return 0;
}
/* rows in place on `threads` OpenMP threads (the benchmark's CPU arm: PolyBenchmark forwardNtt over many polynomials) */
int orc_ntt_forward_threads(int64_t n, const uint64_t *moduli, int32_t nmod, uint64_t *data, int64_t rows, int32_t threads) {
    for (int m = 0; m < nmod; m++)
        if (!ntt_tables_get(n, moduli[m])) return -1; /* build the tables before the parallel region */
    if (threads < 1) threads = 1;
#pragma omp parallel for num_threads(threads) schedule(static)
    for (i64 r = 0; r < rows; r++) ntt_forward_row(ntt_tables_get(n, moduli[r % nmod]), data + r * n);
    return 0;
}
int orc_ntt_inverse(int64_t n, const uint64_t *moduli, int32_t nmod, uint64_t *data, int64_t rows) {
    for (i64 r = 0; r < rows; r++) {
        const ntt_tables *t = ntt_tables_get(n, moduli[r % nmod]);
        if (!t) return -1;
        ntt_inverse_row(t, data + r * n);
    }
    return 0;
}

/* =====================================================================================
 * PolyRq coefficient-wise ops -- Sources/HomomorphicEncryption/PolyRq/PolyRq.swift
 * ===================================================================================== */

void orc_poly_add(int64_t n, const uint64_t *moduli, int32_t nmod, uint64_t *lhs, const uint64_t *rhs) { /* :147-157 */
    for (int r = 0; r < nmod; r++)
        for (i64 i = 0; i < n; i++) lhs[r * n + i] = add_mod(lhs[r * n + i], rhs[r * n + i], moduli[r]);
}
void orc_poly_sub(int64_t n, const uint64_t *moduli, int32_t nmod, uint64_t *lhs, const uint64_t *rhs) { /* :164-174 */
    for (int r = 0; r < nmod; r++)
        for (i64 i = 0; i < n; i++) lhs[r * n + i] = sub_mod(lhs[r * n + i], rhs[r * n + i], moduli[r]);
}
void orc_poly_mul(int64_t n, const uint64_t *moduli, int32_t nmod, uint64_t *lhs, const uint64_t *rhs) { /* :184-204 */
    for (int r = 0; r < nmod; r++) {
        modulus_t m = modulus_make(moduli[r]);
        for (i64 i = 0; i < n; i++) lhs[r * n + i] = mul_mod(&m, lhs[r * n + i], rhs[r * n + i]);
    }
}
/* poly *= scalarResidues, PolyRq.swift:232-245 */
static void poly_mul_scalar_rows(i64 n, const u64 *moduli, int nmod, u64 *data, const u64 *scalars) {
    for (int r = 0; r < nmod; r++) {
        shoup_t s = shoup_make(scalars[r], moduli[r]);
        for (i64 i = 0; i < n; i++) data[r * n + i] = shoup_mul(&s, data[r * n + i]);
    }
}

/* divideAndRoundQLast, PolyRq.swift:365-393; inverseQLast from PolyContext.swift:108-111 */
int orc_divide_round_qlast(int64_t n, const uint64_t *moduli, int32_t nmod, uint64_t *data) {
    if (nmod < 2) return -1;
    const u64 qlast = moduli[nmod - 1];
    u64 *last = data + (i64)(nmod - 1) * n;
    const u64 half = qlast >> 1;
    for (i64 i = 0; i < n; i++) last[i] = add_mod(last[i], half, qlast);
    for (int r = 0; r < nmod - 1; r++) {
        modulus_t qi = modulus_make(moduli[r]);
        u64 inv = orc_inverse_mod(qlast % moduli[r], moduli[r]);
        if (!inv) return -2;
        shoup_t inv_qlast = shoup_make(inv, moduli[r]);
        const u64 half_mod_qi = reduce_single(&qi, half);
        u64 *row = data + (i64)r * n;
        for (i64 i = 0; i < n; i++) {
            u64 tmp = reduce_single(&qi, last[i]);
            u64 c = sub_mod(add_mod(row[i], half_mod_qi, qi.p), tmp, qi.p);
            row[i] = shoup_mul(&inv_qlast, c);
        }
    }
    return 0;
}

/* =====================================================================================
 * Fast base conversion -- Sources/HomomorphicEncryption/RnsBaseConverter.swift, CrtComposer.swift
 * ===================================================================================== */

typedef struct {
    int nin, nout;
    u64 in[ORC_MAX_MODULI], out[ORC_MAX_MODULI];
    modulus_t outmod[ORC_MAX_MODULI];
    shoup_t inv_punct[ORC_MAX_MODULI];         /* (q/q_i)^{-1} mod q_i, CrtComposer.swift:26-45 */
    u64 punct[ORC_MAX_MODULI][ORC_MAX_MODULI]; /* [j][i] = (q/q_i) mod t_j, RnsBaseConverter.swift:41-50 */
} baseconv_t;

static int baseconv_init(baseconv_t *bc, const u64 *in, int nin, const u64 *out, int nout) {
    bc->nin = nin;
    bc->nout = nout;
    for (int i = 0; i < nin; i++) bc->in[i] = in[i];
    for (int j = 0; j < nout; j++) { bc->out[j] = out[j]; bc->outmod[j] = modulus_make(out[j]); }
    for (int j = 0; j < nout; j++)
        for (int i = 0; i < nin; i++) {
            u64 prod = 1 % out[j];
            for (int k = 0; k < nin; k++)
                if (in[k] != in[i]) prod = reduce_double(&bc->outmod[j], (u128)prod * in[k]);
            bc->punct[j][i] = prod;
        }
    for (int i = 0; i < nin; i++) {
        modulus_t qi = modulus_make(in[i]);
        u64 prod = 1 % in[i];
        for (int k = 0; k < nin; k++)
            if (in[k] != in[i]) prod = reduce_double(&qi, (u128)prod * in[k]);
        u64 inv = orc_inverse_mod(prod, in[i]);
        if (!inv && in[i] != 1) return -1;
        bc->inv_punct[i] = shoup_make(inv, in[i]);
    }
    return 0;
}

/* convertApproximateProducts, RnsBaseConverter.swift:97-106 (in place on nin x n) */
static void baseconv_products(const baseconv_t *bc, i64 n, u64 *data) {
    for (int i = 0; i < bc->nin; i++)
        for (i64 c = 0; c < n; c++) data[i * n + c] = shoup_mul(&bc->inv_punct[i], data[i * n + c]);
}
/* convertApproximate(using:), RnsBaseConverter.swift:117-143 */
static void baseconv_from_products(const baseconv_t *bc, i64 n, const u64 *products, u64 *out) {
    u128 *sums = (u128 *)malloc(sizeof(u128) * n);
    for (int j = 0; j < bc->nout; j++) {
        memset(sums, 0, sizeof(u128) * n);
        for (int i = 0; i < bc->nin; i++) {
            const u64 pp = bc->punct[j][i];
            const u64 *row = products + (i64)i * n;
            for (i64 c = 0; c < n; c++) sums[c] += (u128)row[c] * pp;
        }
        for (i64 c = 0; c < n; c++) out[(i64)j * n + c] = reduce_double(&bc->outmod[j], sums[c]);
    }
    free(sums);
}
/* convertApproximate(poly:), RnsBaseConverter.swift:68-73 */
static void baseconv_convert(const baseconv_t *bc, i64 n, const u64 *in, u64 *out) {
    u64 *tmp = (u64 *)malloc(sizeof(u64) * n * bc->nin);
    memcpy(tmp, in, sizeof(u64) * n * bc->nin);
    baseconv_products(bc, n, tmp);
    baseconv_from_products(bc, n, tmp, out);
    free(tmp);
}

void orc_convert_approximate(int64_t n, const uint64_t *q, int32_t nq, const uint64_t *tmod, int32_t nt,
                             const uint64_t *in, uint64_t *out) {
    baseconv_t *bc = (baseconv_t *)malloc(sizeof(baseconv_t));
    baseconv_init(bc, q, nq, tmod, nt);
    baseconv_convert(bc, n, in, out);
    free(bc);
}

/* =====================================================================================
 * BEHZ RNS tool -- Sources/HomomorphicEncryption/RnsTool.swift
 * ===================================================================================== */

/* Word-size constants (ModularArithmetic/Scalar.swift:498-525): Bfv<UInt64> uses m~ = 2^32, gamma = 2^62 - 40797 and
 * 61-bit Bsk primes; Bfv<UInt32> uses m~ = 2^16, gamma = 2^30 - 20405 and 29-bit Bsk primes (bitWidth - 3).  The residues
 * are the same numbers whatever the storage width, so one restatement with these three constants covers both. */
#define ORC_MTILDE (rt->mtilde)
#define ORC_GAMMA (rt->gamma)

struct orc_rnstool {
    int word_bits;     /* 64 or 32 */
    u64 mtilde, gamma; /* T.mTilde, T.rnsCorrectionFactor */
    i64 n;
    int nq, nb; /* nb = nq + 1 = |Bsk| */
    u64 q[ORC_MAX_MODULI];
    modulus_t qmod[ORC_MAX_MODULI];
    u64 bsk[ORC_MAX_MODULI];
    u64 qbsk[2 * ORC_MAX_MODULI]; /* [Q, Bsk] */
    u64 t;
    modulus_t tmod;
    u64 m_tilde_mod_q[ORC_MAX_MODULI];     /* RnsTool.swift:226 */
    shoup_t neg_inv_q_mod_mtilde;          /* :159-165 */
    shoup_t q_mod_bsk[ORC_MAX_MODULI];     /* :217-220 */
    shoup_t inv_mtilde_mod_bsk[ORC_MAX_MODULI]; /* :221-224 */
    shoup_t inv_q_mod_bsk[ORC_MAX_MODULI]; /* :235-239 */
    shoup_t inv_b_mod_msk;                 /* :242-246 */
    shoup_t b_mod_q[ORC_MAX_MODULI], neg_b_mod_q[ORC_MAX_MODULI]; /* :200-215 */
    u64 prod_gamma_t_mod_q[ORC_MAX_MODULI]; /* :145-146 */
    shoup_t inv_gamma_mod_t;                /* :147-150 */
    u64 neg_inv_q_mod_tgamma[2];            /* :154-157 */
    u64 q_mod_t;                            /* :167 */
    shoup_t q_div_t[ORC_MAX_MODULI];        /* :176-182 */
    baseconv_t q_to_bsk, q_to_bsk_mtilde, b_to_msk, b_to_q, q_to_tgamma; /* :247-250,153 */
    const ntt_tables *qbsk_tables[2 * ORC_MAX_MODULI]; /* filled by orc_context_create (NTT-friendly Q only) */
};

static u64 q_remainder(const u64 *moduli, int n, u64 p) { /* PolyContext.qRemainder, PolyContext.swift:184-190 */
    u64 prod = 1 % p;
    for (int i = 0; i < n; i++) prod = (u64)(((u128)prod * moduli[i]) % p);
    return prod;
}

/* floor(Q / t) mod q_i with schoolbook multi-limb arithmetic (reference uses Width32<T>, RnsTool.swift:170-182) */
static u64 q_div_t_mod(const u64 *moduli, int n, u64 t, u64 qi) {
    u64 limbs[ORC_MAX_MODULI + 1];
    int nl = 1;
    limbs[0] = 1;
    for (int i = 0; i < n; i++) {
        u64 carry = 0;
        for (int k = 0; k < nl; k++) {
            u128 v = (u128)limbs[k] * moduli[i] + carry;
            limbs[k] = (u64)v;
            carry = (u64)(v >> 64);
        }
        if (carry) limbs[nl++] = carry;
    }
    u64 rem = 0; /* long division by t, most significant limb first */
    for (int k = nl - 1; k >= 0; k--) {
        u128 v = (((u128)rem) << 64) | limbs[k];
        limbs[k] = (u64)(v / t);
        rem = (u64)(v % t);
    }
    u64 r = 0;
    for (int k = nl - 1; k >= 0; k--) r = (u64)(((((u128)r) << 64) | limbs[k]) % qi);
    return r;
}

orc_rnstool *orc_rnstool_create_w(int64_t n, const uint64_t *q, int32_t nq, uint64_t t, int32_t word_bits) {
    if (nq < 1 || nq + 2 > ORC_MAX_MODULI || (word_bits != 64 && word_bits != 32)) return NULL;
    for (int i = 0; i < nq; i++)
        if (word_bits == 32 && q[i] >= ((u64)1 << 30)) return NULL; /* Modulus<UInt32>.max, Modulus.swift:177-180 */
    orc_rnstool *rt = (orc_rnstool *)calloc(1, sizeof(orc_rnstool));
    rt->word_bits = word_bits;
    rt->mtilde = word_bits == 64 ? ((u64)1 << 32) : ((u64)1 << 16);
    rt->gamma = word_bits == 64 ? ((((u64)1) << 62) - 40797) : ((((u64)1) << 30) - 20405);
    rt->n = n;
    rt->nq = nq;
    rt->nb = nq + 1;
    rt->t = t;
    rt->tmod = modulus_make(t);
    for (int i = 0; i < nq; i++) { rt->q[i] = q[i]; rt->qmod[i] = modulus_make(q[i]); }
    /* Bsk primes: RnsTool.swift:30-33 -- (bitWidth-3)-bit, ascending, NTT-friendly for degree n */
    int32_t bits[ORC_MAX_MODULI];
    for (int i = 0; i < rt->nb; i++) bits[i] = word_bits - 3;
    if (orc_generate_primes(bits, rt->nb, 1, n, rt->bsk) != rt->nb) { free(rt); return NULL; }
    for (int i = 0; i < nq; i++) rt->qbsk[i] = q[i];
    for (int j = 0; j < rt->nb; j++) rt->qbsk[nq + j] = rt->bsk[j];
    const u64 msk = rt->bsk[rt->nb - 1];
    const int nB = rt->nb - 1; /* base B = Bsk without m_sk */

    u128 gamma_t = (u128)ORC_GAMMA * t;
    for (int i = 0; i < nq; i++) rt->prod_gamma_t_mod_q[i] = reduce_double(&rt->qmod[i], gamma_t);
    rt->inv_gamma_mod_t = shoup_make(orc_inverse_mod(ORC_GAMMA, t), t);
    u64 tgamma[2] = {t, ORC_GAMMA};
    baseconv_init(&rt->q_to_tgamma, q, nq, tgamma, 2);
    for (int k = 0; k < 2; k++) {
        u64 qm = q_remainder(q, nq, tgamma[k]);
        rt->neg_inv_q_mod_tgamma[k] = neg_mod(orc_inverse_mod(qm, tgamma[k]), tgamma[k]);
    }
    {
        u64 qm = q_remainder(q, nq, ORC_MTILDE);
        u64 v = neg_mod(orc_inverse_mod(qm, ORC_MTILDE), ORC_MTILDE);
        rt->neg_inv_q_mod_mtilde = shoup_make(v, ORC_MTILDE);
    }
    rt->q_mod_t = q_remainder(q, nq, t);
    for (int i = 0; i < nq; i++) rt->q_div_t[i] = shoup_make(q_div_t_mod(q, nq, t, q[i]), q[i]);
    for (int i = 0; i < nq; i++) {
        u64 b = q_remainder(rt->bsk, nB, q[i]);
        rt->b_mod_q[i] = shoup_make(b, q[i]);
        rt->neg_b_mod_q[i] = shoup_make(neg_mod(b, q[i]), q[i]);
        rt->m_tilde_mod_q[i] = reduce_single(&rt->qmod[i], ORC_MTILDE);
    }
    for (int j = 0; j < rt->nb; j++) {
        u64 bj = rt->bsk[j];
        u64 qm = q_remainder(q, nq, bj);
        rt->q_mod_bsk[j] = shoup_make(qm, bj);
        rt->inv_mtilde_mod_bsk[j] = shoup_make(orc_inverse_mod(ORC_MTILDE, bj), bj);
        rt->inv_q_mod_bsk[j] = shoup_make(orc_inverse_mod(qm, bj), bj);
    }
    rt->inv_b_mod_msk = shoup_make(orc_inverse_mod(q_remainder(rt->bsk, nB, msk), msk), msk);
    u64 bsk_mtilde[ORC_MAX_MODULI];
    for (int j = 0; j < rt->nb; j++) bsk_mtilde[j] = rt->bsk[j];
    bsk_mtilde[rt->nb] = ORC_MTILDE;
    baseconv_init(&rt->q_to_bsk, q, nq, rt->bsk, rt->nb);
    baseconv_init(&rt->q_to_bsk_mtilde, q, nq, bsk_mtilde, rt->nb + 1);
    baseconv_init(&rt->b_to_msk, rt->bsk, nB, &msk, 1);
    baseconv_init(&rt->b_to_q, rt->bsk, nB, q, nq);
    return rt;
}
orc_rnstool *orc_rnstool_create(int64_t n, const uint64_t *q, int32_t nq, uint64_t t) {
    return orc_rnstool_create_w(n, q, nq, t, 64);
}
void orc_rnstool_destroy(orc_rnstool *rt) { free(rt); }
int32_t orc_rnstool_bsk(const orc_rnstool *rt, uint64_t *out) {
    for (int j = 0; j < rt->nb; j++) out[j] = rt->bsk[j];
    return rt->nb;
}

/* convertApproximateBskMTilde, RnsTool.swift:313-316 */
void orc_rnstool_convert_bsk_mtilde(const orc_rnstool *rt, const uint64_t *in, uint64_t *out) {
    const i64 n = rt->n;
    u64 *scaled = (u64 *)malloc(sizeof(u64) * n * rt->nq);
    memcpy(scaled, in, sizeof(u64) * n * rt->nq);
    poly_mul_scalar_rows(n, rt->q, rt->nq, scaled, rt->m_tilde_mod_q);
    baseconv_convert(&rt->q_to_bsk_mtilde, n, scaled, out);
    free(scaled);
}

/* smallMontgomeryReduce, RnsTool.swift:339-368; data = (nb+1) x n in base [Bsk, m~], result in first nb rows */
void orc_rnstool_small_montgomery_reduce(const orc_rnstool *rt, uint64_t *data) {
    const i64 n = rt->n;
    const u64 threshold = ORC_MTILDE >> 1;
    u64 *mrow = data + (i64)rt->nb * n;
    for (i64 c = 0; c < n; c++) mrow[c] = shoup_mul(&rt->neg_inv_q_mod_mtilde, mrow[c]);
    for (int j = 0; j < rt->nb; j++) {
        const u64 bj = rt->bsk[j];
        u64 *row = data + (i64)j * n;
        for (i64 c = 0; c < n; c++) {
            u64 r = mrow[c];
            if (!(r < threshold)) r = r + bj - ORC_MTILDE;
            u64 v = row[c] + shoup_lazy(&rt->q_mod_bsk[j], r);
            row[c] = shoup_mul(&rt->inv_mtilde_mod_bsk[j], v);
        }
    }
}

/* liftQToQBsk, RnsTool.swift:324-331 */
void orc_rnstool_lift(const orc_rnstool *rt, const uint64_t *in, uint64_t *out) {
    const i64 n = rt->n;
    u64 *tmp = (u64 *)malloc(sizeof(u64) * n * (rt->nb + 1));
    orc_rnstool_convert_bsk_mtilde(rt, in, tmp);
    orc_rnstool_small_montgomery_reduce(rt, tmp);
    memcpy(out, in, sizeof(u64) * n * rt->nq);
    memcpy(out + (i64)rt->nq * n, tmp, sizeof(u64) * n * rt->nb);
    free(tmp);
}

/* approximateFloor, RnsTool.swift:378-398 */
void orc_rnstool_approximate_floor(const orc_rnstool *rt, const uint64_t *in, uint64_t *out) {
    const i64 n = rt->n;
    baseconv_convert(&rt->q_to_bsk, n, in, out); /* converts the Q rows */
    const u64 *in_bsk = in + (i64)rt->nq * n;
    for (int j = 0; j < rt->nb; j++) {
        const u64 bj = rt->bsk[j];
        for (i64 c = 0; c < n; c++) {
            i64 idx = (i64)j * n + c;
            out[idx] = shoup_mul(&rt->inv_q_mod_bsk[j], in_bsk[idx] + bj - out[idx]);
        }
    }
}

/* convertApproximateBskToQ, RnsTool.swift:402-450 */
void orc_rnstool_bsk_to_q(const orc_rnstool *rt, const uint64_t *in, uint64_t *out) {
    const i64 n = rt->n;
    const int nB = rt->nb - 1;
    const u64 msk = rt->bsk[rt->nb - 1];
    const u64 *in_msk = in + (i64)nB * n;
    u64 *poly_b = (u64 *)malloc(sizeof(u64) * n * nB);
    memcpy(poly_b, in, sizeof(u64) * n * nB);
    baseconv_products(&rt->b_to_msk, n, poly_b); /* same products serve B->msk and B->Q */
    u64 *alpha = (u64 *)malloc(sizeof(u64) * n);
    baseconv_from_products(&rt->b_to_msk, n, poly_b, alpha);
    const u64 threshold = msk >> 1;
    for (i64 c = 0; c < n; c++) alpha[c] = shoup_mul(&rt->inv_b_mod_msk, alpha[c] + msk - in_msk[c]);
    baseconv_from_products(&rt->b_to_q, n, poly_b, out);
    for (int i = 0; i < rt->nq; i++) {
        const u64 qi = rt->q[i];
        for (i64 c = 0; c < n; c++) {
            u64 a = alpha[c];
            u64 adjust = (a > threshold) ? shoup_mul(&rt->b_mod_q[i], msk - a) : shoup_mul(&rt->neg_b_mod_q[i], a);
            out[(i64)i * n + c] = add_mod(out[(i64)i * n + c], adjust, qi);
        }
    }
    free(alpha);
    free(poly_b);
}

/* floorQBskToQ, RnsTool.swift:453-456 */
void orc_rnstool_floor(const orc_rnstool *rt, const uint64_t *in, uint64_t *out) {
    u64 *floored = (u64 *)malloc(sizeof(u64) * rt->n * rt->nb);
    orc_rnstool_approximate_floor(rt, in, floored);
    orc_rnstool_bsk_to_q(rt, floored, out);
    free(floored);
}

/* scaleAndRound, RnsTool.swift:272-302 */
void orc_rnstool_scale_and_round(const orc_rnstool *rt, const uint64_t *in, uint64_t scaling_factor, uint64_t *out) {
    const i64 n = rt->n;
    const u64 t = rt->t;
    u64 *poly = (u64 *)malloc(sizeof(u64) * n * rt->nq);
    memcpy(poly, in, sizeof(u64) * n * rt->nq);
    poly_mul_scalar_rows(n, rt->q, rt->nq, poly, rt->prod_gamma_t_mod_q);
    u64 *tg = (u64 *)malloc(sizeof(u64) * n * 2);
    baseconv_convert(&rt->q_to_tgamma, n, poly, tg);
    u64 tgamma[2] = {t, ORC_GAMMA};
    poly_mul_scalar_rows(n, tgamma, 2, tg, rt->neg_inv_q_mod_tgamma);
    const u64 corrected_gamma = ORC_GAMMA / 2;
    for (i64 c = 0; c < n; c++) {
        u64 mod_t = tg[c], mod_g = tg[n + c];
        u64 s_greater = neg_mod(reduce_single(&rt->tmod, ORC_GAMMA - mod_g), t);
        u64 s_less = reduce_single(&rt->tmod, mod_g);
        u64 s = (mod_g > corrected_gamma) ? s_greater : s_less;
        out[c] = sub_mod(mod_t, s, t);
    }
    u64 scaled = shoup_mul(&rt->inv_gamma_mod_t, scaling_factor);
    poly_mul_scalar_rows(n, &t, 1, out, &scaled);
    free(tg);
    free(poly);
}

/* =====================================================================================
 * BFV context and scheme ops -- Context.swift, Bfv/ *.swift
 * ===================================================================================== */

struct orc_context {
    i64 n;
    int L; /* ciphertext moduli */
    u64 q[ORC_MAX_MODULI]; /* q_0..q_{L-1}, then q_ks at index L */
    u64 t;
    orc_rnstool *tools[ORC_MAX_MODULI]; /* tools[l] for l = 1..L input moduli (Context.swift:129-141) */
    const ntt_tables *qt[ORC_MAX_MODULI]; /* NTT tables of q_0..q_{L-1}, q_ks */
};

orc_context *orc_context_create(int64_t n, const uint64_t *coeff_moduli, int32_t nmod, uint64_t t) {
    return orc_context_create_w(n, coeff_moduli, nmod, t, 64);
}
/* word_bits = 32: Context<Bfv<UInt32>> (same residues, Bfv<UInt32>'s m~ / gamma / Bsk) */
orc_context *orc_context_create_w(int64_t n, const uint64_t *coeff_moduli, int32_t nmod, uint64_t t, int32_t word_bits) {
    if (nmod < 2 || nmod > ORC_MAX_MODULI / 2 - 2) return NULL;
    /* keep the per-call scratch on the heap (no mmap/munmap + page faults per multiply): a fair CPU baseline */
    mallopt(M_MMAP_THRESHOLD, 1 << 30);
    mallopt(M_TRIM_THRESHOLD, 1 << 30);
    mallopt(M_ARENA_MAX, 256);
    orc_context *ctx = (orc_context *)calloc(1, sizeof(orc_context));
    ctx->n = n;
    ctx->L = nmod - 1;
    ctx->t = t;
    for (int i = 0; i < nmod; i++) {
        ctx->q[i] = coeff_moduli[i];
        ctx->qt[i] = ntt_tables_get(n, coeff_moduli[i]);
        if (!ctx->qt[i]) { free(ctx); return NULL; }
    }
    for (int l = 1; l <= ctx->L; l++) {
        /* NOTE: below the top level the reference carves its Bsk/m~ base out of the top-level one
         * (RnsTool.swift:185-186); only scaleAndRound -- which does not touch that base -- is used from
         * tools[l < L] here.  ct x ct multiply is top level only ("levels"). */
        ctx->tools[l] = orc_rnstool_create_w(n, ctx->q, l, t, word_bits);
        if (!ctx->tools[l]) { free(ctx); return NULL; }
    }
    orc_rnstool *top = ctx->tools[ctx->L];
    for (int r = 0; r < top->nq + top->nb; r++) {
        top->qbsk_tables[r] = ntt_tables_get(n, top->qbsk[r]);
        if (!top->qbsk_tables[r]) return NULL;
    }
    return ctx;
}
void orc_context_destroy(orc_context *ctx) {
    if (!ctx) return;
    for (int l = 1; l <= ctx->L; l++) orc_rnstool_destroy(ctx->tools[l]);
    free(ctx);
}
int32_t orc_context_L(const orc_context *ctx) { return ctx->L; }
int32_t orc_context_bsk(const orc_context *ctx, uint64_t *out) { return orc_rnstool_bsk(ctx->tools[ctx->L], out); }

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* computeBehzPolys for one poly, Bfv+Multiply.swift:51-57 */
static void behz_poly(const orc_context *ctx, const u64 *in, u64 *out) {
    const orc_rnstool *rt = ctx->tools[ctx->L];
    const int R = rt->nq + rt->nb;
    orc_rnstool_lift(rt, in, out);
    for (int r = 0; r < R; r++) ntt_forward_row(rt->qbsk_tables[r], out + (i64)r * ctx->n);
}

int orc_bfv_lift_ntt(const orc_context *ctx, const uint64_t *polys, uint64_t *out, int64_t npolys) {
    const int L = ctx->L, R = 2 * L + 1;
    for (i64 k = 0; k < npolys; k++) behz_poly(ctx, polys + k * L * ctx->n, out + k * R * ctx->n);
    return 0;
}

/* mulAssign = multiplyWithoutScaling + dropExtendedBase, Bfv+Multiply.swift:18-85 */
static void bfv_mul_one(const orc_context *ctx, const u64 *a, const u64 *b, u64 *out) {
    const i64 n = ctx->n;
    const int L = ctx->L, R = 2 * L + 1;
    const orc_rnstool *rt = ctx->tools[L];
    const i64 psz = (i64)R * n;
    u64 *buf = (u64 *)malloc(sizeof(u64) * psz * 7);
    u64 *l0 = buf, *l1 = buf + psz, *r0 = buf + 2 * psz, *r1 = buf + 3 * psz;
    u64 *p0 = buf + 4 * psz, *p1 = buf + 5 * psz, *p2 = buf + 6 * psz;
    behz_poly(ctx, a, l0);
    behz_poly(ctx, a + (i64)L * n, l1);
    behz_poly(ctx, b, r0);
    behz_poly(ctx, b + (i64)L * n, r1);
    for (int r = 0; r < R; r++) { /* :80-82 */
        modulus_t m = modulus_make(rt->qbsk[r]);
        for (i64 c = 0; c < n; c++) {
            i64 i = (i64)r * n + c;
            p0[i] = mul_mod(&m, l0[i], r0[i]);
            p1[i] = add_mod(mul_mod(&m, l0[i], r1[i]), mul_mod(&m, l1[i], r0[i]), m.p);
            p2[i] = mul_mod(&m, l1[i], r1[i]);
        }
    }
    u64 tvec[2 * ORC_MAX_MODULI];
    for (int r = 0; r < R; r++) tvec[r] = ctx->t;
    u64 *polys[3] = {p0, p1, p2};
    for (int k = 0; k < 3; k++) { /* dropExtendedBase :31-48 */
        poly_mul_scalar_rows(n, rt->qbsk, R, polys[k], tvec);
        for (int r = 0; r < R; r++) ntt_inverse_row(rt->qbsk_tables[r], polys[k] + (i64)r * n);
        orc_rnstool_floor(rt, polys[k], out + (i64)k * L * n);
    }
    free(buf);
}

int orc_bfv_mul(const orc_context *ctx, const uint64_t *a, const uint64_t *b, uint64_t *out, int64_t batch,
                int32_t threads) {
    const i64 in_sz = (i64)2 * ctx->L * ctx->n, out_sz = (i64)3 * ctx->L * ctx->n;
    if (threads <= 0) threads = orc_num_threads();
#pragma omp parallel for num_threads(threads) schedule(dynamic, 1)
    for (i64 k = 0; k < batch; k++) bfv_mul_one(ctx, a + k * in_sz, b + k * in_sz, out + k * out_sz);
    return 0;
}

/* ---- deterministic PRNG for synthetic inputs / keys (NOT the reference's AES-CTR-DRBG; out of scope) ---- */
static inline u64 splitmix64(u64 *s) {
    u64 z = (*s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
void orc_fill_uniform(uint64_t seed, const uint64_t *moduli, int32_t nmod, int64_t n, uint64_t *data, int64_t rows) {
    /* like randomizeUniform (PolyRq+Randomize.swift:57-78): a 128-bit draw reduced mod q */
    for (i64 r = 0; r < rows; r++) {
        u64 s = seed * 0x100000001B3ull + (u64)r * 0x9E3779B97F4A7C15ull + 1;
        const u64 p = moduli[r % nmod];
        for (i64 c = 0; c < n; c++) {
            u128 v = (((u128)splitmix64(&s)) << 64) | splitmix64(&s);
            data[r * n + c] = (u64)(v % p);
        }
    }
}
/* randomizeTernary, PolyRq+Randomize.swift:91-106 */
static void sample_ternary(u64 *s, const u64 *moduli, int nmod, i64 n, u64 *data) {
    for (i64 c = 0; c < n; c++) {
        u128 v = (((u128)splitmix64(s)) << 32) | (splitmix64(s) & 0xFFFFFFFFull);
        u64 val = (u64)(v % 3);
        for (int r = 0; r < nmod; r++) data[(i64)r * n + c] = sub_mod(val, 1, moduli[r]);
    }
}
/* randomizeCenteredBinomialDistribution with sigma = 3.2 => k = 21, PolyRq+Randomize.swift:121-167 */
static void sample_cbd(u64 *s, const u64 *moduli, int nmod, i64 n, u64 *data) {
    const u64 mask = (((u64)1) << 21) - 1;
    for (i64 c = 0; c < n; c++) {
        u64 pos = (u64)__builtin_popcountll(splitmix64(s) & mask);
        u64 neg = (u64)__builtin_popcountll(splitmix64(s) & mask);
        for (int r = 0; r < nmod; r++) data[(i64)r * n + c] = sub_mod(pos, neg, moduli[r]);
    }
}

/* encryptZero, Bfv+Encrypt.swift:150-181: ct = (-(a*s + e), a), Coeff format, nmod rows from `moduli` */
static void encrypt_zero(const orc_context *ctx, u64 *rng, const u64 *moduli, int nmod, const u64 *sk_rows[],
                         u64 *c0, u64 *c1) {
    const i64 n = ctx->n;
    u64 *a = c1; /* sampled directly in Eval form */
    for (int r = 0; r < nmod; r++) {
        for (i64 c = 0; c < n; c++) {
            u128 v = (((u128)splitmix64(rng)) << 64) | splitmix64(rng);
            a[(i64)r * n + c] = (u64)(v % moduli[r]);
        }
    }
    u64 *err = (u64 *)malloc(sizeof(u64) * n * nmod);
    sample_cbd(rng, moduli, nmod, n, err);
    for (int r = 0; r < nmod; r++) {
        modulus_t m = modulus_make(moduli[r]);
        const ntt_tables *tb = ntt_tables_get(n, moduli[r]);
        u64 *row0 = c0 + (i64)r * n, *rowa = a + (i64)r * n;
        for (i64 c = 0; c < n; c++) row0[c] = mul_mod(&m, rowa[c], sk_rows[r][c]);
        ntt_inverse_row(tb, row0);
        for (i64 c = 0; c < n; c++) row0[c] = neg_mod(add_mod(row0[c], err[(i64)r * n + c], m.p), m.p);
        ntt_inverse_row(tb, rowa);
    }
    free(err);
}

/* generateSecretKey, Bfv+Keys.swift:20-26 -- ternary secret over all coefficient moduli, Eval format */
static void gen_secret(const orc_context *ctx, u64 *rng, u64 *sk) {
    const int K = ctx->L + 1;
    sample_ternary(rng, ctx->q, K, ctx->n, sk);
    for (int r = 0; r < K; r++) ntt_forward_row(ntt_tables_get(ctx->n, ctx->q[r]), sk + (i64)r * ctx->n);
}

/* _generateKeySwitchKey, Bfv+Keys.swift:69-103 */
int orc_gen_keyswitch_key(const orc_context *ctx, uint64_t seed, const uint64_t *sk, const uint64_t *cur,
                          uint64_t *ksk) {
    const i64 n = ctx->n;
    const int L = ctx->L, K = L + 1;
    u64 rng = seed ^ 0xA5A5A5A5DEADBEEFull;
    const u64 key_modulus = ctx->q[L];
    const u64 *sk_rows[ORC_MAX_MODULI];
    for (int r = 0; r < K; r++) sk_rows[r] = sk + (i64)r * n;
    for (int j = 0; j < L; j++) {
        u64 *c0 = ksk + ((i64)j * 2 + 0) * K * n, *c1 = ksk + ((i64)j * 2 + 1) * K * n;
        encrypt_zero(ctx, &rng, ctx->q, K, sk_rows, c0, c1);
        for (int r = 0; r < K; r++) {
            const ntt_tables *tb = ntt_tables_get(n, ctx->q[r]);
            ntt_forward_row(tb, c0 + (i64)r * n);
            ntt_forward_row(tb, c1 + (i64)r * n);
        }
        modulus_t m = modulus_make(ctx->q[j]);
        shoup_t prod = shoup_make(reduce_single(&m, key_modulus), ctx->q[j]);
        for (i64 c = 0; c < n; c++) {
            u64 v = shoup_mul(&prod, cur[(i64)j * n + c]);
            c0[(i64)j * n + c] = add_mod(c0[(i64)j * n + c], v, m.p);
        }
    }
    return 0;
}

/* generateSecretKey + generateRelinearizationKey, Bfv+Keys.swift:20-26,58-66 */
int orc_keygen(const orc_context *ctx, uint64_t seed, uint64_t *sk, uint64_t *relin_key) {
    const i64 n = ctx->n;
    const int K = ctx->L + 1;
    u64 rng = seed;
    gen_secret(ctx, &rng, sk);
    if (!relin_key) return 0;
    u64 *s2 = (u64 *)malloc(sizeof(u64) * n * K);
    for (int r = 0; r < K; r++) {
        modulus_t m = modulus_make(ctx->q[r]);
        for (i64 c = 0; c < n; c++) s2[(i64)r * n + c] = mul_mod(&m, sk[(i64)r * n + c], sk[(i64)r * n + c]);
    }
    int rc = orc_gen_keyswitch_key(ctx, seed + 1, sk, s2, relin_key);
    free(s2);
    return rc;
}

/* _computeKeySwitchingUpdate, Bfv+Keys.swift:123-208 */
int orc_keyswitch_update(const orc_context *ctx, const uint64_t *target, int32_t l, const uint64_t *ksk,
                         uint64_t *out) {
    const i64 n = ctx->n;
    const int L = ctx->L, K = L + 1;
    if (l < 1 || l > L) return -1;
    const int rns_count = l + 1;
    u64 ksm[ORC_MAX_MODULI]; /* keySwitchingContexts[l-1].moduli = q_0..q_{l-1}, q_ks (Context.swift:114-127) */
    for (int i = 0; i < l; i++) ksm[i] = ctx->q[i];
    ksm[l] = ctx->q[L];
    u64 *prod = (u64 *)malloc(sizeof(u64) * 2 * rns_count * n);
    u128 *acc = (u128 *)malloc(sizeof(u128) * 2 * n);
    u64 *buf = (u64 *)malloc(sizeof(u64) * n);
    for (int r = 0; r < rns_count; r++) {
        const int key_index = (r == rns_count - 1) ? K - 1 : r; /* :153 */
        modulus_t km = modulus_make(ksm[r]);
        const ntt_tables *tb = ctx->qt[(r == l) ? L : r];
        memset(acc, 0, sizeof(u128) * 2 * n);
        for (int j = 0; j < l; j++) {
            memcpy(buf, target + (i64)j * n, sizeof(u64) * n);
            if (ksm[j] > km.p) /* :168-172 */
                for (i64 c = 0; c < n; c++) buf[c] = reduce_single(&km, buf[c]);
            ntt_forward_row(tb, buf); /* :174-179 */
            for (int comp = 0; comp < 2; comp++) { /* :180-191 */
                const u64 *krow = ksk + (((i64)j * 2 + comp) * K + key_index) * n;
                u128 *a = acc + (i64)comp * n;
                for (i64 c = 0; c < n; c++) a[c] += (u128)buf[c] * krow[c];
            }
        }
        for (int comp = 0; comp < 2; comp++) /* :193-202 */
            for (i64 c = 0; c < n; c++)
                prod[((i64)comp * rns_count + r) * n + c] = reduce_double(&km, acc[(i64)comp * n + c]);
    }
    for (int comp = 0; comp < 2; comp++) { /* :204-206 */
        u64 *p = prod + (i64)comp * rns_count * n;
        for (int r = 0; r < rns_count; r++) ntt_inverse_row(ctx->qt[(r == l) ? L : r], p + (i64)r * n);
        orc_divide_round_qlast(n, ksm, rns_count, p);
        memcpy(out + (i64)comp * l * n, p, sizeof(u64) * l * n);
    }
    free(buf);
    free(acc);
    free(prod);
    return 0;
}

/* relinearize, Bfv.swift:201-219 */
int orc_bfv_relinearize(const orc_context *ctx, const uint64_t *ct3, int32_t l, const uint64_t *relin_key,
                        uint64_t *out, int64_t batch, int32_t threads) {
    const i64 n = ctx->n;
    const i64 psz = (i64)l * n;
    if (threads <= 0) threads = orc_num_threads();
    int rc = 0;
#pragma omp parallel for num_threads(threads) schedule(dynamic, 1)
    for (i64 k = 0; k < batch; k++) {
        const u64 *ct = ct3 + k * 3 * psz;
        u64 *o = out + k * 2 * psz;
        u64 *upd = (u64 *)malloc(sizeof(u64) * 2 * psz);
        if (orc_keyswitch_update(ctx, ct + 2 * psz, l, relin_key, upd) != 0) rc = -1;
        memcpy(o, ct, sizeof(u64) * 2 * psz);
        orc_poly_add(n, ctx->q, l, o, upd);
        orc_poly_add(n, ctx->q, l, o + psz, upd + psz);
        free(upd);
    }
    return rc;
}

/* modSwitchDown, Bfv.swift:163-171 */
int orc_bfv_mod_switch_down(const orc_context *ctx, const uint64_t *ct, int32_t npoly, int32_t l, uint64_t *out,
                            int64_t batch, int32_t threads) {
    const i64 n = ctx->n;
    if (l < 2 || l > ctx->L) return -1;
    if (threads <= 0) threads = orc_num_threads();
    const i64 polys = batch * npoly;
#pragma omp parallel for num_threads(threads) schedule(dynamic, 4)
    for (i64 k = 0; k < polys; k++) {
        u64 *tmp = (u64 *)malloc(sizeof(u64) * l * n);
        memcpy(tmp, ct + k * l * n, sizeof(u64) * l * n);
        orc_divide_round_qlast(n, ctx->q, l, tmp);
        memcpy(out + k * (l - 1) * n, tmp, sizeof(u64) * (l - 1) * n);
        free(tmp);
    }
    return 0;
}

/* encrypt = encryptZero + plaintextTranslate(Add), Bfv+Encrypt.swift:64-139 */
int orc_encrypt(const orc_context *ctx, uint64_t seed, const uint64_t *sk, const uint64_t *plain, uint64_t *ct) {
    const i64 n = ctx->n;
    const int L = ctx->L;
    const orc_rnstool *rt = ctx->tools[L];
    u64 rng = seed ^ 0x0123456789ABCDEFull;
    const u64 *sk_rows[ORC_MAX_MODULI];
    for (int r = 0; r < L; r++) sk_rows[r] = sk + (i64)r * n;
    u64 *c0 = ct, *c1 = ct + (i64)L * n;
    encrypt_zero(ctx, &rng, ctx->q, L, sk_rows, c0, c1);
    const u64 t = ctx->t;
    const u64 t_threshold = (t + 1) / 2; /* RnsTool.swift:123-125 */
    for (i64 c = 0; c < n; c++) {
        u64 adjust = (u64)((((u128)rt->q_mod_t * plain[c]) + t_threshold) / t); /* :87-105 */
        for (int r = 0; r < L; r++) {
            u64 delta = shoup_mul(&rt->q_div_t[r], plain[c]);
            u64 v = add_mod(delta, adjust, ctx->q[r]);
            c0[(i64)r * n + c] = add_mod(c0[(i64)r * n + c], v, ctx->q[r]);
        }
    }
    return 0;
}

/* decryptCoeff = forwardNtt + dotProduct + scaleAndRound, Bfv+Decrypt.swift:21-41,188-204 */
int orc_decrypt(const orc_context *ctx, const uint64_t *sk, const uint64_t *ct, int32_t npoly, int32_t l,
                uint64_t *plain) {
    const i64 n = ctx->n;
    if (l < 1 || l > ctx->L || npoly < 2) return -1;
    u64 *dot = (u64 *)malloc(sizeof(u64) * l * n);
    u64 *ci = (u64 *)malloc(sizeof(u64) * n);
    u64 *spow = (u64 *)malloc(sizeof(u64) * n);
    for (int r = 0; r < l; r++) {
        modulus_t m = modulus_make(ctx->q[r]);
        const ntt_tables *tb = ntt_tables_get(n, ctx->q[r]);
        u64 *drow = dot + (i64)r * n;
        memcpy(drow, ct + (i64)r * n, sizeof(u64) * n);
        ntt_forward_row(tb, drow);
        memcpy(spow, sk + (i64)r * n, sizeof(u64) * n);
        for (int k = 1; k < npoly; k++) {
            memcpy(ci, ct + ((i64)k * l + r) * n, sizeof(u64) * n);
            ntt_forward_row(tb, ci);
            for (i64 c = 0; c < n; c++) drow[c] = add_mod(drow[c], mul_mod(&m, ci[c], spow[c]), m.p);
            if (k != npoly - 1)
                for (i64 c = 0; c < n; c++) spow[c] = mul_mod(&m, spow[c], sk[(i64)r * n + c]);
        }
        ntt_inverse_row(tb, drow);
    }
    orc_rnstool_scale_and_round(ctx->tools[l], dot, 1, plain);
    free(spow);
    free(ci);
    free(dot);
    return 0;
}

/* =====================================================================================
 * Galois automorphisms -- PolyRq/Galois.swift, Bfv/Bfv.swift:174-198, Bfv+Keys.swift:42-49
 * ===================================================================================== */

/* PolyRq<Coeff>.applyGalois, Galois.swift:115-141 with GaloisCoeffIterator :18-60 */
void orc_galois_coeff(int64_t n, const uint64_t *moduli, int32_t nmod, int64_t element, const uint64_t *in, uint64_t *out) {
    const int logn = ilog2((u64)n);
    for (int r = 0; r < nmod; r++) {
        i64 raw = 0;
        for (i64 i = 0; i < n; i++) {
            const int negate = (raw >> logn) & 1;
            const i64 oi = raw & (n - 1);
            const u64 v = in[(i64)r * n + i];
            out[(i64)r * n + oi] = negate ? neg_mod(v, moduli[r]) : v;
            raw += element;
        }
    }
}
/* PolyRq<Eval>.applyGalois, Galois.swift:151-166 with GaloisEvalIterator :62-98 */
void orc_galois_eval(int64_t n, int32_t nmod, int64_t element, const uint64_t *in, uint64_t *out) {
    const int logn = ilog2((u64)n);
    for (i64 i = 0; i < n; i++) {
        const u64 reversed = orc_reverse_bits((uint32_t)(i + n), logn + 1);
        u64 raw = ((u64)element * reversed) >> 1;
        raw &= (u64)(n - 1);
        const i64 src = logn ? (i64)orc_reverse_bits((uint32_t)raw, logn) : 0;
        for (int r = 0; r < nmod; r++) out[(i64)r * n + i] = in[(i64)r * n + src];
    }
}
/* GaloisElement.rotatingColumns / swappingRows, Galois.swift:181-211 */
int64_t orc_galois_element_swapping_rows(int64_t degree) { return (degree << 1) - 1; }
int64_t orc_galois_element_rotating_columns(int64_t step, int64_t degree) {
    i64 pos = step < 0 ? -step : step;
    if (pos >= (degree >> 1) || pos <= 0) return 0;
    if (step > 0) pos = (degree >> 1) - pos;
    return (int64_t)orc_pow_mod(3, (u64)pos, (u64)degree << 1);
}
/* generateEvaluationKey's Galois branch, Bfv+Keys.swift:42-49 */
int orc_gen_galois_key(const orc_context *ctx, uint64_t seed, const uint64_t *sk, int64_t element, uint64_t *ksk) {
    const int K = ctx->L + 1;
    u64 *switched = (u64 *)malloc(sizeof(u64) * K * ctx->n);
    orc_galois_eval(ctx->n, K, element, sk, switched);
    int rc = orc_gen_keyswitch_key(ctx, seed, sk, switched, ksk);
    free(switched);
    return rc;
}
/* Bfv.applyGalois, Bfv.swift:174-198 */
int orc_bfv_apply_galois(const orc_context *ctx, const uint64_t *ct, int32_t l, int64_t element, const uint64_t *gkey,
                         uint64_t *out, int64_t batch, int32_t threads) {
    const i64 n = ctx->n, psz = (i64)l * n;
    if (threads <= 0) threads = orc_num_threads();
    int rc = 0;
#pragma omp parallel for num_threads(threads) schedule(dynamic, 1)
    for (i64 k = 0; k < batch; k++) {
        const u64 *c = ct + k * 2 * psz;
        u64 *o = out + k * 2 * psz;
        u64 *tmp = (u64 *)malloc(sizeof(u64) * psz);
        u64 *upd = (u64 *)malloc(sizeof(u64) * 2 * psz);
        orc_galois_coeff(n, ctx->q, l, element, c, o);
        orc_galois_coeff(n, ctx->q, l, element, c + psz, tmp);
        if (orc_keyswitch_update(ctx, tmp, l, gkey, upd) != 0) rc = -1;
        orc_poly_add(n, ctx->q, l, o, upd);
        memcpy(o + psz, upd + psz, sizeof(u64) * psz);
        free(upd);
        free(tmp);
    }
    return rc;
}

/* =====================================================================================
 * Lazy ciphertext-plaintext inner product -- Bfv/Bfv.swift:402-505, Plaintext.swift:149-171
 * ===================================================================================== */

/* Plaintext.convertToEvalFormat(moduliCount:), Plaintext.swift:149-171: centered lift of the t-residues to each
 * q_i (values >= (t+1)/2 get + (q_i - t), RnsTool.swift:168 tIncrement) followed by the forward NTT. */
int orc_plaintext_to_eval(const orc_context *ctx, const uint64_t *plain, int32_t l, uint64_t *out) {
    const i64 n = ctx->n;
    if (l < 1 || l > ctx->L) return -1;
    const u64 t = ctx->t, threshold = (t + 1) / 2;
    for (int r = 0; r < l; r++) {
        const u64 inc = ctx->q[r] - t;
        u64 *row = out + (i64)r * n;
        for (i64 c = 0; c < n; c++) row[c] = plain[c] < threshold ? plain[c] : plain[c] + inc;
        ntt_forward_row(ctx->qt[r], row);
    }
    return 0;
}

/* Bfv.innerProduct(ciphertexts:plaintexts:), Bfv.swift:476-505 with lazyMultiply :388-400, reduceInPlace :365-377 and
 * reduceToCiphertext :380-394, evaluated for `out_count` independent plaintext rows against the same `terms`
 * ciphertexts (the MulPir first-dimension scan, IndexPir/PirUtil.swift:437-442).
 *   cts: terms x npoly x l x n (Eval); pts: out_count x terms x l x n (Eval); present: out_count x terms flags
 *   (0 = nil plaintext, skipped, :493) or NULL; out: out_count x npoly x l x n (Eval). */
int orc_inner_product_plain(const orc_context *ctx, const uint64_t *cts, int32_t npoly, int32_t l, int64_t terms,
                            const uint64_t *pts, const uint8_t *present, uint64_t *out, int64_t out_count,
                            int32_t threads) {
    const i64 n = ctx->n;
    if (l < 1 || l > ctx->L || npoly < 1) return -1;
    if (threads <= 0) threads = orc_num_threads();
    /* maxLazyProductAccumulationCount, PolyContext.swift:246-253 */
    u64 qmax = 0;
    for (int r = 0; r < l; r++) qmax = ctx->q[r] > qmax ? ctx->q[r] : qmax;
    const u128 max_product = (u128)(qmax - 1) * (qmax - 1);
    const u128 max_count128 = ((~(u128)0) - qmax) / max_product;
    const i64 max_count = max_count128 > (u128)INT64_MAX ? INT64_MAX : (i64)max_count128;
#pragma omp parallel for num_threads(threads) schedule(dynamic, 1)
    for (i64 o = 0; o < out_count; o++) {
        u128 *acc = (u128 *)calloc((size_t)npoly * l * n, sizeof(u128));
        i64 reduce_count = 0;
        for (i64 k = 0; k < terms; k++) {
            if (present && !present[o * terms + k]) continue;
            const u64 *pt = pts + (o * terms + k) * l * n;
            for (int p = 0; p < npoly; p++) {
                const u64 *ct = cts + ((k * npoly + p) * l) * n;
                u128 *a = acc + (i64)p * l * n;
                for (i64 i = 0; i < (i64)l * n; i++) a[i] += (u128)ct[i] * pt[i]; /* addingLazyProduct, PolyRq.swift:210-225 */
            }
            if (++reduce_count >= max_count) {
                reduce_count = 0;
                for (int p = 0; p < npoly; p++)
                    for (int r = 0; r < l; r++) {
                        modulus_t m = modulus_make(ctx->q[r]);
                        u128 *a = acc + ((i64)p * l + r) * n;
                        for (i64 c = 0; c < n; c++) a[c] = reduce_double(&m, a[c]);
                    }
            }
        }
        for (int p = 0; p < npoly; p++)
            for (int r = 0; r < l; r++) {
                modulus_t m = modulus_make(ctx->q[r]);
                const u128 *a = acc + ((i64)p * l + r) * n;
                u64 *dst = out + ((o * npoly + p) * l + r) * n;
                for (i64 c = 0; c < n; c++) dst[c] = reduce_double(&m, a[c]);
            }
        free(acc);
    }
    return 0;
}

/* Bfv.innerProduct(_:_:) over ciphertext pairs, Bfv.swift:315-361: lazy 128-bit accumulation of the tensor products in
 * [Q, Bsk] (lazyMultiply :319-331), reduceToCiphertext (:380-394), one dropExtendedBase (Bfv+Multiply.swift:31-48).
 * lhs, rhs: groups x pairs x 2 x L x n (Coeff); out: groups x 3 x L x n (Coeff). */
int orc_bfv_inner_product(const orc_context *ctx, const uint64_t *lhs, const uint64_t *rhs, uint64_t *out, int64_t pairs,
                          int64_t groups, int32_t threads) {
    const i64 n = ctx->n;
    const int L = ctx->L, R = 2 * L + 1;
    const orc_rnstool *rt = ctx->tools[L];
    const i64 psz = (i64)R * n, ct_in = (i64)2 * L * n;
    if (threads <= 0) threads = orc_num_threads();
    /* maxProductCount = maxLazyProductAccumulationCount(qBsk) / 2, Bfv.swift:331 */
    u64 pmax = 0;
    for (int r = 0; r < R; r++) pmax = rt->qbsk[r] > pmax ? rt->qbsk[r] : pmax;
    const u128 max_lazy = ((~(u128)0) - pmax) / ((u128)(pmax - 1) * (pmax - 1));
    const i64 max_count = (i64)((max_lazy / 2) > (u128)INT64_MAX ? (u128)INT64_MAX : (max_lazy / 2));
#pragma omp parallel for num_threads(threads) schedule(dynamic, 1)
    for (i64 g = 0; g < groups; g++) {
        u128 *acc = (u128 *)calloc((size_t)3 * psz, sizeof(u128));
        u64 *buf = (u64 *)malloc(sizeof(u64) * psz * 4);
        u64 *l0 = buf, *l1 = buf + psz, *r0 = buf + 2 * psz, *r1 = buf + 3 * psz;
        i64 count = 0;
        for (i64 k = 0; k < pairs; k++) {
            const u64 *a = lhs + (g * pairs + k) * ct_in, *b = rhs + (g * pairs + k) * ct_in;
            behz_poly(ctx, a, l0);
            behz_poly(ctx, a + (i64)L * n, l1);
            behz_poly(ctx, b, r0);
            behz_poly(ctx, b + (i64)L * n, r1);
            for (i64 i = 0; i < psz; i++) {
                acc[i] += (u128)l0[i] * r0[i];
                acc[psz + i] += (u128)l0[i] * r1[i];
                acc[psz + i] += (u128)l1[i] * r0[i];
                acc[2 * psz + i] += (u128)l1[i] * r1[i];
            }
            if (++count >= (max_count > 0 ? max_count : 1)) {
                count = 0;
                for (int p = 0; p < 3; p++)
                    for (int r = 0; r < R; r++) {
                        modulus_t m = modulus_make(rt->qbsk[r]);
                        u128 *ar = acc + (i64)p * psz + (i64)r * n;
                        for (i64 c = 0; c < n; c++) ar[c] = reduce_double(&m, ar[c]);
                    }
            }
        }
        u64 *sum = (u64 *)malloc(sizeof(u64) * psz);
        u64 tvec[2 * ORC_MAX_MODULI];
        for (int r = 0; r < R; r++) tvec[r] = ctx->t;
        for (int p = 0; p < 3; p++) {
            for (int r = 0; r < R; r++) {
                modulus_t m = modulus_make(rt->qbsk[r]);
                for (i64 c = 0; c < n; c++) sum[(i64)r * n + c] = reduce_double(&m, acc[(i64)p * psz + (i64)r * n + c]);
            }
            poly_mul_scalar_rows(n, rt->qbsk, R, sum, tvec);
            for (int r = 0; r < R; r++) ntt_inverse_row(rt->qbsk_tables[r], sum + (i64)r * n);
            orc_rnstool_floor(rt, sum, out + (g * 3 + p) * (i64)L * n);
        }
        free(sum);
        free(buf);
        free(acc);
    }
    return 0;
}

/* PolyRq<Coeff>.multiplyPowerOfX, PolyRq.swift:398-422: multiplication by X^power in Z_q[X]/(X^N + 1) (power may be
 * negative): rotate the coefficient columns and negate the wrapped ones. */
void orc_multiply_power_of_x(int64_t n, const uint64_t *moduli, int32_t nmod, int64_t power, const uint64_t *in,
                             uint64_t *out) {
    const i64 two_n = 2 * n;
    i64 s = power % two_n;
    if (s < 0) s += two_n;
    for (int r = 0; r < nmod; r++)
        for (i64 i = 0; i < n; i++) {
            const i64 raw = (i + s) % two_n;
            const u64 v = in[(i64)r * n + i];
            out[(i64)r * n + (raw % n)] = raw >= n ? neg_mod(v, moduli[r]) : v;
        }
}
