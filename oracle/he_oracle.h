/*
 * he_oracle.h -- CPU ORACLE (TEST INFRASTRUCTURE ONLY; never linked into the product).
 *
 * A plain-C restatement of the RNS-BFV hot path of apple/swift-homomorphic-encryption
 * (reference @ 68675885a9b1).  Every function cites the reference file:line it follows.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
 * load this library.  Parity status: pinned against the reference's own known-answer tests
 * (tests/test_oracle_kats.py) -- NTT vectors, minimal roots, generatePrimes, divideAndRoundQLast,
 * smallMontgomeryReduce KATs, lift/floor/Shenoy-Kumaresan big-integer properties,
 * decrypt-correctness of multiply+relinearize.  The Swift reference itself cannot be run in this
 * environment (no Swift toolchain), so full-pipeline golden ciphertexts come from this oracle.
 */
#ifndef HE_ORACLE_H
#define HE_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_context orc_context;

/* ---- scalar helpers (Sources/HomomorphicEncryption/Scalar.swift) ---- */
uint64_t orc_pow_mod(uint64_t base, uint64_t exp, uint64_t p);
uint64_t orc_inverse_mod(uint64_t a, uint64_t p); /* 0 if not invertible */
int orc_is_prime(uint64_t p);
/* returns number of primes found (== n on success) */
int orc_generate_primes(const int32_t *bit_counts, int32_t n, int32_t prefer_small, int64_t ntt_degree,
                        uint64_t *out);
uint64_t orc_min_primitive_root(int64_t degree, uint64_t p); /* 0 if none */
uint32_t orc_reverse_bits(uint32_t x, int32_t bit_count);

/* ---- modular arithmetic KAT hooks (Sources/ModularArithmetic/Modulus.swift) ---- */
uint64_t orc_barrett_reduce_single(uint64_t x, uint64_t p);
uint64_t orc_barrett_reduce_double(uint64_t hi, uint64_t lo, uint64_t p);
uint64_t orc_barrett_reduce_product(uint64_t x, uint64_t y, uint64_t p);
uint64_t orc_shoup_mul(uint64_t x, uint64_t multiplicand, uint64_t p);
uint64_t orc_shoup_mul_lazy(uint64_t x, uint64_t multiplicand, uint64_t p);

/* ---- standalone ring ops on rows (PolyRq+Ntt.swift, PolyRq.swift) ---- */
/* In-place forward / inverse negacyclic NTT of `rows` rows, row i under modulus moduli[i % nmod]. */
int orc_ntt_forward(int64_t n, const uint64_t *moduli, int32_t nmod, uint64_t *data, int64_t rows);
int orc_ntt_inverse(int64_t n, const uint64_t *moduli, int32_t nmod, uint64_t *data, int64_t rows);
int orc_ntt_forward_threads(int64_t n, const uint64_t *moduli, int32_t nmod, uint64_t *data, int64_t rows, int32_t threads);
/* Root tables as the reference builds them (for parity of the product's setup). */
int orc_ntt_tables(int64_t n, uint64_t p, uint64_t *roots, uint64_t *inv_roots_reordered, uint64_t *inv_degree,
                   uint64_t *inv_degree_root);
/* divideAndRoundQLast: data = nmod x n in, (nmod-1) x n out (in place, first rows). */
int orc_divide_round_qlast(int64_t n, const uint64_t *moduli, int32_t nmod, uint64_t *data);
/* coefficient-wise ops on a poly with `nmod` rows */
void orc_poly_add(int64_t n, const uint64_t *moduli, int32_t nmod, uint64_t *lhs, const uint64_t *rhs);
void orc_poly_sub(int64_t n, const uint64_t *moduli, int32_t nmod, uint64_t *lhs, const uint64_t *rhs);
void orc_poly_mul(int64_t n, const uint64_t *moduli, int32_t nmod, uint64_t *lhs, const uint64_t *rhs);

/* ---- RNS tool built from arbitrary Q (for the reference's RnsToolTests KATs) ---- */
typedef struct orc_rnstool orc_rnstool;
orc_rnstool *orc_rnstool_create(int64_t n, const uint64_t *q, int32_t nq, uint64_t t);
orc_rnstool *orc_rnstool_create_w(int64_t n, const uint64_t *q, int32_t nq, uint64_t t, int32_t word_bits);
void orc_rnstool_destroy(orc_rnstool *);
int32_t orc_rnstool_bsk(const orc_rnstool *, uint64_t *out); /* writes nq+1 Bsk primes, returns count */
/* poly in base [Bsk, m~] ((nq+2) x n) -> base Bsk ((nq+1) x n), in place */
void orc_rnstool_small_montgomery_reduce(const orc_rnstool *, uint64_t *data);
/* in: nq x n (base Q)  -> out: (2nq+1) x n (base [Q,Bsk]) */
void orc_rnstool_lift(const orc_rnstool *, const uint64_t *in, uint64_t *out);
/* in: (2nq+1) x n -> out: (nq+1) x n (base Bsk) */
void orc_rnstool_approximate_floor(const orc_rnstool *, const uint64_t *in, uint64_t *out);
/* in: (nq+1) x n (base Bsk) -> out: nq x n (base Q) */
void orc_rnstool_bsk_to_q(const orc_rnstool *, const uint64_t *in, uint64_t *out);
/* in: (2nq+1) x n -> out nq x n */
void orc_rnstool_floor(const orc_rnstool *, const uint64_t *in, uint64_t *out);
/* in: nq x n -> out: 1 x n mod t (decrypt scaling), scalingFactor as in the reference */
void orc_rnstool_scale_and_round(const orc_rnstool *, const uint64_t *in, uint64_t scaling_factor, uint64_t *out);
/* in: nq x n -> out (nq+2) x n in base [Bsk, m~] */
void orc_rnstool_convert_bsk_mtilde(const orc_rnstool *, const uint64_t *in, uint64_t *out);
/* generic fast base conversion q[] -> tmod[]: in nq x n, out nt x n (RnsBaseConverter.swift:68-73) */
void orc_convert_approximate(int64_t n, const uint64_t *q, int32_t nq, const uint64_t *tmod, int32_t nt,
                             const uint64_t *in, uint64_t *out);

/* ---- BFV context (Context.swift:94-143) ---- */
/* coeff_moduli = q_0..q_{L-1}, q_ks (nmod = L+1 >= 2); t = plaintext modulus */
orc_context *orc_context_create(int64_t n, const uint64_t *coeff_moduli, int32_t nmod, uint64_t t);
orc_context *orc_context_create_w(int64_t n, const uint64_t *coeff_moduli, int32_t nmod, uint64_t t, int32_t word_bits);
void orc_context_destroy(orc_context *);
int32_t orc_context_L(const orc_context *);
int32_t orc_context_bsk(const orc_context *, uint64_t *out);

/* ct x ct multiply at the top level (Bfv+Multiply.swift:18-85).
 * a, b: batch x 2 x L x n (Coeff) ; out: batch x 3 x L x n (Coeff).  threads<=0 -> all cores. */
int orc_bfv_mul(const orc_context *, const uint64_t *a, const uint64_t *b, uint64_t *out, int64_t batch,
                int32_t threads);
/* stages of the multiply, for stage-level parity */
int orc_bfv_lift_ntt(const orc_context *, const uint64_t *polys, uint64_t *out, int64_t npolys); /* L x n -> R x n, Eval */
/* relinearization key layout: L x 2 x K x n (Eval), K = L+1 (Keys.swift:66-99) */
int orc_keygen(const orc_context *, uint64_t seed, uint64_t *secret_key_eval /* (L+1) x n */,
               uint64_t *relin_key /* L x 2 x K x n */);
/* generic key-switch key for current key `cur` (Eval, (L+1) x n rows; only first L used) -> target sk */
int orc_gen_keyswitch_key(const orc_context *, uint64_t seed, const uint64_t *secret_key_eval,
                          const uint64_t *current_key_eval, uint64_t *ksk);
/* key switching update of one target poly with l rows (Coeff): out 2 x l x n (Bfv+Keys.swift:123-208) */
int orc_keyswitch_update(const orc_context *, const uint64_t *target, int32_t l, const uint64_t *ksk,
                         uint64_t *out);
/* relinearize: in batch x 3 x l x n -> out batch x 2 x l x n (Bfv.swift:201-219) */
int orc_bfv_relinearize(const orc_context *, const uint64_t *ct3, int32_t l, const uint64_t *relin_key,
                        uint64_t *out, int64_t batch, int32_t threads);
/* modSwitchDown: in batch x npoly x l x n -> out batch x npoly x (l-1) x n (Bfv.swift:163-171) */
int orc_bfv_mod_switch_down(const orc_context *, const uint64_t *ct, int32_t npoly, int32_t l, uint64_t *out,
                            int64_t batch, int32_t threads);
/* encrypt a coefficient-encoded plaintext (n values < t) under sk: out 2 x L x n (Bfv+Encrypt.swift:64-181) */
int orc_encrypt(const orc_context *, uint64_t seed, const uint64_t *secret_key_eval, const uint64_t *plain,
                uint64_t *ct);
/* decrypt npoly-poly ciphertext with l rows: out n values < t (Bfv+Decrypt.swift:21-41,188-204) */
int orc_decrypt(const orc_context *, const uint64_t *secret_key_eval, const uint64_t *ct, int32_t npoly,
                int32_t l, uint64_t *plain);

/* ---- Galois automorphisms: PolyRq/Galois.swift:115-166, Bfv.swift:174-198 ---- */
void orc_galois_coeff(int64_t n, const uint64_t *moduli, int32_t nmod, int64_t element, const uint64_t *in, uint64_t *out);
void orc_galois_eval(int64_t n, int32_t nmod, int64_t element, const uint64_t *in, uint64_t *out);
int64_t orc_galois_element_rotating_columns(int64_t step, int64_t degree); /* 0 on invalid step */
int64_t orc_galois_element_swapping_rows(int64_t degree);
/* Galois key for `element`: key-switch key from s(x^element) to s (Bfv+Keys.swift:42-49); layout like the relin key */
int orc_gen_galois_key(const orc_context *, uint64_t seed, const uint64_t *secret_key_eval, int64_t element, uint64_t *ksk);
/* applyGalois: ct batch x 2 x l x n (Coeff) -> out batch x 2 x l x n */
int orc_bfv_apply_galois(const orc_context *, const uint64_t *ct, int32_t l, int64_t element, const uint64_t *galois_key,
                         uint64_t *out, int64_t batch, int32_t threads);

/* ---- lazy ciphertext x plaintext inner product: Bfv.swift:402-505, Plaintext.swift:149-171 ---- */
int orc_plaintext_to_eval(const orc_context *, const uint64_t *plain /* n values < t */, int32_t l, uint64_t *out /* l x n */);
int orc_inner_product_plain(const orc_context *, const uint64_t *cts, int32_t npoly, int32_t l, int64_t terms,
                            const uint64_t *pts, const uint8_t *present, uint64_t *out, int64_t out_count, int32_t threads);

int orc_bfv_inner_product(const orc_context *, const uint64_t *lhs, const uint64_t *rhs, uint64_t *out, int64_t pairs,
                          int64_t groups, int32_t threads);

void orc_multiply_power_of_x(int64_t n, const uint64_t *moduli, int32_t nmod, int64_t power, const uint64_t *in,
                             uint64_t *out); /* PolyRq.swift:398-422 */

/* deterministic test inputs: uniform residues row r < moduli[r % nmod] (splitmix64, rejection-free mod) */
void orc_fill_uniform(uint64_t seed, const uint64_t *moduli, int32_t nmod, int64_t n, uint64_t *data, int64_t rows);
int orc_num_threads(void);

#ifdef __cplusplus
}
#endif
#endif
