/* Minimal C client of the hecuda C ABI (include/hecuda.h): one batch of BFV ct x ct multiplications followed by
 * relinearization and a modulus switch, exactly the three HeScheme calls the reference's RlweBenchmark times
 * (Benchmarks/RlweBenchmark/RlweBenchmark.swift:387-493).  The inputs here are synthetic uniform residues.
 *
 *   gcc -std=c99 -O2 -Iinclude examples/multiply_relinearize.c -Lswift-homomorphic-encryption_b200 -lhecuda \
 *       -Wl,-rpath,$PWD/swift-homomorphic-encryption_b200 -o multiply_relinearize && ./multiply_relinearize
 */
#include <stdio.h>
#include <stdlib.h>

#include "hecuda.h"

#define CHECK(call)                                                                  \
    do {                                                                             \
        int32_t rc_ = (call);                                                        \
        if (rc_ != HECUDA_OK) {                                                      \
            fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, hecuda_last_error()); \
            return 1;                                                                \
        }                                                                            \
    } while (0)

static uint64_t next_random(uint64_t *state) { /* splitmix64 */
    uint64_t z = (*state += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

int main(void) {
    /* n_8192_logq_3x55 plus one key-switch modulus: BASELINE config 2 */
    const uint64_t moduli[4] = {36028797018652673ull, 36028797017571329ull, 36028797017456641ull, 36028797017276417ull};
    const int64_t n = 8192, batch = 16;
    const int32_t L = 3, K = 4;
    hecuda_context *ctx = NULL;
    hecuda_evk *evk = NULL;
    CHECK(hecuda_context_create(n, moduli, 4, 557057, &ctx));

    const size_t poly = (size_t)L * n;
    uint64_t *lhs, *rhs, *product, *relinearized, *switched, *key;
    /* pinned host memory lets the library overlap its copies with the kernels */
    CHECK(hecuda_host_alloc((void **)&lhs, batch * 2 * poly * sizeof(uint64_t)));
    CHECK(hecuda_host_alloc((void **)&rhs, batch * 2 * poly * sizeof(uint64_t)));
    CHECK(hecuda_host_alloc((void **)&product, batch * 3 * poly * sizeof(uint64_t)));
    CHECK(hecuda_host_alloc((void **)&relinearized, batch * 2 * poly * sizeof(uint64_t)));
    CHECK(hecuda_host_alloc((void **)&switched, batch * 2 * (poly - n) * sizeof(uint64_t)));
    key = (uint64_t *)malloc((size_t)L * 2 * K * n * sizeof(uint64_t));
    if (!key) return 1;

    uint64_t seed = 1;
    for (int64_t b = 0; b < batch * 2; ++b)
        for (int r = 0; r < L; ++r)
            for (int64_t c = 0; c < n; ++c) {
                lhs[(b * L + r) * n + c] = next_random(&seed) % moduli[r];
                rhs[(b * L + r) * n + c] = next_random(&seed) % moduli[r];
            }
    for (int i = 0; i < L * 2; ++i)
        for (int r = 0; r < K; ++r)
            for (int64_t c = 0; c < n; ++c) key[((size_t)i * K + r) * n + c] = next_random(&seed) % moduli[r];
    CHECK(hecuda_evk_create(ctx, key, &evk));

    CHECK(hecuda_bfv_multiply(ctx, lhs, rhs, product, batch));                     /* Bfv.mulAssign      */
    CHECK(hecuda_bfv_relinearize(ctx, evk, product, L, relinearized, batch));      /* Bfv.relinearize    */
    CHECK(hecuda_bfv_mod_switch_down(ctx, relinearized, 2, L, switched, batch));   /* Bfv.modSwitchDown  */

    printf("%lld products; first residue of the result: %llu; %llu kernel launches\n", (long long)batch,
           (unsigned long long)switched[0], (unsigned long long)hecuda_kernel_launch_count());
    hecuda_evk_destroy(evk);
    hecuda_context_destroy(ctx);
    hecuda_host_free(lhs), hecuda_host_free(rhs), hecuda_host_free(product), hecuda_host_free(relinearized), hecuda_host_free(switched);
    free(key);
    return 0;
}
