/* Multi-GPU setup through the C ABI alone (no torch, no MPI): one process per GPU, the evaluation key created on rank 0
 * and broadcast over NCCL, then every rank relinearizes its own slice of a batch (no data-path collective).
 *
 *   evk_broadcast <rank> <world_size> <id-file>
 *
 * Rank 0 writes the 128-byte NCCL unique id to <id-file>; the other ranks wait for it (the host's own channel -- a
 * Swift server would hand it over its RPC layer).  Each rank uses GPU <rank> and prints a checksum of its relinearized
 * slice; the sum over ranks equals the checksum a single process prints for the whole batch (world_size 1).
 *
 *   gcc -std=c99 -O2 -Iinclude examples/evk_broadcast.c -Lswift-homomorphic-encryption_b200 -lhecuda \
 *       -Wl,-rpath,$PWD/swift-homomorphic-encryption_b200 -o evk_broadcast
 */
#define _POSIX_C_SOURCE 200809L
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

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

int main(int argc, char **argv) {
    if (argc != 4) {
        fprintf(stderr, "usage: %s <rank> <world_size> <id-file>\n", argv[0]);
        return 2;
    }
    const int rank = atoi(argv[1]), world = atoi(argv[2]);
    const char *id_path = argv[3];
    const uint64_t moduli[4] = {36028797018652673ull, 36028797017571329ull, 36028797017456641ull, 36028797017276417ull};
    const int64_t n = 8192, batch = 8;
    const int32_t L = 3, K = 4;
    const uint32_t galois_element = 3;

    CHECK(hecuda_set_device(rank));
    int32_t node = -1, cpus = 0;
    CHECK(hecuda_bind_host_to_device(rank, &node, &cpus));

    hecuda_comm *comm = NULL;
    if (world > 1) {
        uint8_t id[HECUDA_COMM_UNIQUE_ID_BYTES];
        if (rank == 0) {
            CHECK(hecuda_comm_unique_id(id));
            char tmp[4096];
            snprintf(tmp, sizeof(tmp), "%s.tmp", id_path);
            FILE *f = fopen(tmp, "wb");
            if (!f || fwrite(id, 1, sizeof(id), f) != sizeof(id) || fclose(f) != 0 || rename(tmp, id_path) != 0) return 3;
        } else {
            FILE *f = NULL;
            for (int tries = 0; tries < 3000 && !(f = fopen(id_path, "rb")); ++tries) {
                struct timespec ts = {0, 10 * 1000 * 1000};
                nanosleep(&ts, NULL);
            }
            if (!f || fread(id, 1, sizeof(id), f) != sizeof(id)) return 3;
            fclose(f);
        }
        CHECK(hecuda_comm_create(id, rank, world, &comm));
    }

    hecuda_context *ctx = NULL;
    hecuda_evk *evk = NULL;
    CHECK(hecuda_context_create(n, moduli, 4, 557057, &ctx));
    const size_t key_words = (size_t)L * 2 * K * n;
    if (rank == 0) { /* the rank that received the key from the client */
        uint64_t *key = (uint64_t *)malloc(2 * key_words * sizeof(uint64_t));
        if (!key) return 1;
        uint64_t seed = 99;
        for (int which = 0; which < 2; ++which)
            for (int i = 0; i < L * 2; ++i)
                for (int r = 0; r < K; ++r)
                    for (int64_t c = 0; c < n; ++c)
                        key[which * key_words + ((size_t)i * K + r) * n + c] = next_random(&seed) % moduli[r];
        CHECK(hecuda_evk_create(ctx, key, &evk));
        CHECK(hecuda_evk_set_galois_key(evk, galois_element, key + key_words));
        free(key);
    } else {
        CHECK(hecuda_evk_create_empty(ctx, &evk));
    }
    if (world > 1) CHECK(hecuda_evk_broadcast(evk, comm, 0, 1, &galois_element, 1));

    /* every rank generates the same synthetic batch and works on its contiguous slice */
    const size_t poly = (size_t)L * n;
    uint64_t *ct3 = (uint64_t *)malloc(batch * 3 * poly * sizeof(uint64_t));
    uint64_t *out = (uint64_t *)malloc(batch * 2 * poly * sizeof(uint64_t));
    uint64_t *rot = (uint64_t *)malloc(batch * 2 * poly * sizeof(uint64_t));
    if (!ct3 || !out || !rot) return 1;
    uint64_t seed = 7;
    for (int64_t b = 0; b < batch * 3; ++b)
        for (int r = 0; r < L; ++r)
            for (int64_t c = 0; c < n; ++c) ct3[(b * L + r) * n + c] = next_random(&seed) % moduli[r];
    const int64_t lo = batch * rank / world, hi = batch * (rank + 1) / world;
    CHECK(hecuda_bfv_relinearize(ctx, evk, ct3 + lo * 3 * poly, L, out + lo * 2 * poly, hi - lo));
    CHECK(hecuda_bfv_apply_galois(ctx, evk, out + lo * 2 * poly, L, galois_element, rot + lo * 2 * poly, hi - lo));
    uint64_t checksum = 0;
    for (size_t i = (size_t)lo * 2 * poly; i < (size_t)hi * 2 * poly; ++i) checksum += out[i] * 31 + rot[i];
    printf("rank %d of %d (numa node %d): ciphertexts [%lld, %lld) checksum %llu\n", rank, world, (int)node, (long long)lo,
           (long long)hi, (unsigned long long)checksum);

    free(ct3), free(out), free(rot);
    CHECK(hecuda_evk_destroy(evk));
    CHECK(hecuda_context_destroy(ctx));
    CHECK(hecuda_comm_destroy(comm));
    return 0;
}
