// drbg_emulate.cu -- replays on the host the __host__ __device__ AES-128 / CTR_DRBG functions the device kernels call
// (csrc/drbg.cuh): prints the first `count` stream bytes of NistAes128Ctr(seed) as hex.
//   usage: drbg_emulate <seed hex (64 chars)> <count>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../swift-homomorphic-encryption_b200/csrc/drbg.cuh"

using namespace hecuda;
using namespace hecuda::drbg;

int main(int argc, char **argv) {
    if (argc < 3 || std::strlen(argv[1]) != 64) return 2;
    unsigned char seed[32], sbox[256];
    for (int i = 0; i < 32; ++i) {
        unsigned v = 0;
        std::sscanf(argv[1] + 2 * i, "%2x", &v);
        seed[i] = (unsigned char)v;
    }
    const long count = std::atol(argv[2]);
    make_sbox(sbox);
    unsigned char key[16] = {0}, rk[kRoundKeyBytes];
    u64 hi = 0, lo = 0;
    expand_key(key, rk, sbox);
    drbg_update(key, hi, lo, rk, seed, sbox);
    std::vector<unsigned char> out;
    while ((long)out.size() < count) {  // one 4096-byte segment per iteration, exactly like the two kernels
        expand_key(key, rk, sbox);
        for (int i = 0; i < kSegmentBlocks; ++i) {
            unsigned char block[16];
            counter_block(hi, lo, 1 + (u64)i, block);
            encrypt_block(block, rk, sbox);
            out.insert(out.end(), block, block + 16);
        }
        const u64 l = lo + kSegmentBlocks;
        hi += l < lo ? 1 : 0;
        lo = l;
        drbg_update(key, hi, lo, rk, nullptr, sbox);
    }
    for (long i = 0; i < count; ++i) std::printf("%02x", out[i]);
    std::printf("\n");
    return 0;
}
