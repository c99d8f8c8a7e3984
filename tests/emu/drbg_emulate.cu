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
    u32w te0[256];
    for (int i = 0; i < 32; ++i) {
        unsigned v = 0;
        std::sscanf(argv[1] + 2 * i, "%2x", &v);
        seed[i] = (unsigned char)v;
    }
    const long count = std::atol(argv[2]);
    make_tables(sbox, te0);
    u32w key[4] = {0, 0, 0, 0}, rk[kRoundKeyWords], provided[8], b0[4], b1[4];
    for (int i = 0; i < 8; ++i)
        provided[i] = ((u32w)seed[4 * i] << 24) | ((u32w)seed[4 * i + 1] << 16) | ((u32w)seed[4 * i + 2] << 8) | seed[4 * i + 3];
    u64 hi = 0, lo = 0;
    std::vector<unsigned char> out;
    for (int s = -1; (long)out.size() < count; ++s) {  // the same loop as drbg_chain_kernel, both lanes in turn
        expand_key(key, rk, sbox);
        if (s >= 0) {
            for (int i = 0; i < kSegmentBlocks; ++i) {  // drbg_fill_kernel
                u32w blk[4];
                counter_block(hi, lo, 1 + (u64)i, blk);
                encrypt_block(blk, rk, te0, sbox);
                for (int w = 0; w < 4; ++w)
                    for (int k = 3; k >= 0; --k) out.push_back((unsigned char)(blk[w] >> (8 * k)));
            }
            const u64 l = lo + kSegmentBlocks;
            hi += l < lo ? 1 : 0;
            lo = l;
        }
        counter_block(hi, lo, 1, b0);
        counter_block(hi, lo, 2, b1);
        encrypt_block(b0, rk, te0, sbox);
        encrypt_block(b1, rk, te0, sbox);
        drbg_absorb(key, hi, lo, b0, b1, s < 0 ? provided : nullptr);
    }
    for (long i = 0; i < count; ++i) std::printf("%02x", out[i]);
    std::printf("\n");
    return 0;
}
