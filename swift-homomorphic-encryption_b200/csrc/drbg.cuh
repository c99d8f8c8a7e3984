// drbg.cuh -- AES-128 (FIPS-197) block encryption and the CTR_DRBG update step as __host__ __device__ functions, shared by
// drbg.cu and its host-side emulation test (tests/emu/drbg_emulate.cu).
#pragma once
#include "modarith.cuh"

namespace hecuda {
namespace drbg {

constexpr int kSegmentBytes = 4096, kSegmentBlocks = kSegmentBytes / 16, kRoundKeyBytes = 176;

inline void make_sbox(unsigned char *sbox) {  // FIPS-197 5.1.1: multiplicative inverse in GF(2^8) followed by the affine map
    unsigned char p = 1, q = 1;
    do {
        p = (unsigned char)(p ^ (p << 1) ^ ((p & 0x80) ? 0x1B : 0));  // p *= 3
        q ^= (unsigned char)(q << 1);                                  // q /= 3
        q ^= (unsigned char)(q << 2);
        q ^= (unsigned char)(q << 4);
        if (q & 0x80) q ^= 0x09;
        const unsigned char x = (unsigned char)(q ^ (q << 1 | q >> 7) ^ (q << 2 | q >> 6) ^ (q << 3 | q >> 5) ^ (q << 4 | q >> 4));
        sbox[p] = (unsigned char)(x ^ 0x63);
    } while (p != 1);
    sbox[0] = 0x63;
}

HE_HD unsigned char xtime(unsigned char x) { return (unsigned char)((x << 1) ^ ((x & 0x80) ? 0x1B : 0)); }

// KeyExpansion (FIPS-197 5.2) for a 128-bit key: 11 round keys of 16 bytes
HE_HD void expand_key(const unsigned char *key, unsigned char *rk, const unsigned char *sbox) {
    for (int i = 0; i < 16; ++i) rk[i] = key[i];
    unsigned char rcon = 1;
    for (int i = 16; i < kRoundKeyBytes; i += 4) {
        unsigned char t0 = rk[i - 4], t1 = rk[i - 3], t2 = rk[i - 2], t3 = rk[i - 1];
        if ((i & 15) == 0) {
            const unsigned char r0 = sbox[t1] ^ rcon, r1 = sbox[t2], r2 = sbox[t3], r3 = sbox[t0];
            t0 = r0, t1 = r1, t2 = r2, t3 = r3;
            rcon = xtime(rcon);
        }
        rk[i] = rk[i - 16] ^ t0;
        rk[i + 1] = rk[i - 15] ^ t1;
        rk[i + 2] = rk[i - 14] ^ t2;
        rk[i + 3] = rk[i - 13] ^ t3;
    }
}

// Cipher (FIPS-197 5.1); state is column-major: byte r of column c at s[4c + r]
HE_HD void encrypt_block(unsigned char *s, const unsigned char *rk, const unsigned char *sbox) {
#pragma unroll
    for (int i = 0; i < 16; ++i) s[i] ^= rk[i];
    for (int round = 1; round <= 10; ++round) {
        unsigned char t[16];
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int r = 0; r < 4; ++r) t[4 * c + r] = sbox[s[4 * ((c + r) & 3) + r]];  // SubBytes + ShiftRows
        if (round < 10) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {  // MixColumns
                const unsigned char a0 = t[4 * c], a1 = t[4 * c + 1], a2 = t[4 * c + 2], a3 = t[4 * c + 3];
                const unsigned char all = a0 ^ a1 ^ a2 ^ a3;
                s[4 * c] = a0 ^ all ^ xtime(a0 ^ a1);
                s[4 * c + 1] = a1 ^ all ^ xtime(a1 ^ a2);
                s[4 * c + 2] = a2 ^ all ^ xtime(a2 ^ a3);
                s[4 * c + 3] = a3 ^ all ^ xtime(a3 ^ a0);
            }
        } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) s[i] = t[i];
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) s[i] ^= rk[16 * round + i];
    }
}

HE_HD void counter_block(u64 hi, u64 lo, u64 add, unsigned char *out) {  // (V + add) big-endian
    const u64 l = lo + add, h = hi + (l < lo ? 1 : 0);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        out[i] = (unsigned char)(h >> (56 - 8 * i));
        out[8 + i] = (unsigned char)(l >> (56 - 8 * i));
    }
}

// ctrDrbgUpdate (NistCtrDrbg.swift:62-69): (key, V) <- first 32 keystream bytes at V+1, V+2, xored with `provided`
HE_HD void drbg_update(unsigned char *key, u64 &hi, u64 &lo, const unsigned char *rk, const unsigned char *provided,
                            const unsigned char *sbox) {
    unsigned char b0[16], b1[16];
    counter_block(hi, lo, 1, b0);
    counter_block(hi, lo, 2, b1);
    encrypt_block(b0, rk, sbox);
    encrypt_block(b1, rk, sbox);
    hi = lo = 0;
    for (int i = 0; i < 16; ++i) {
        key[i] = b0[i] ^ (provided ? provided[i] : 0);
        const unsigned char v = b1[i] ^ (provided ? provided[16 + i] : 0);
        if (i < 8) hi = (hi << 8) | v; else lo = (lo << 8) | v;
    }
}

}  // namespace drbg
}  // namespace hecuda
