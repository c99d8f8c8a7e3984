// drbg.cuh -- AES-128 (FIPS-197) block encryption and the CTR_DRBG update step as __host__ __device__ functions, shared by
// drbg.cu and its host-side emulation test (tests/emu/drbg_emulate.cu).
//
// State and round keys are big-endian 32-bit columns.  Rounds 1..9 use one 1 KB table Te0[x] = (2 S[x], S[x], S[x], 3 S[x])
// -- SubBytes, ShiftRows and MixColumns of FIPS-197 5.1.1-5.1.3 merged, the other three column tables being byte rotations of
// it -- and the last round the S-box alone; both tables are derived from the field arithmetic in make_tables().
#pragma once
#include "modarith.cuh"

namespace hecuda {
namespace drbg {

typedef unsigned int u32w;
constexpr int kSegmentBytes = 4096, kSegmentBlocks = kSegmentBytes / 16, kRoundKeyWords = 44;

inline void make_tables(unsigned char *sbox, u32w *te0) {  // S-box: inverse in GF(2^8), then the affine map (FIPS-197 5.1.1)
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
    for (int i = 0; i < 256; ++i) {
        const unsigned char s = sbox[i], s2 = (unsigned char)((s << 1) ^ ((s & 0x80) ? 0x1B : 0)), s3 = (unsigned char)(s2 ^ s);
        te0[i] = ((u32w)s2 << 24) | ((u32w)s << 16) | ((u32w)s << 8) | s3;
    }
}

HE_HD u32w ror32(u32w x, int r) { return (x >> r) | (x << (32 - r)); }
HE_HD u32w sub_word(u32w w, const unsigned char *sbox) {
    return ((u32w)sbox[w >> 24] << 24) | ((u32w)sbox[(w >> 16) & 255] << 16) | ((u32w)sbox[(w >> 8) & 255] << 8) | sbox[w & 255];
}

// KeyExpansion (FIPS-197 5.2), 128-bit key: 44 words
HE_HD void expand_key(const u32w key[4], u32w *rk, const unsigned char *sbox) {
    for (int i = 0; i < 4; ++i) rk[i] = key[i];
    u32w rcon = 0x01000000u;
    for (int i = 4; i < kRoundKeyWords; ++i) {
        u32w t = rk[i - 1];
        if ((i & 3) == 0) {
            t = sub_word((t << 8) | (t >> 24), sbox) ^ rcon;
            rcon = (rcon << 1) ^ ((rcon & 0x80000000u) ? 0x1B000000u : 0);
        }
        rk[i] = rk[i - 4] ^ t;
    }
}

// Cipher (FIPS-197 5.1) on the four big-endian columns s[0..3]
HE_HD void encrypt_block(u32w s[4], const u32w *rk, const u32w *te0, const unsigned char *sbox) {
    u32w a = s[0] ^ rk[0], b = s[1] ^ rk[1], c = s[2] ^ rk[2], d = s[3] ^ rk[3];
    for (int round = 1; round < 10; ++round) {
        const u32w *k = rk + 4 * round;
        const u32w na = te0[a >> 24] ^ ror32(te0[(b >> 16) & 255], 8) ^ ror32(te0[(c >> 8) & 255], 16) ^ ror32(te0[d & 255], 24) ^ k[0];
        const u32w nb = te0[b >> 24] ^ ror32(te0[(c >> 16) & 255], 8) ^ ror32(te0[(d >> 8) & 255], 16) ^ ror32(te0[a & 255], 24) ^ k[1];
        const u32w nc = te0[c >> 24] ^ ror32(te0[(d >> 16) & 255], 8) ^ ror32(te0[(a >> 8) & 255], 16) ^ ror32(te0[b & 255], 24) ^ k[2];
        const u32w nd = te0[d >> 24] ^ ror32(te0[(a >> 16) & 255], 8) ^ ror32(te0[(b >> 8) & 255], 16) ^ ror32(te0[c & 255], 24) ^ k[3];
        a = na, b = nb, c = nc, d = nd;
    }
    const u32w *k = rk + 40;
    s[0] = (((u32w)sbox[a >> 24] << 24) | ((u32w)sbox[(b >> 16) & 255] << 16) | ((u32w)sbox[(c >> 8) & 255] << 8) | sbox[d & 255]) ^ k[0];
    s[1] = (((u32w)sbox[b >> 24] << 24) | ((u32w)sbox[(c >> 16) & 255] << 16) | ((u32w)sbox[(d >> 8) & 255] << 8) | sbox[a & 255]) ^ k[1];
    s[2] = (((u32w)sbox[c >> 24] << 24) | ((u32w)sbox[(d >> 16) & 255] << 16) | ((u32w)sbox[(a >> 8) & 255] << 8) | sbox[b & 255]) ^ k[2];
    s[3] = (((u32w)sbox[d >> 24] << 24) | ((u32w)sbox[(a >> 16) & 255] << 16) | ((u32w)sbox[(b >> 8) & 255] << 8) | sbox[c & 255]) ^ k[3];
}

// the counter block V + add (128-bit big-endian increment, as AES._CTR with a 16-byte nonce)
HE_HD void counter_block(u64 hi, u64 lo, u64 add, u32w out[4]) {
    const u64 l = lo + add, h = hi + (l < lo ? 1 : 0);
    out[0] = (u32w)(h >> 32), out[1] = (u32w)h, out[2] = (u32w)(l >> 32), out[3] = (u32w)l;
}

// ctrDrbgUpdate (NistCtrDrbg.swift:62-69) given the two keystream blocks at V+1 and V+2: key <- block0 ^ provided[0..15],
// V <- block1 ^ provided[16..31] (provided as 8 big-endian words, or null for the all-zero additional input)
HE_HD void drbg_absorb(u32w key[4], u64 &hi, u64 &lo, const u32w b0[4], const u32w b1[4], const u32w *provided) {
    u32w v[4];
    for (int i = 0; i < 4; ++i) {
        key[i] = b0[i] ^ (provided ? provided[i] : 0);
        v[i] = b1[i] ^ (provided ? provided[4 + i] : 0);
    }
    hi = ((u64)v[0] << 32) | v[1];
    lo = ((u64)v[2] << 32) | v[3];
}

}  // namespace drbg
}  // namespace hecuda
