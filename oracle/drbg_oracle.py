"""TEST INFRASTRUCTURE ONLY -- restatement of the reference's seeded-ciphertext expansion.

  * NistCtrDrbg          HomomorphicEncryption/Random/NistCtrDrbg.swift:25-84 (NIST SP 800-90A CTR_DRBG, AES-128,
                         no derivation function).  The block cipher is a third-party dependency of the reference
                         (swift-crypto 3.15.1, `AES._CTR`, un-vendored); here it is the `cryptography` package's AES --
                         the published algorithm (FIPS 197), pinned by the reference's own NIST vectors.
  * NistAes128Ctr        Random/NistAes128Ctr.swift:17-40 = BufferedRng<NistCtrDrbg> with a 4096-byte buffer
                         (Random/BufferedRng.swift:17-67): the byte stream is the concatenation of 4096-byte generates.
  * PolyRq.randomizeUniform  PolyRq/PolyRq+Randomize.swift:49-81: coefficient k of row r = the k-th little-endian
                         128-bit word of the stream (rows consecutive) reduced modulo q_r.
  * Ciphertext(deserialize: .seeded)  SerializedCiphertext.swift:41-60: poly0 deserialized, poly1 = the random Eval
                         polynomial converted to the canonical (Coeff) format.

Only tests/, smoke() and bench.py's baseline legs may import this module.
"""
from __future__ import annotations

import numpy as np
from cryptography.hazmat.primitives.ciphers import Cipher, algorithms, modes

from . import oracle as O
from .pir_oracle import load_poly

BUFFER_COUNT = 4096  # NistAes128Ctr bufferCount
MASK128 = (1 << 128) - 1


class NistCtrDrbg:
    def __init__(self, entropy: bytes):
        assert len(entropy) == 32
        self.key, self.nonce, self.reseed_counter = bytes(16), 0, 1
        self._update(entropy)

    def _keystream(self, count: int) -> bytes:
        # AES._CTR.encrypt(zeros, nonce: (V + 1).bigEndianBytes): counter blocks V+1, V+2, ... (128-bit big-endian increment)
        enc = Cipher(algorithms.AES(self.key), modes.ECB()).encryptor()
        blocks = -(-count // 16)
        data = b"".join(((self.nonce + 1 + i) & MASK128).to_bytes(16, "big") for i in range(blocks))
        return enc.update(data)[:count]

    def _update(self, provided: bytes):
        xor = bytes(a ^ b for a, b in zip(self._keystream(32), provided))
        self.key, self.nonce = xor[:16], int.from_bytes(xor[16:], "big")

    def generate(self, count: int) -> bytes:
        assert count <= 1 << 16
        out = self._keystream(count)
        self.nonce = (self.nonce + -(-count // 16)) & MASK128
        self._update(bytes(32))
        self.reseed_counter += 1
        return out


class NistAes128Ctr:
    """BufferedRng<NistCtrDrbg>(bufferCount: 4096)."""

    def __init__(self, seed: bytes):
        self.rng, self.buffer, self.offset = NistCtrDrbg(seed), b"", 0

    def fill(self, count: int) -> bytes:
        out = bytearray()
        while len(out) < count:
            if self.offset == len(self.buffer):
                self.buffer, self.offset = self.rng.generate(BUFFER_COUNT), 0
            take = min(count - len(out), len(self.buffer) - self.offset)
            out += self.buffer[self.offset:self.offset + take]
            self.offset += take
        return bytes(out)


def random_poly(n: int, moduli, seed: bytes) -> np.ndarray:
    """PolyRq.random(context:using:) with NistAes128Ctr(seed:) -> rows x N residues."""
    rng = NistAes128Ctr(seed)
    chunk = min(n, 1024)
    out = np.zeros((len(moduli), n), dtype=np.uint64)
    for r, q in enumerate(moduli):
        for start in range(0, n, chunk):
            data = rng.fill(chunk * 16)
            for i in range(chunk):
                out[r, start + i] = int.from_bytes(data[16 * i:16 * i + 16], "little") % int(q)
    return out


def expand_seeded_ciphertext(ctx: O.Context, poly0_serialized: bytes, seed: bytes, moduli_count: int = 0) -> np.ndarray:
    """Ciphertext(deserialize: .seeded(poly0:seed:)) for BFV (canonical format Coeff) -> (2, l, N)."""
    l = moduli_count or ctx.L
    q = ctx.q[:l]
    poly0 = load_poly(ctx.n, q, poly0_serialized)
    a = random_poly(ctx.n, q, seed)
    return np.stack([poly0, O.ntt_inverse(ctx.n, q, a)])
