"""GPU parity for seeded-ciphertext expansion: AES-128 CTR_DRBG stream, uniform sampling and
Ciphertext(deserialize: .seeded) against oracle/drbg_oracle.py (pinned on the reference's NIST vectors)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import hecuda
from oracle import drbg_oracle as drbg
from oracle import oracle as orc
from oracle import pir_oracle as opir


@pytest.mark.parametrize("n,bits", [(16, [55, 52, 62, 58]), (64, [30, 61, 40]), (1024, [55, 55, 55]), (4096, [27, 28, 28]),
                                    (8192, [55, 55, 55, 55])])
def test_random_polys_and_seeded_expansion_match_oracle(n, bits):
    moduli = orc.generate_primes(bits, False, n)
    g, o = hecuda.Context(n, moduli, 17), orc.Context(n, moduli, 17)
    L = o.L
    rng = np.random.default_rng(n)
    seeds = rng.integers(0, 256, size=(3, 32), dtype=np.uint8)
    seeds[0] = np.frombuffer(bytes.fromhex("69a09f6bf5dda15cd4af29e14cf5e0cddd7d07ac39bba587f8bc331104f9c448"), dtype=np.uint8)
    for l in (L, 1):
        got = hecuda.Bfv.randomPolys(g, seeds, l)
        for b in range(3):
            assert np.array_equal(got[b], drbg.random_poly(n, o.q[:l], seeds[b].tobytes())), (l, b)
    poly0 = orc.fill_uniform(3, o.q, n, 3 * L).reshape(3, L, n)
    packed = np.stack([np.frombuffer(opir.serialize_poly(n, o.q, poly0[b]), dtype=np.uint8) for b in range(3)])
    cts = hecuda.Bfv.expandSeeded(g, packed, seeds)
    assert cts.shape == (3, 2, L, n)
    for b in range(3):
        expected = drbg.expand_seeded_ciphertext(o, packed[b].tobytes(), seeds[b].tobytes())
        assert np.array_equal(cts[b], expected), b
    with pytest.raises(hecuda.HeError):   # serializedBufferSizeMismatch
        hecuda.Bfv.expandSeeded(g, packed[:, :-1], seeds)
    g.close()
