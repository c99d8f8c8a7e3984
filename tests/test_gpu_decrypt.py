"""GPU parity for Bfv.decryptCoeff against the oracle's decrypt (pinned on the reference's
scaleAndRound test, RnsToolTests.swift:21-64, and on encrypt/decrypt round trips)."""
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import hecuda
from oracle import oracle as orc


@pytest.mark.parametrize("n,bits,t", [(16, [55, 52, 62, 58], 1153), (64, [55, 55, 55], 65537), (4096, [27, 28, 28], 17),
                                      (8192, [55, 55, 55, 55], 557057)])
def test_decrypt_matches_oracle_at_every_level(n, bits, t):
    moduli = orc.generate_primes(bits, False, n)
    g, o = hecuda.Context(n, moduli, t), orc.Context(n, moduli, t)
    rng = random.Random(n)
    sk, rk = o.keygen(4)
    messages = [np.array([rng.randrange(t) for _ in range(n)], dtype=np.uint64) for _ in range(3)]
    fresh = np.stack([o.encrypt(10 + i, sk, m) for i, m in enumerate(messages)])
    got = hecuda.Bfv.decrypt(g, fresh, sk)
    for i, m in enumerate(messages):
        assert np.array_equal(got[i], m) and np.array_equal(got[i], o.decrypt(sk, fresh[i]))
    # three-poly product (Bfv+Decrypt.swift:188-204 handles any poly count), then relinearized, then every lower level
    product = o.mul(fresh[:1], fresh[1:2])
    assert np.array_equal(hecuda.Bfv.decrypt(g, product, sk)[0], o.decrypt(sk, product[0]))
    ct = o.relinearize(product, rk)
    while True:
        assert np.array_equal(hecuda.Bfv.decrypt(g, ct, sk)[0], o.decrypt(sk, ct[0])), ct.shape
        if ct.shape[-2] == 1:
            break
        ct = o.mod_switch_down(ct)
    # uniformly random "ciphertexts" exercise every branch of the gamma correction
    junk = orc.fill_uniform(5, moduli[: o.L], n, 4 * 2 * o.L).reshape(4, 2, o.L, n)
    got = hecuda.Bfv.decrypt(g, junk, sk)
    for i in range(4):
        assert np.array_equal(got[i], o.decrypt(sk, junk[i]))
    with pytest.raises(hecuda.HeError):
        hecuda.Bfv.decrypt(g, np.zeros((1, 4, o.L, n), dtype=np.uint64), sk)
    g.close()
