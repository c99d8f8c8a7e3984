"""GPU parity for the Galois row: PolyRq.applyGalois in both formats and Bfv.applyGalois
(rotations / row swap) against the oracle, which is pinned on the reference's GaloisTests KATs."""
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import hecuda
from oracle import oracle as orc


def primes(bits, count, n):
    return orc.generate_primes([bits] * count, False, n)


def test_poly_apply_galois_reference_kats():
    """GaloisTests.swift:21-60 (element 3) through the GPU path; 17 and 97 are NTT-friendly for N = 4 / 8."""
    for n, moduli, data, expected in [
        (4, [17], [[0, 1, 2, 3]], [[0, 3, 15, 1]]),
        (8, [17], [[0, 1, 2, 3, 4, 5, 6, 7]], [[0, 14, 6, 1, 13, 7, 2, 12]]),
        (8, [17, 97], [[0, 1, 2, 3, 4, 5, 6, 7], [7, 6, 5, 4, 3, 2, 1, 0]],
         [[0, 14, 6, 1, 13, 7, 2, 12], [7, 93, 1, 6, 94, 0, 5, 95]]),
    ]:
        ks = [m for m in orc.generate_primes([30, 30, 30], True, n) if m not in moduli][0]
        g = hecuda.Context(n, moduli + [ks], 2)
        assert hecuda.Bfv.polyApplyGalois(g, np.array(data, dtype=np.uint64), 3).tolist() == expected
        g.close()


@pytest.mark.parametrize("n", [8, 64, 1024, 8192])
def test_poly_apply_galois_matches_oracle(n):
    moduli = primes(55, 4, n)
    g = hecuda.Context(n, moduli, 2)
    q = moduli[:3]
    x = orc.fill_uniform(n, q, n, 2 * 3).reshape(2, 3, n)
    x[0, :, 0] = 0  # negating zero must stay zero
    elements = [3, 2 * n - 1, orc.galois_element_rotating_columns(1, n), orc.galois_element_rotating_columns(-(n // 4), n)]
    for el in elements:
        got = hecuda.Bfv.polyApplyGalois(g, x, el)
        assert np.array_equal(got[0], orc.galois_coeff(n, q, el, x[0]))
        assert np.array_equal(got[1], orc.galois_coeff(n, q, el, x[1]))
        ev = hecuda.Bfv.forwardNtt(g, x)
        got_ev = hecuda.Bfv.polyApplyGalois(g, ev, el, evalFormat=True)
        assert np.array_equal(got_ev[0], orc.galois_eval(n, 3, el, ev[0]))
        # automorphism commutes with the NTT (GaloisTests.swift:66-71)
        assert np.array_equal(hecuda.Bfv.forwardNtt(g, got), got_ev)
    with pytest.raises(hecuda.HeError):
        hecuda.Bfv.polyApplyGalois(g, x, 4)  # even elements are invalid (Galois.swift:100-105)
    g.close()


@pytest.mark.parametrize("n,bits,nmod", [(16, 40, 3), (64, 55, 4), (4096, 55, 4), (8192, 55, 4)])
def test_apply_galois_matches_oracle_and_decrypts(n, bits, nmod):
    moduli = primes(bits, nmod, n)
    t = orc.generate_primes([12], True, 1)[0]
    g, o = hecuda.Context(n, moduli, t), orc.Context(n, moduli, t)
    L = o.L
    rnd = random.Random(n)
    sk, _ = o.keygen(3, relin=False)
    m = np.array([rnd.randrange(t) for _ in range(n)], dtype=np.uint64)
    fresh = o.encrypt(1, sk, m)
    batch = 3
    cts = np.stack([fresh] + [orc.fill_uniform(7 + k, moduli[:L], n, 2 * L).reshape(2, L, n) for k in range(batch - 1)])
    key = hecuda.EvaluationKey(g, None)
    for el in (3, orc.galois_element_rotating_columns(2 if n > 16 else 1, n), orc.galois_element_swapping_rows(n)):
        gk = o.galois_keygen(100 + el, sk, el)
        key.setGaloisKey(el, gk)
        got = hecuda.Bfv.applyGalois(g, cts, el, key)
        assert np.array_equal(got, o.apply_galois(cts, el, gk)), f"element {el}"
        assert o.decrypt(sk, got[0]).tolist() == orc.galois_coeff(n, [t], el, [m])[0].tolist()
        if L >= 2:  # one level down (key-switch context l < L)
            low = o.mod_switch_down(cts)
            assert np.array_equal(hecuda.Bfv.applyGalois(g, low, el, key), o.apply_galois(low, el, gk))
    with pytest.raises(hecuda.HeError):  # missingGaloisElement (Bfv.swift:188-190)
        hecuda.Bfv.applyGalois(g, cts, 5, key)
    with pytest.raises(hecuda.HeError):  # three polys (Bfv.swift:180)
        hecuda.Bfv.applyGalois(g, np.zeros((1, 3, L, n), dtype=np.uint64), 3, key)
    key.close()
    g.close()


@pytest.mark.parametrize("n", [8, 64, 8192])
def test_multiply_power_of_x_matches_oracle(n):
    moduli = primes(55, 4, n)
    g = hecuda.Context(n, moduli, 2)
    q = moduli[:3]
    x = orc.fill_uniform(n + 1, q, n, 2 * 3).reshape(2, 3, n)
    x[0, :, 0] = 0
    for power in (0, 1, -1, n // 2, n, n + 3, -(n + 3), 2 * n, 5 * n + 1, -7 * n - 2):
        got = hecuda.Bfv.multiplyPowerOfX(g, x, power)
        assert np.array_equal(got[0], orc.multiply_power_of_x(n, q, power, x[0])), power
        assert np.array_equal(got[1], orc.multiply_power_of_x(n, q, power, x[1])), power
    g.close()
