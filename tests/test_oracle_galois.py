"""Oracle pins for the Galois row: the reference's KATs and properties
(Tests/HomomorphicEncryptionTests/PolyRqTests/GaloisTests.swift:21-113) and decrypt-correctness of applyGalois."""
import random

import numpy as np
import pytest

from oracle import oracle as orc

KATS = [  # GaloisTests.swift:21-60, element 3
    (4, [17], [[0, 1, 2, 3]], [[0, 3, 15, 1]]),
    (8, [17], [[0, 1, 2, 3, 4, 5, 6, 7]], [[0, 14, 6, 1, 13, 7, 2, 12]]),
    (8, [17, 97], [[0, 1, 2, 3, 4, 5, 6, 7], [7, 6, 5, 4, 3, 2, 1, 0]],
     [[0, 14, 6, 1, 13, 7, 2, 12], [7, 93, 1, 6, 94, 0, 5, 95]]),
]


@pytest.mark.parametrize("n,moduli,data,expected", KATS)
def test_apply_galois_kats(n, moduli, data, expected):
    assert orc.galois_coeff(n, moduli, 3, data).tolist() == expected
    # Eval-format automorphism commutes with the NTT (GaloisTests.swift:66-71)
    ev = orc.ntt_forward(n, moduli, data)
    assert orc.ntt_inverse(n, moduli, orc.galois_eval(n, len(moduli), 3, ev)).tolist() == expected
    for index in range(1, n):
        element = 2 * index + 1
        lhs = orc.ntt_forward(n, moduli, orc.galois_coeff(n, moduli, element, data))
        assert np.array_equal(lhs, orc.galois_eval(n, len(moduli), element, ev))
    fwd = orc.galois_element_swapping_rows(n)
    assert orc.galois_coeff(n, moduli, fwd, orc.galois_coeff(n, moduli, fwd, data)).tolist() == data
    for step in range(1, n // 2):
        f = orc.galois_element_rotating_columns(step, n)
        b = orc.galois_element_rotating_columns(n // 2 - step, n)
        assert orc.galois_coeff(n, moduli, b, orc.galois_coeff(n, moduli, f, data)).tolist() == data


def test_galois_elements():  # GaloisTests.swift:115-121: elements 3, 9, 11 <-> steps 3, 2, 1 at degree 8
    assert orc.galois_element_rotating_columns(3, 8) == 3
    assert orc.galois_element_rotating_columns(2, 8) == 9
    assert orc.galois_element_rotating_columns(1, 8) == 11
    with pytest.raises(ValueError):
        orc.galois_element_rotating_columns(4, 8)


@pytest.mark.parametrize("n,bits,nmod", [(16, 40, 3), (64, 55, 4)])
def test_apply_galois_decrypts_to_automorphism(n, bits, nmod):
    """Dec(applyGalois(Enc(m), g)) = m(x^g): the reference checks this through rotateColumns/swapRows on SIMD slots
    (HeApiTestUtils.swift schemeRotationTest); with coefficient encoding the same identity is the polynomial one."""
    moduli = orc.generate_primes([bits] * nmod, False, n)
    t = orc.generate_primes([12], True, 1)[0]
    ctx = orc.Context(n, moduli, t)
    rnd = random.Random(n)
    sk, _ = ctx.keygen(3, relin=False)
    m = np.array([rnd.randrange(t) for _ in range(n)], dtype=np.uint64)
    ct = ctx.encrypt(1, sk, m)
    for element in (3, orc.galois_element_rotating_columns(1, n), orc.galois_element_swapping_rows(n)):
        gk = ctx.galois_keygen(50 + element, sk, element)
        out = ctx.apply_galois(ct[None], element, gk)[0]
        expect = orc.galois_coeff(n, [t], element, [m])[0]
        assert ctx.decrypt(sk, out).tolist() == expect.tolist()
        if ctx.L >= 2:  # below the top level too
            low = ctx.mod_switch_down(ct[None])
            out_low = ctx.apply_galois(low, element, gk)[0]
            assert ctx.decrypt(sk, out_low).tolist() == expect.tolist()


def test_multiply_power_of_x_kats():  # PolyRqTests.swift:179-244
    moduli, data = [2, 3, 5], [[0, 1, 0, 1], [0, 1, 2, 0], [0, 1, 2, 3]]
    neg = [orc.multiply_power_of_x(4, moduli, -i, data).tolist() for i in range(8)]
    pos = [orc.multiply_power_of_x(4, moduli, i, data).tolist() for i in range(8)]
    assert neg[0] == data and pos[0] == data
    assert neg[1] == [[1, 0, 1, 0], [1, 2, 0, 0], [1, 2, 3, 0]]
    assert neg[2] == [[0, 1, 0, 1], [2, 0, 0, 2], [2, 3, 0, 4]]
    assert neg[3] == [[1, 0, 1, 0], [0, 0, 2, 1], [3, 0, 4, 3]]
    assert pos[1] == [[1, 0, 1, 0], [0, 0, 1, 2], [2, 0, 1, 2]]
    assert pos[2] == [[0, 1, 0, 1], [1, 0, 0, 1], [3, 2, 0, 1]]
    assert pos[3] == [[1, 0, 1, 0], [2, 1, 0, 0], [4, 3, 2, 0]]
    for ys in (neg, pos):  # X^(i+N) = -X^i
        for i in range(4):
            z = orc.poly_op("add", 4, moduli, ys[i], ys[i + 4])
            assert not z.any()
