"""GPU parity for the lazy ciphertext x plaintext inner product and Plaintext.convertToEvalFormat
against the oracle, including nil plaintexts, a single term, and decrypt-correctness of the GPU result."""
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import hecuda
from oracle import oracle as orc


@pytest.mark.parametrize("n,nmod,terms,rows,polys", [(16, 3, 5, 3, 2), (64, 4, 1, 2, 2), (4096, 4, 9, 4, 2),
                                                     (8192, 4, 33, 3, 2), (8192, 3, 4, 2, 3)])
def test_inner_product_matches_oracle(n, nmod, terms, rows, polys):
    moduli = orc.generate_primes([55] * nmod, False, n)
    g, o = hecuda.Context(n, moduli, 65537), orc.Context(n, moduli, 65537)
    L = o.L
    for l in sorted({L, 1}):
        q = moduli[:l]
        cts = orc.fill_uniform(1, q, n, terms * polys * l).reshape(terms, polys, l, n)
        pts = orc.fill_uniform(2, q, n, rows * terms * l).reshape(rows, terms, l, n)
        for i in range(l):  # extreme residues
            cts[0, 0, i, :4] = q[i] - 1
            pts[0, 0, i, :4] = q[i] - 1
        assert np.array_equal(hecuda.Bfv.innerProduct(g, cts, pts), o.inner_product_plain(cts, pts))
        present = np.ones((rows, terms), dtype=np.uint8)
        present[0, terms // 2] = 0
        present[-1, :] = 0  # every plaintext nil -> zero ciphertext
        got = hecuda.Bfv.innerProduct(g, cts, pts, present)
        assert np.array_equal(got, o.inner_product_plain(cts, pts, present))
        assert not got[-1].any()
    g.close()


def test_plaintext_to_eval_and_decrypt_correctness():
    n = 4096
    moduli = orc.generate_primes([55] * 4, False, n)
    t = orc.generate_primes([17], True, 1)[0]
    g, o = hecuda.Context(n, moduli, t), orc.Context(n, moduli, t)
    L = o.L
    rnd = random.Random(8)
    sk, _ = o.keygen(2, relin=False)
    terms = 3
    ms = np.array([[rnd.randrange(t) for _ in range(n)] for _ in range(terms)], dtype=np.uint64)
    ps = np.zeros((terms, n), dtype=np.uint64)
    for k in range(terms):  # sparse plaintexts: the expected product is cheap to compute exactly
        ps[k, 0], ps[k, 1 + k] = 1 + k, t - 2
    pts = hecuda.Bfv.plaintextToEval(g, ps)
    assert np.array_equal(pts, np.stack([o.plaintext_to_eval(ps[k]) for k in range(terms)]))
    assert np.array_equal(hecuda.Bfv.plaintextToEval(g, ps, 2), np.stack([o.plaintext_to_eval(ps[k], 2) for k in range(terms)]))
    cts = hecuda.Bfv.forwardNtt(g, np.stack([o.encrypt(5 + k, sk, ms[k]) for k in range(terms)]))
    out = hecuda.Bfv.innerProduct(g, cts, pts[None])[0]
    coeff = hecuda.Bfv.inverseNtt(g, out)
    expect = [0] * n
    for k in range(terms):
        mk = [int(v) for v in ms[k]]
        for j, c in ((0, 1 + k), (1 + k, t - 2)):
            for i in range(n):
                idx = i + j
                if idx < n:
                    expect[idx] = (expect[idx] + mk[i] * c) % t
                else:
                    expect[idx - n] = (expect[idx - n] - mk[i] * c) % t
    assert o.decrypt(sk, coeff).tolist() == expect
    g.close()


def test_inner_product_errors():
    n = 64
    moduli = orc.generate_primes([50] * 3, False, n)
    g = hecuda.Context(n, moduli, 257)
    with pytest.raises(hecuda.HeError):  # too many polys
        hecuda.Bfv.innerProduct(g, np.zeros((2, 4, 2, n), dtype=np.uint64), np.zeros((1, 2, 2, n), dtype=np.uint64))
    with pytest.raises(hecuda.HeError):  # moduli count above the ciphertext context
        hecuda.Bfv.innerProduct(g, np.zeros((2, 2, 3, n), dtype=np.uint64), np.zeros((1, 2, 3, n), dtype=np.uint64))
    g.close()


@pytest.mark.parametrize("n,nmod,pairs,groups", [(16, 3, 4, 2), (64, 4, 1, 3), (4096, 4, 8, 2), (8192, 4, 5, 2)])
def test_ct_ct_inner_product_matches_oracle(n, nmod, pairs, groups):
    """Bfv.innerProduct(_:_:) (Bfv.swift:315-361); RlweBenchmark uses 8 ct x ct terms (RlweBenchmark.swift:816-820)."""
    moduli = orc.generate_primes([55] * nmod, False, n)
    t = orc.generate_primes([12], True, 1)[0]
    g, o = hecuda.Context(n, moduli, t), orc.Context(n, moduli, t)
    L = o.L
    lhs = orc.fill_uniform(3, moduli[:L], n, groups * pairs * 2 * L).reshape(groups, pairs, 2, L, n)
    rhs = orc.fill_uniform(4, moduli[:L], n, groups * pairs * 2 * L).reshape(groups, pairs, 2, L, n)
    got = hecuda.Bfv.innerProductCiphertexts(g, lhs, rhs)
    assert np.array_equal(got, o.inner_product(lhs, rhs))
    if pairs == 1:
        assert np.array_equal(got, hecuda.Bfv.mulAssign(g, lhs[:, 0], rhs[:, 0]))
    g.close()
