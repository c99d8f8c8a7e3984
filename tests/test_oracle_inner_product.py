"""Oracle pins for the lazy ct x pt inner product: exact big-integer definition and
decrypt-correctness (the reference checks innerProduct by decryption, HeApiTestUtils.swift schemeCiphertextPlaintextInnerProductTest)."""
import random

import numpy as np

from oracle import oracle as orc


def negacyclic_mul(a, b, t):
    n = len(a)
    out = [0] * n
    for i in range(n):
        for j in range(n):
            k = i + j
            if k < n:
                out[k] = (out[k] + a[i] * b[j]) % t
            else:
                out[k - n] = (out[k - n] - a[i] * b[j]) % t
    return out


def test_inner_product_is_exact_sum_mod_q():
    n = 32
    moduli = orc.generate_primes([55, 55, 55], False, n)
    ctx = orc.Context(n, moduli, 65537)
    L, terms, outs = ctx.L, 5, 3
    cts = orc.fill_uniform(1, ctx.q, n, terms * 2 * L).reshape(terms, 2, L, n)
    pts = orc.fill_uniform(2, ctx.q, n, outs * terms * L).reshape(outs, terms, L, n)
    present = np.ones((outs, terms), dtype=np.uint8)
    present[1, 2] = 0
    present[2, :] = 0  # all nil -> zero ciphertext
    got = ctx.inner_product_plain(cts, pts, present)
    for o in range(outs):
        for p in range(2):
            for r in range(L):
                q = ctx.q[r]
                expect = [sum(int(cts[k, p, r, c]) * int(pts[o, k, r, c]) for k in range(terms) if present[o, k]) % q
                          for c in range(n)]
                assert [int(v) for v in got[o, p, r]] == expect


def test_plaintext_to_eval_and_inner_product_decrypts():
    n = 32
    moduli = orc.generate_primes([55, 55, 55], False, n)
    t = orc.generate_primes([12], True, 1)[0]
    ctx = orc.Context(n, moduli, t)
    L = ctx.L
    rnd = random.Random(4)
    sk, _ = ctx.keygen(9, relin=False)
    terms = 4
    ms = [[rnd.randrange(t) for _ in range(n)] for _ in range(terms)]
    ps = [[rnd.randrange(t) if i < 3 else 0 for i in range(n)] for _ in range(terms)]
    cts = np.stack([orc.ntt_forward(n, ctx.q, ctx.encrypt(10 + k, sk, ms[k]).reshape(2 * L, n)).reshape(2, L, n)
                    for k in range(terms)])
    pts = np.stack([ctx.plaintext_to_eval(ps[k]) for k in range(terms)])
    # centered lift: residues are p or p + q_i - t (Plaintext.swift:160-166)
    back = orc.ntt_inverse(n, ctx.q, pts[0])
    for r in range(L):
        for c in range(n):
            v, p = int(back[r, c]), ps[0][c]
            assert v == (p if p < (t + 1) // 2 else p + ctx.q[r] - t)
    out = ctx.inner_product_plain(cts, pts[None])[0]
    coeff = orc.ntt_inverse(n, ctx.q, out.reshape(2 * L, n)).reshape(2, L, n)
    expect = [0] * n
    for k in range(terms):
        prod = negacyclic_mul(ms[k], ps[k], t)
        expect = [(a + b) % t for a, b in zip(expect, prod)]
    assert ctx.decrypt(sk, coeff).tolist() == expect


def test_ct_ct_inner_product_decrypts_and_matches_single_multiply():
    n = 32
    moduli = orc.generate_primes([55, 55, 55, 55], False, n)
    t = orc.generate_primes([10], True, 1)[0]
    ctx = orc.Context(n, moduli, t)
    L = ctx.L
    rnd = random.Random(6)
    sk, _ = ctx.keygen(1, relin=False)
    pairs = 3
    ms = [[rnd.randrange(t) for _ in range(n)] for _ in range(2 * pairs)]
    lhs = np.stack([ctx.encrypt(10 + k, sk, ms[2 * k]) for k in range(pairs)])[None]
    rhs = np.stack([ctx.encrypt(20 + k, sk, ms[2 * k + 1]) for k in range(pairs)])[None]
    out = ctx.inner_product(lhs, rhs)[0]
    expect = [0] * n
    for k in range(pairs):
        expect = [(a + b) % t for a, b in zip(expect, negacyclic_mul(ms[2 * k], ms[2 * k + 1], t))]
    assert ctx.decrypt(sk, out).tolist() == expect
    # with a single pair the inner product IS the multiply (same tensor, same dropExtendedBase)
    one = ctx.inner_product(lhs[:, :1], rhs[:, :1])[0]
    assert np.array_equal(one, ctx.mul(lhs[0, :1], rhs[0, :1])[0])
