"""Scheme-level pins for the oracle: decrypt-correctness (the reference's only scheme-level check,
Sources/_TestUtilities/HeApiTestUtils.swift:494-557, HeAPITests.swift:118-142) plus bit-exact agreement with the
independent big-integer model in oracle/bigint_model.py."""
import random

import numpy as np
import pytest

from oracle import bigint_model as bm
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


def make_ctx(n, bits, nmod, t_bits=None, t=None):
    moduli = orc.generate_primes([bits] * nmod, False, n)
    if t is None:
        (t,) = orc.generate_primes([t_bits], True, n)
    return orc.Context(n, moduli, t)


def test_derived_check_values():
    """SURVEY.md section 8(c) 'derived check values' (first three N=8192 moduli equal the reference's predefined
    n_8192_logq_3x55 set, EncryptionParameters.swift:406-410; t = 557057 matches :383)."""
    q = orc.generate_primes([55] * 4, False, 8192)
    assert q == [36028797018652673, 36028797017571329, 36028797017456641, 36028797017276417]
    ctx = orc.Context(8192, q, 557057)
    assert ctx.bsk == [1152921504606994433, 1152921504607191041, 1152921504607223809, 1152921504607338497]
    assert orc.generate_primes([20], True, 8192) == [557057]
    q16 = orc.generate_primes([55] * 8, False, 16384)
    assert q16[0] == 36028797017456641 and q16[-1] == 36028797013098497


@pytest.mark.parametrize("n,bits,nmod", [(16, 36, 4), (16, 55, 3), (32, 40, 2)])
def test_mul_matches_bigint_model(n, bits, nmod):
    ctx = make_ctx(n, bits, nmod, t_bits=12)
    L = ctx.L
    a = orc.fill_uniform(1, ctx.q, n, 2 * L).reshape(2, L, n)
    b = orc.fill_uniform(2, ctx.q, n, 2 * L).reshape(2, L, n)
    got = ctx.mul(a[None], b[None])[0]
    expect = bm.bfv_mul(a.tolist(), b.tolist(), ctx.q, ctx.bsk, ctx.t)
    assert got.tolist() == expect


@pytest.mark.parametrize("n,bits,nmod", [(16, 36, 4), (16, 55, 3)])
def test_keyswitch_matches_bigint_model(n, bits, nmod):
    ctx = make_ctx(n, bits, nmod, t_bits=12)
    L = ctx.L
    sk, rk = ctx.keygen(5)
    for l in range(1, L + 1):
        target = orc.fill_uniform(3 + l, ctx.q, n, l)
        got = ctx.keyswitch_update(target, rk)
        expect = bm.keyswitch_update(target.tolist(), ctx.moduli, l, rk.tolist())
        assert got.tolist() == expect


@pytest.mark.parametrize("n,bits,nmod,t_bits", [(16, 40, 4, 12), (64, 55, 4, 17), (32, 50, 3, 10), (16, 60, 3, 12)])
def test_multiply_relinearize_decrypts(n, bits, nmod, t_bits):
    ctx = make_ctx(n, bits, nmod, t_bits=t_bits)
    t = ctx.t
    rnd = random.Random(n + bits)
    sk, rk = ctx.keygen(11)
    m1 = [rnd.randrange(t) for _ in range(n)]
    m2 = [rnd.randrange(t) for _ in range(n)]
    c1 = ctx.encrypt(1, sk, m1)
    c2 = ctx.encrypt(2, sk, m2)
    assert ctx.decrypt(sk, c1).tolist() == m1
    assert ctx.decrypt(sk, c2).tolist() == m2
    prod3 = ctx.mul(c1[None], c2[None])[0]
    expect = negacyclic_mul(m1, m2, t)
    assert ctx.decrypt(sk, prod3).tolist() == expect  # 3-poly decrypt (Bfv+Decrypt.swift:188-204)
    relin = ctx.relinearize(prod3[None], rk)[0]
    assert ctx.decrypt(sk, relin).tolist() == expect
    if ctx.L >= 2:
        down = ctx.mod_switch_down(relin[None])[0]
        assert down.shape == (2, ctx.L - 1, n)
        assert ctx.decrypt(sk, down).tolist() == expect
        # relinearize after mod-switch (key switching below the top level, keySwitchingContexts[l-1])
        down3 = ctx.mod_switch_down(prod3[None])[0]
        relin_low = ctx.relinearize(down3[None], rk)[0]
        assert ctx.decrypt(sk, relin_low).tolist() == expect


def test_batch_and_threads_agree():
    ctx = make_ctx(64, 55, 4, t=557057 if False else None, t_bits=17)
    L, n = ctx.L, 64
    a = orc.fill_uniform(21, ctx.q, n, 8 * 2 * L).reshape(8, 2, L, n)
    b = orc.fill_uniform(22, ctx.q, n, 8 * 2 * L).reshape(8, 2, L, n)
    one = ctx.mul(a, b, threads=1)
    many = ctx.mul(a, b, threads=4)
    assert np.array_equal(one, many)
    for k in (0, 7):
        assert np.array_equal(ctx.mul(a[k][None], b[k][None])[0], one[k])
