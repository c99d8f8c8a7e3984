"""Pins the CPU oracle against the reference's own known-answer tests and big-integer properties.

Each test names the reference test it transcribes (paths relative to /root/reference/Tests/HomomorphicEncryptionTests).
Big-integer ground truth uses Python ints where the reference uses OctoWidth<T>.
"""
import random
from math import prod

import numpy as np
import pytest

from oracle import oracle as orc


def crt_decompose(x, moduli):
    return [x % m for m in moduli]


# ---------------------------------------------------------------- ScalarTests.swift
def test_generate_primes_kats():  # ScalarTests.swift:173-210
    assert orc.generate_primes([62], False) == [4_611_686_018_427_387_847]
    assert orc.generate_primes([61], False) == [2_305_843_009_213_693_951]
    assert orc.generate_primes([60], True) == [576_460_752_303_423_619]
    assert orc.generate_primes([60], False) == [1_152_921_504_606_846_883]
    assert orc.generate_primes([45, 46, 46], False) == [35_184_372_088_777, 70_368_744_177_643, 70_368_744_177_607]
    assert orc.generate_primes([45, 46, 45, 46, 45], True) == [
        17_592_186_044_423, 35_184_372_088_891, 17_592_186_044_437, 35_184_372_088_907, 17_592_186_044_443]
    # UInt32 cases: same algorithm, values fit in 64 bits
    assert orc.generate_primes([27, 28, 28], True, 1024) == [67_127_297, 134_246_401, 134_250_497]
    assert orc.generate_primes([30], False, 2048) == [1_073_692_673]
    with pytest.raises(ValueError):
        orc.generate_primes([5], True, 1024)


def test_is_prime():  # ScalarTests.swift:157-170
    primes = [2, 3, 5, (1 << 14) - 65, (1 << 15) - 49, (1 << 16) - 17, (1 << 28) - 183, (1 << 29) - 3]
    for p in primes:
        assert orc.is_prime(p)
    assert not orc.is_prime(1)
    rnd = random.Random(1)
    for a in primes:
        assert not orc.is_prime(a * rnd.choice(primes))
    assert sum(orc.is_prime(i) for i in range(1, 1000)) == 168


def test_reverse_bits():  # ScalarTests.swift:245-254
    assert orc.reverse_bits(0, 1) == 0
    assert orc.reverse_bits(1 << 31, 32) == 1
    assert orc.reverse_bits(0xFFFFFFFF, 32) == 0xFFFFFFFF
    assert orc.reverse_bits(0xFF00F00F, 32) == 0xF00F00FF
    assert orc.reverse_bits(0xFF00, 16) == 0x00FF


def test_barrett_shoup_random():  # ScalarTests.swift:213-405 (randomised equality vs slow %)
    L = orc.lib()
    rnd = random.Random(7)
    for _ in range(2000):
        bits = rnd.randint(2, 62)
        p = rnd.randrange(1 << (bits - 1), 1 << bits) | 1
        if p >= (1 << 62):
            continue
        x = rnd.randrange(1 << 64)
        assert L.orc_barrett_reduce_single(x, p) == x % p
        hi, lo = rnd.randrange(1 << 64), rnd.randrange(1 << 64)
        assert L.orc_barrett_reduce_double(hi, lo, p) == ((hi << 64) | lo) % p
        a, b = rnd.randrange(p), rnd.randrange(p)
        assert L.orc_barrett_reduce_product(a, b, p) == a * b % p
        assert L.orc_shoup_mul(x, a, p) == x * a % p
        lz = L.orc_shoup_mul_lazy(x, a, p)
        assert lz < 2 * p and lz % p == x * a % p
    # power-of-two modulus (m~ = 2^32 row of the BEHZ base)
    p = 1 << 32
    for _ in range(200):
        x = rnd.randrange(1 << 64)
        a = rnd.randrange(p)
        assert L.orc_shoup_mul(x, a, p) == x * a % p
        hi, lo = rnd.randrange(1 << 64), rnd.randrange(1 << 64)
        assert L.orc_barrett_reduce_double(hi, lo, p) == lo % p


# ---------------------------------------------------------------- NttTests.swift
def test_min_primitive_root():  # NttTests.swift:39-45
    assert orc.min_primitive_root(2, 11) == 10
    assert orc.min_primitive_root(2, 29) == 28
    assert orc.min_primitive_root(4, 29) == 12
    assert orc.min_primitive_root(2, 1_234_565_441) == 1_234_565_440
    assert orc.min_primitive_root(8, 1_234_565_441) == 249_725_733
    # derived check values recorded in SURVEY.md section 8(c)
    assert orc.min_primitive_root(2 * 4096, 36028797018652673) == 4306037850660
    assert orc.min_primitive_root(2 * 8192, 36028797018652673) == 15372713853695


def run_ntt_test(moduli, coeff, evald):
    n = len(coeff[0])
    assert np.array_equal(orc.ntt_forward(n, moduli, coeff), np.array(evald, dtype=np.uint64))
    assert np.array_equal(orc.ntt_inverse(n, moduli, evald), np.array(coeff, dtype=np.uint64))


def test_ntt_kats():  # NttTests.swift:73-191
    run_ntt_test([97], [[0, 0]], [[0, 0]])
    run_ntt_test([97], [[1, 0]], [[1, 1]])
    run_ntt_test([97], [[1, 2]], [[45, 54]])
    run_ntt_test([113], [[3, 4]], [[63, 56]])
    run_ntt_test([97, 113], [[1, 2], [3, 4]], [[45, 54], [63, 56]])
    run_ntt_test([97], [[1, 0, 0, 0]], [[1, 1, 1, 1]])
    run_ntt_test([97], [[1, 2, 3, 4]], [[30, 7, 64, 0]])
    run_ntt_test([97, 113], [[1, 2, 3, 4], [5, 6, 7, 8]], [[30, 7, 64, 0], [108, 31, 103, 4]])
    run_ntt_test([4_194_353], [[1, 0, 0, 0, 0, 0, 0, 0]], [[1] * 8])
    e8 = [3_372_683, 765_982, 387_853, 2_657_954, 2_013_665, 1_280_882, 2_457_874, 3_840_527]
    run_ntt_test([4_194_353], [[1, 2, 3, 4, 5, 6, 7, 8]], [e8])
    run_ntt_test([4_194_353, 113], [[1, 2, 3, 4, 5, 6, 7, 8], [1, 0, 0, 0, 0, 0, 0, 0]], [e8, [1] * 8])
    c16 = [477_051_601, 421_524_611, 456_257_859, 247_136_825, 128_775_020, 76_785_070, 49_764_016, 525_812_772,
           325_605_371, 88_935_943, 255_470_762, 39_507_048, 404_978_219, 379_383_003, 244_420_585, 346_826_612]
    e16 = [230_846_094, 480_599_401, 157_364_576, 360_442_736, 531_052_463, 294_311_347, 432_899_854, 219_721_533,
           286_807_067, 260_650_843, 362_842_688, 315_862_017, 493_042_020, 520_739_674, 167_758_416, 370_401_491]
    run_ntt_test([536_870_849], [[1] + [0] * 15], [[1] * 16])
    run_ntt_test([536_870_849], [c16], [e16])
    c32 = [401, 203, 221, 352, 487, 151, 405, 356, 343, 424, 635, 757, 457, 280, 624, 353,
           496, 353, 624, 280, 457, 757, 635, 424, 343, 356, 405, 151, 487, 352, 221, 203]
    run_ntt_test([769], [c32], [list(range(1, 33))])
    one_hot = [1] + [0] * 4095
    run_ntt_test([557_057], [one_hot], [[1] * 4096])
    run_ntt_test([557_057], [[0] * 4096], [[0] * 4096])


def test_ntt_roundtrip_large_moduli():  # NttTests.swift:194-206
    n = 256
    moduli = orc.generate_primes([60, 62], False, n)
    x = orc.fill_uniform(3, moduli, n, 2)
    assert np.array_equal(orc.ntt_inverse(n, moduli, orc.ntt_forward(n, moduli, x)), x)


def test_ntt_matches_naive():  # NttTests.swift:209-250
    n = 128
    (p,) = orc.generate_primes([30], False, n)
    x = orc.fill_uniform(5, [p], n, 1)
    y = orc.fill_uniform(6, [p], n, 1)
    prod_ntt = orc.ntt_inverse(n, [p], orc.poly_op("mul", n, [p], orc.ntt_forward(n, [p], x), orc.ntt_forward(n, [p], y)))
    xs, ys = [int(v) for v in x[0]], [int(v) for v in y[0]]
    naive = [0] * n
    for i in range(n):
        for j in range(n):
            k = i + j
            if k < n:
                naive[k] = (naive[k] + xs[i] * ys[j]) % p
            else:
                naive[k - n] = (naive[k - n] - xs[i] * ys[j]) % p
    assert [int(v) for v in prod_ntt[0]] == naive


def test_ntt_bit_reversed_evaluation_order():
    """Defines the Eval layout: out[i] = poly(psi^(2*bitrev(i)+1)) with psi the minimal 2N-th root."""
    n, p = 16, 97 * 0 + 4_194_353 - 0
    (p,) = orc.generate_primes([30], False, n)
    psi = orc.min_primitive_root(2 * n, p)
    x = [int(v) for v in orc.fill_uniform(9, [p], n, 1)[0]]
    out = orc.ntt_forward(n, [p], [x])[0]
    for i in range(n):
        e = 2 * orc.reverse_bits(i, 4) + 1
        w = pow(psi, e, p)
        assert int(out[i]) == sum(c * pow(w, k, p) for k, c in enumerate(x)) % p


# ---------------------------------------------------------------- PolyRqTests.swift
def test_divide_and_round_qlast_kats():  # PolyRqTests.swift:146-176
    out = orc.divide_round_qlast(4, [13, 17], [[2, 2, 3, 4], [2, 7, 8, 9]])
    assert out.tolist() == [[0, 2, 2, 3]]
    out = orc.divide_round_qlast(2, [13, 17, 29], [[12, 12], [8, 9], [25, 8]])
    assert out.tolist() == [[1, 10], [1, 10]]


def test_divide_and_round_matches_bigint():
    n = 64
    moduli = orc.generate_primes([50, 55, 55, 57], False, n)
    q = prod(moduli)
    rnd = random.Random(11)
    xs = [rnd.randrange(q) for _ in range(n)]
    data = np.array([crt_decompose(x, moduli) for x in xs], dtype=np.uint64).T.copy()
    out = orc.divide_round_qlast(n, moduli, data)
    ql = moduli[-1]
    for c, x in enumerate(xs):
        expect = (x + (ql >> 1)) // ql
        assert [int(out[r, c]) for r in range(3)] == crt_decompose(expect, moduli[:3])


def test_poly_ops():  # PolyRqTests.swift:46-143 (add / sub / mul semantics)
    n, moduli = 4, [13, 17]
    a, b = [[1, 2, 3, 4], [5, 6, 7, 8]], [[12, 12, 0, 5], [16, 1, 2, 3]]
    for op, f in (("add", lambda x, y, m: (x + y) % m), ("sub", lambda x, y, m: (x - y) % m),
                  ("mul", lambda x, y, m: x * y % m)):
        got = orc.poly_op(op, n, moduli, a, b)
        assert got.tolist() == [[f(x, y, m) for x, y in zip(ra, rb)] for ra, rb, m in zip(a, b, moduli)]


# ---------------------------------------------------------------- RnsBaseConverterTests.swift
@pytest.mark.parametrize("n,bits", [(32, [20, 20]), (16, [30, 30, 30]), (8, [40, 40, 40, 40])])
def test_convert_approximate(n, bits):  # RnsBaseConverterTests.swift:21-65
    q = orc.generate_primes(bits, True)
    tm = orc.generate_primes([bits[0] + 3] * 2, True)
    Q = prod(q)
    rnd = random.Random(13)
    xs = [rnd.randrange(Q) for _ in range(n)]
    data = np.array([crt_decompose(x, q) for x in xs], dtype=np.uint64).T.copy()
    out = orc.convert_approximate(n, q, tm, data)
    for c, x in enumerate(xs):
        got = [int(out[j, c]) for j in range(len(tm))]
        assert any(got == crt_decompose(x + a * Q, tm) for a in range(len(q)))
        # exact value: sum_i [x_i (Q/q_i)^-1]_{q_i} (Q/q_i) mod t_j (RnsBaseConverter.swift:117-143)
        s = sum((x % qi) * pow(Q // qi, -1, qi) % qi * (Q // qi) for qi in q)
        assert got == crt_decompose(s, tm)


# ---------------------------------------------------------------- RnsToolTests.swift
def test_montgomery_reduce_kat():  # RnsToolTests.swift:119-166
    m = 1 << 32
    (q0,) = orc.generate_primes([36], True)
    rt = orc.RnsTool(2, [q0], 2)
    out = rt.small_montgomery_reduce([[m, 2 * m], [m, 2 * m], [0, 0]])
    assert out.tolist() == [[1, 2], [1, 2]]
    q = orc.generate_primes([36, 36], True)
    rt = orc.RnsTool(2, q, 2)
    out = rt.small_montgomery_reduce([[m, 2 * m]] * 3 + [[0, 0]])
    assert out.tolist() == [[1, 2]] * 3


@pytest.mark.parametrize("n,bits", [(4, [20, 20]), (8, [30, 30, 30]), (16, [40, 40, 40, 40]), (8, [55, 55, 55])])
def test_lift_q_to_qbsk(n, bits):  # RnsToolTests.swift:169-208: lift == CRT of the centered value, exactly
    q = orc.generate_primes(bits, True)
    rt = orc.RnsTool(n, q, 2)
    Q = prod(q)
    qbsk = q + rt.bsk
    QB = prod(qbsk)
    rnd = random.Random(17)
    xs = [rnd.randrange(Q) for _ in range(n)]
    data = np.array([crt_decompose(x, q) for x in xs], dtype=np.uint64).T.copy()
    out = rt.lift(data)
    for c, x in enumerate(xs):
        expected = QB - (Q - x) if x > Q // 2 else x
        assert [int(out[r, c]) for r in range(len(qbsk))] == crt_decompose(expected, qbsk)


@pytest.mark.parametrize("n,bits", [(32, [20, 20]), (16, [30, 30, 30]), (8, [40, 40, 40, 40])])
def test_convert_bsk_mtilde(n, bits):  # RnsToolTests.swift:66-117
    q = orc.generate_primes(bits, True)
    rt = orc.RnsTool(n, q, 2)
    Q = prod(q)
    base = rt.bsk + [1 << 32]
    rnd = random.Random(19)
    xs = [rnd.randrange(Q) for _ in range(n)]
    data = np.array([crt_decompose(x, q) for x in xs], dtype=np.uint64).T.copy()
    out = rt.convert_bsk_mtilde(data)
    for c, x in enumerate(xs):
        got = [int(out[r, c]) for r in range(len(base))]
        assert any(got == crt_decompose((x << 32) % Q + a * Q, base) for a in range(len(q)))


@pytest.mark.parametrize("n,bits", [(4, [20, 20]), (8, [30, 30, 30]), (16, [40, 40, 40, 40])])
def test_approximate_floor(n, bits):  # RnsToolTests.swift:211-260: floor within +-(L-1)
    q = orc.generate_primes(bits, True)
    rt = orc.RnsTool(n, q, 2)
    Q, BSK = prod(q), prod(rt.bsk)
    qbsk = q + rt.bsk
    QB = Q * BSK
    rnd = random.Random(23)
    xs = [QB - 1, 1] + [rnd.randrange(QB) for _ in range(n - 2)]
    data = np.array([crt_decompose(x, qbsk) for x in xs], dtype=np.uint64).T.copy()
    out = rt.approximate_floor(data)
    for c, x in enumerate(xs):
        got = [int(out[r, c]) for r in range(len(rt.bsk))]
        cands = []
        for a in range(len(q)):
            cands += [(x // Q + a) % BSK, (x // Q + BSK - a) % BSK]
        assert any(got == crt_decompose(v, rt.bsk) for v in cands)


@pytest.mark.parametrize("n,bits", [(4, [20, 20]), (8, [30, 30, 30]), (8, [55, 55, 55])])
def test_convert_bsk_to_q(n, bits):  # RnsToolTests.swift:263-305: exact
    q = orc.generate_primes(bits, True)
    rt = orc.RnsTool(n, q, 2)
    Q, BSK = prod(q), prod(rt.bsk)
    rnd = random.Random(29)
    xs = [rnd.randrange(Q) for _ in range(n)]
    data = np.array([crt_decompose(x, rt.bsk) for x in xs], dtype=np.uint64).T.copy()
    out = rt.bsk_to_q(data)
    for c, x in enumerate(xs):
        expected = Q - ((BSK - x) % Q) if x > BSK // 2 else x % Q
        assert [int(out[r, c]) for r in range(len(q))] == crt_decompose(expected % Q, q)


@pytest.mark.parametrize("n,bits,t", [(8, [30, 30, 30], 1153), (16, [55, 55], 557057)])
def test_scale_and_round(n, bits, t):  # RnsToolTests.swift:21-64: recovers m from Delta*m + small noise
    q = orc.generate_primes(bits, True)
    rt = orc.RnsTool(n, q, t)
    Q = prod(q)
    rnd = random.Random(31)
    ms = [rnd.randrange(t) for _ in range(n)]
    xs = [((Q // t) * m + rnd.randrange(-1000, 1000)) % Q for m in ms]
    data = np.array([crt_decompose(x, q) for x in xs], dtype=np.uint64).T.copy()
    out = rt.scale_and_round(data)
    assert [int(v) for v in out] == ms


# ---------------------------------------------------------------- the same RnsToolTests at the reference's UInt32 word size
# (RnsToolTests.swift runs every test for UInt32 and UInt64: m~ = 2^16, gamma = 2^30 - 20405, 29-bit Bsk,
#  ModularArithmetic/Scalar.swift:498-511)
def test_word32_constants():
    q = orc.generate_primes([27, 28], True, 8)
    rt = orc.RnsTool(8, q, 2, word_bits=32)
    assert rt.bsk == orc.generate_primes([29, 29, 29], True, 8)  # RnsTool.swift:30-33 with T.bitWidth - 3 = 29
    assert all(1 << 28 <= b < 1 << 29 for b in rt.bsk)
    with pytest.raises(ValueError):
        orc.RnsTool(8, orc.generate_primes([36], True, 8), 2, word_bits=32)  # above Modulus<UInt32>.max


def test_montgomery_reduce_kat_word32():  # RnsToolTests.swift:119-166 with T = UInt32
    m = 1 << 16
    (q0,) = orc.generate_primes([28], True)
    rt = orc.RnsTool(2, [q0], 2, word_bits=32)
    assert rt.small_montgomery_reduce([[m, 2 * m], [m, 2 * m], [0, 0]]).tolist() == [[1, 2], [1, 2]]


@pytest.mark.parametrize("n,bits", [(4, [20, 20]), (8, [27, 28, 28]), (16, [30, 30, 30, 30])])
def test_lift_q_to_qbsk_word32(n, bits):  # RnsToolTests.swift:169-208
    q = orc.generate_primes(bits, True)
    rt = orc.RnsTool(n, q, 2, word_bits=32)
    Q, qbsk = prod(q), q + rt.bsk
    QB = prod(qbsk)
    rnd = random.Random(37)
    xs = [rnd.randrange(Q) for _ in range(n)]
    data = np.array([crt_decompose(x, q) for x in xs], dtype=np.uint64).T.copy()
    out = rt.lift(data)
    for c, x in enumerate(xs):
        expected = QB - (Q - x) if x > Q // 2 else x
        assert [int(out[r, c]) for r in range(len(qbsk))] == crt_decompose(expected, qbsk)


@pytest.mark.parametrize("n,bits", [(4, [20, 20]), (8, [27, 28, 28])])
def test_floor_stages_word32(n, bits):  # RnsToolTests.swift:211-305
    q = orc.generate_primes(bits, True)
    rt = orc.RnsTool(n, q, 2, word_bits=32)
    Q, BSK = prod(q), prod(rt.bsk)
    qbsk = q + rt.bsk
    rnd = random.Random(41)
    xs = [Q * BSK - 1, 1] + [rnd.randrange(Q * BSK) for _ in range(n - 2)]
    out = rt.approximate_floor(np.array([crt_decompose(x, qbsk) for x in xs], dtype=np.uint64).T.copy())
    for c, x in enumerate(xs):
        got = [int(out[r, c]) for r in range(len(rt.bsk))]
        cands = []
        for a in range(len(q)):
            cands += [(x // Q + a) % BSK, (x // Q + BSK - a) % BSK]
        assert any(got == crt_decompose(v, rt.bsk) for v in cands)
    ys = [rnd.randrange(Q) for _ in range(n)]
    out = rt.bsk_to_q(np.array([crt_decompose(y, rt.bsk) for y in ys], dtype=np.uint64).T.copy())
    for c, y in enumerate(ys):
        expected = Q - ((BSK - y) % Q) if y > BSK // 2 else y % Q
        assert [int(out[r, c]) for r in range(len(q))] == crt_decompose(expected % Q, q)


def test_scale_and_round_word32():  # RnsToolTests.swift:21-64
    n, t = 8, 257
    q = orc.generate_primes([28, 28, 29], True)
    rt = orc.RnsTool(n, q, t, word_bits=32)
    Q = prod(q)
    rnd = random.Random(43)
    ms = [rnd.randrange(t) for _ in range(n)]
    xs = [((Q // t) * m + rnd.randrange(-1000, 1000)) % Q for m in ms]
    out = rt.scale_and_round(np.array([crt_decompose(x, q) for x in xs], dtype=np.uint64).T.copy())
    assert [int(v) for v in out] == ms


def test_multiply_relinearize_decrypts_word32():  # HeApiTestUtils.swift:494-557 run for Bfv<UInt32> (HeAPITests.swift:222-230)
    n = 64
    moduli = orc.generate_primes([28, 28, 29], False, n)
    t = orc.generate_primes([10], True, n)[0]
    o = orc.Context(n, moduli, t, word_bits=32)
    sk, rk = o.keygen(7)
    rnd = np.random.default_rng(3)
    m1, m2 = rnd.integers(0, t, n, dtype=np.uint64), rnd.integers(0, t, n, dtype=np.uint64)
    prod3 = o.mul(o.encrypt(1, sk, m1)[None], o.encrypt(2, sk, m2)[None])
    expect = [0] * n
    for i in range(n):
        for j in range(n):
            v = int(m1[i]) * int(m2[j])
            if i + j < n:
                expect[i + j] = (expect[i + j] + v) % t
            else:
                expect[i + j - n] = (expect[i + j - n] - v) % t
    relin = o.relinearize(prod3, rk)
    assert o.decrypt(sk, prod3[0]).tolist() == expect
    assert o.decrypt(sk, relin[0]).tolist() == expect
    assert o.decrypt(sk, o.mod_switch_down(relin)[0]).tolist() == expect
