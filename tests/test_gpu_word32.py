"""Bfv<UInt32> on the GPU (the reference's second scalar type, Scalar.swift:498-511; HeAPITests.swift:222-230 and
RlweBenchmark.swift:732-847 run every test / benchmark at both widths): uint32 buffers through the C ABI against the
oracle in its 32-bit mode (m~ = 2^16, gamma = 2^30 - 20405, 29-bit Bsk), at the reference's 32-bit parameter sets."""
import numpy as np
import pytest

import hecuda
from oracle import oracle as orc

pytestmark = pytest.mark.gpu

PIR_MODULI = [134176769, 268369921, 268361729]  # n_4096_logq_27_28_28 (EncryptionParameters.swift:357-367)
CASES = [(64, None, 641), (1024, None, 12289), (4096, PIR_MODULI, 17), (8192, None, 65537)]


def setup(n, moduli, t):
    moduli = moduli or orc.generate_primes([28, 28, 29], False, n)
    return moduli, hecuda.Context(n, moduli, t, scalar=np.uint32), orc.Context(n, moduli, t, word_bits=32)


def test_context_is_32_bit():
    moduli, g, o = setup(64, None, 641)
    assert g.bskModuli == orc.RnsTool(64, moduli[:2], 641, word_bits=32).bsk
    assert all(1 << 28 <= b < 1 << 29 for b in g.bskModuli)
    with pytest.raises(hecuda.HeError):
        hecuda.Context(64, orc.generate_primes([40, 40], False, 64), 641, scalar=np.uint32)


@pytest.mark.parametrize("n,moduli,t", CASES)
def test_ntt_u32_matches_oracle(n, moduli, t):
    moduli, g, o = setup(n, moduli, t)
    L = len(moduli) - 1
    x = orc.fill_uniform(3, moduli[:L], n, 4 * L).reshape(4, L, n)
    x[0, :, :3] = [[0, 1, m - 1] for m in moduli[:L]]
    fwd = hecuda.Bfv32.forwardNtt(g, x.astype(np.uint32))
    assert fwd.dtype == np.uint32
    assert np.array_equal(fwd.astype(np.uint64).reshape(-1, n), orc.ntt_forward(n, moduli[:L], x))
    assert np.array_equal(hecuda.Bfv32.inverseNtt(g, fwd).astype(np.uint64), x)


@pytest.mark.parametrize("n,moduli,t", CASES)
def test_multiply_relinearize_modswitch_u32_match_oracle(n, moduli, t):
    moduli, g, o = setup(n, moduli, t)
    L = len(moduli) - 1
    batch = 3
    a = orc.fill_uniform(11, moduli[:L], n, batch * 2 * L).reshape(batch, 2, L, n)
    b = orc.fill_uniform(12, moduli[:L], n, batch * 2 * L).reshape(batch, 2, L, n)
    for i in range(L):
        a[0, :, i, :4] = moduli[i] - 1
        b[0, :, i, :4] = [0, 1, moduli[i] - 1, moduli[i] // 2]
    prod = hecuda.Bfv32.mulAssign(g, a.astype(np.uint32), b.astype(np.uint32))
    want = o.mul(a, b)
    assert np.array_equal(prod.astype(np.uint64), want)
    _, rk = o.keygen(5)
    key = hecuda.EvaluationKey32(g, rk.astype(np.uint32))
    relin = hecuda.Bfv32.relinearize(g, prod, key)
    want_relin = o.relinearize(want, rk)
    assert np.array_equal(relin.astype(np.uint64), want_relin)
    down = hecuda.Bfv32.modSwitchDown(g, relin)
    assert np.array_equal(down.astype(np.uint64), o.mod_switch_down(want_relin))
    key.close()


def test_encrypted_product_decrypts_u32():
    """HeApiTestUtils.swift:494-557 at the PIR default parameters in 32-bit: Enc(a) * Enc(b) -> relinearize decrypts to a*b."""
    n, t = 4096, 17
    moduli, g, o = setup(n, PIR_MODULI, t)
    sk, rk = o.keygen(9)
    rng = np.random.default_rng(1)
    m1 = np.zeros(n, dtype=np.uint64)
    m2 = np.zeros(n, dtype=np.uint64)
    m1[:8] = rng.integers(0, t, 8)
    m2[0], m2[1] = 3, 5
    prod = hecuda.Bfv32.mulAssign(g, o.encrypt(1, sk, m1)[None].astype(np.uint32), o.encrypt(2, sk, m2)[None].astype(np.uint32))
    key = hecuda.EvaluationKey32(g, rk.astype(np.uint32))
    relin = hecuda.Bfv32.relinearize(g, prod, key)
    expect = np.zeros(n, dtype=np.int64)
    for j, c in ((0, 3), (1, 5)):
        expect[j:j + 8] += m1[:8].astype(np.int64) * c
    assert o.decrypt(sk, relin[0].astype(np.uint64)).tolist() == (expect % t).tolist()
    key.close()


@pytest.mark.parametrize("n,moduli,t", CASES[:3])
def test_lift_and_floor_u32_match_oracle(n, moduli, t):
    moduli, g, o = setup(n, moduli, t)
    L = len(moduli) - 1
    tool = orc.RnsTool(n, moduli[:L], t, word_bits=32)
    x = orc.fill_uniform(41, moduli[:L], n, 2 * L).reshape(2, L, n)
    lifted = hecuda.Bfv32.liftQToQBsk(g, x.astype(np.uint32))
    assert np.array_equal(lifted.astype(np.uint64), np.stack([tool.lift(x[k]) for k in range(2)]))
    base = moduli[:L] + tool.bsk
    y = orc.fill_uniform(42, base, n, 2 * (2 * L + 1)).reshape(2, 2 * L + 1, n)
    assert np.array_equal(hecuda.Bfv32.floorQBskToQ(g, y.astype(np.uint32)).astype(np.uint64),
                          np.stack([tool.floor(y[k]) for k in range(2)]))


def test_u64_entry_points_refuse_nothing_but_u32_entry_points_need_a_32_bit_context():
    n = 64
    moduli = orc.generate_primes([28, 28, 29], False, n)
    g64 = hecuda.Context(n, moduli, 641)
    x = np.zeros((1, 2, n), dtype=np.uint32)
    with pytest.raises(hecuda.HeError):
        hecuda.Bfv32.forwardNtt(g64, x)


def test_fused_calls_galois_and_inner_product_u32():
    """The remaining uint32 entry points: multiply+relinearize(+modSwitchDown), relinearize+modSwitchDown, applyGalois and
    the ct x ct inner product, against the 32-bit oracle."""
    n, t = 4096, 17
    moduli, g, o = setup(n, PIR_MODULI, t)
    L = len(moduli) - 1
    a = orc.fill_uniform(21, moduli[:L], n, 4 * 2 * L).reshape(4, 2, L, n)
    b = orc.fill_uniform(22, moduli[:L], n, 4 * 2 * L).reshape(4, 2, L, n)
    sk, rk = o.keygen(5)
    key = hecuda.EvaluationKey32(g, rk.astype(np.uint32))
    prod = o.mul(a, b)
    relin = o.relinearize(prod, rk)
    a32, b32 = a.astype(np.uint32), b.astype(np.uint32)
    assert np.array_equal(hecuda.Bfv32.mulRelinearize(g, a32, b32, key).astype(np.uint64), relin)
    assert np.array_equal(hecuda.Bfv32.mulRelinearize(g, a32, b32, key, modSwitchDown=True).astype(np.uint64), o.mod_switch_down(relin))
    assert np.array_equal(hecuda.Bfv32.relinearizeModSwitchDown(g, prod.astype(np.uint32), key).astype(np.uint64),
                          o.mod_switch_down(relin))
    element = 3
    gk = o.galois_keygen(77, sk, element)
    key.setGaloisKey(element, gk.astype(np.uint32))
    want = o.apply_galois(a[:2], element, gk)
    assert np.array_equal(hecuda.Bfv32.applyGalois(g, a32[:2], element, key).astype(np.uint64), want)
    lhs, rhs = a.reshape(2, 2, 2, L, n), b.reshape(2, 2, 2, L, n)
    assert np.array_equal(hecuda.Bfv32.innerProductCiphertexts(g, lhs.astype(np.uint32), rhs.astype(np.uint32)).astype(np.uint64),
                          o.inner_product(lhs, rhs))
    key.close()


def test_mulpir_on_a_32_bit_context():
    """MulPir computeResponse at the reference's default PIR parameters (n_4096_logq_27_28_28, t = 17) with Bfv<UInt32>'s
    constants: the application driver runs on the 32-bit context through the uint64 entry points (64-bit storage of the
    same residues); reply bit-exact with the 32-bit oracle and decrypting to the database entry."""
    import random

    from hecuda import pir
    from oracle import pir_oracle as opir

    n, t = 4096, 17
    moduli, g, o = setup(n, PIR_MODULI, t)
    entries, entry_size = 30000, 1
    rng = random.Random(entries)
    config = pir.IndexPirConfig(entries, entry_size, 2, 1, True, "hybridCompression", False)
    param = pir.MulPir.generateParameter(config, g)
    oparam = opir.generate_parameter(opir.IndexPirConfig(entries, entry_size, 2, 1, True, "hybridCompression", False), n, t)
    database = [bytes(rng.randrange(256) for _ in range(entry_size)) for _ in range(entries)]
    sk, relin = o.keygen(13)
    key = hecuda.EvaluationKey(g, relin)
    okeys = {}
    for i, e in enumerate(param.evaluationKeyConfig.galoisElements):
        okeys[e] = o.galois_keygen(100 + i, sk, e)
        key.setGaloisKey(e, okeys[e])
    server = pir.MulPirServer(param, g, [pir.MulPirServer.process(database, g, param)])
    index = rng.randrange(entries)
    query = opir.generate_query(o, oparam, [index], sk, 900)
    got = server.computeResponse(np.stack(query), key)
    reply = [[got[0, c] for c in range(server.chunkCount)]]
    assert opir.decrypt_response(o, oparam, reply, [index], sk) == [database[index]]
    expected = opir.compute_response(o, query, 1, okeys, relin, [opir.process_database(o, oparam, database)], oparam)
    for c in range(server.chunkCount):
        assert np.array_equal(got[0, c], expected[0][c])
    key.close()
