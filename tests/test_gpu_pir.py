"""GPU parity for the MulPir server row: PirUtil.expand and computeResponse through the C ABI,
bit-exact against oracle/pir_oracle.py (pinned on the reference's MulPir / Expansion / IndexPir tests), plus the
reference's own end-to-end property: the decrypted response is the database entry."""
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import hecuda
from hecuda import pir
from oracle import oracle as orc
from oracle import pir_oracle as opir

TEST_MODULI_BITS = [55, 52, 62, 58]  # TestUtils.testCoefficientModuli for UInt64 (TestUtilities.swift:312-317)


def contexts(n, bits, t):
    moduli = orc.generate_primes(bits, False, n)
    return hecuda.Context(n, moduli, t), orc.Context(n, moduli, t)


def load_keys(g, o, sk, relin, elements, seed=100):
    key = hecuda.EvaluationKey(g, relin)
    okeys = {}
    for i, e in enumerate(elements):
        okeys[e] = o.galois_keygen(seed + i, sk, e)
        key.setGaloisKey(e, okeys[e])
    return key, okeys


@pytest.mark.parametrize("n,bits,t", [(16, TEST_MODULI_BITS, 1153), (64, [55, 55, 55], 65537), (1024, [50, 50, 50], 17)])
@pytest.mark.parametrize("compression", ["noCompression", "hybridCompression", "maxCompression"])
def test_expand_matches_oracle(n, bits, t, compression):
    g, o = contexts(n, bits, t)
    sk, _ = o.keygen(5, relin=False)
    rng = random.Random(n)
    counts = sorted({1, 2, 3, 5, n // 2 + 1, n - 1, n, n + 1, n + 2, 2 * n, 2 * n + 5} if n <= 64 else {37, 200})
    for count in counts:
        # a single output needs no key; its configuration names 2^(logN+1)+1, which is not a valid element
        elements = [e for e in opir.evaluation_key_config(count, n, compression) if e < 2 * n]
        key, okeys = load_keys(g, o, sk, None, elements)
        ones = [i for i in range(count) if rng.random() < 0.3]
        cts = opir.compress_binary_inputs(o, count, ones, sk, 300 + count)
        expected = np.stack(opir.expand(o, cts, count, okeys))
        got = pir.PirUtil.expand(g, np.stack(cts), count, key)
        assert np.array_equal(got, expected), f"outputCount {count}"
        for index in (0, count // 2, count - 1):  # ExpansionTests: constant polynomial with the queried bit
            dec = o.decrypt(sk, got[index])
            assert int(dec[0]) == (1 if index in ones else 0) and not dec[1:].any()
        key.close()
    g.close()


def test_expand_errors():
    g, o = contexts(16, TEST_MODULI_BITS, 1153)
    sk, _ = o.keygen(5, relin=False)
    cts = np.stack(opir.compress_binary_inputs(o, 8, [1], sk, 1))
    key, _ = load_keys(g, o, sk, None, [9])          # only x -> x^9: cannot serve logStep 1 (target 17)... 9 <= 17 applies twice
    pir.PirUtil.expand(g, cts, 2, key)
    key.close()
    key, _ = load_keys(g, o, sk, None, [17])         # logStep 2 needs an element <= 9
    with pytest.raises(hecuda.HeError) as err:
        pir.PirUtil.expand(g, cts, 4, key)
    assert "missingGaloisKey" in str(err.value)
    with pytest.raises(hecuda.HeError):              # outputCount must fit the ciphertext count (PirUtil.swift:326-327)
        pir.PirUtil.expand(g, cts, 17, key)
    key.close()
    g.close()


CONFIGS = [
    dict(entry_size=1, dims=2, uneven=False, compression="noCompression"),
    dict(entry_size=8, dims=2, uneven=False, compression="noCompression"),
    dict(entry_size=24, dims=2, uneven=True, compression="noCompression"),
    dict(entry_size=24, dims=1, uneven=True, compression="noCompression"),
    dict(entry_size=24, dims=1, uneven=True, compression="hybridCompression"),
    dict(entry_size=24, dims=1, uneven=True, compression="maxCompression"),
    dict(entry_size=47, dims=2, uneven=True, compression="hybridCompression"),   # 3 chunks, two dimensions
]


@pytest.mark.parametrize("encoding", [False, True])
@pytest.mark.parametrize("cfg", CONFIGS)
def test_index_pir_matches_oracle_and_decrypts(cfg, encoding):
    """IndexPirTests.indexPirTest configurations on the reference's test context (N=16, t=1153)."""
    g, o = contexts(16, TEST_MODULI_BITS, 1153)
    rng = random.Random(cfg["entry_size"] * 3 + cfg["dims"] + 17 * encoding)
    config = pir.IndexPirConfig(100, cfg["entry_size"], cfg["dims"], 2, cfg["uneven"], cfg["compression"], encoding)
    param = pir.MulPir.generateParameter(config, g)
    oparam = opir.generate_parameter(opir.IndexPirConfig(100, cfg["entry_size"], cfg["dims"], 2, cfg["uneven"],
                                                         cfg["compression"], encoding), o.n, o.t)
    assert param.dimensions == oparam.dimensions
    database = [bytes(rng.randrange(256) for _ in range(rng.randint(1, cfg["entry_size"]) if encoding else cfg["entry_size"]))
                for _ in range(100)]
    sk, relin = o.keygen(31)
    key, okeys = load_keys(g, o, sk, relin, param.evaluationKeyConfig.galoisElements)
    server = pir.MulPirServer(param, g, [pir.MulPirServer.process(database, g, param)])
    odb = opir.process_database(o, oparam, database)
    for trial, batch in enumerate((2, 1)):
        indices = rng.sample(range(100), batch)
        query = opir.generate_query(o, oparam, indices, sk, 500 + 10 * trial)
        expected = opir.compute_response(o, query, batch, okeys, relin, [odb], oparam)
        got = server.computeResponse(np.stack(query), key, indicesCount=batch)
        assert got.shape == (batch, server.chunkCount, 2, 1, 16)
        for qi in range(batch):
            for chunk in range(server.chunkCount):
                assert np.array_equal(got[qi, chunk], expected[qi][chunk]), (qi, chunk)
        reply = [[got[qi, c] for c in range(server.chunkCount)] for qi in range(batch)]
        assert opir.decrypt_response(o, oparam, reply, indices, sk) == [database[i] for i in indices]
    key.close()
    g.close()


def test_index_pir_one_database_per_query_and_errors():
    g, o = contexts(16, TEST_MODULI_BITS, 1153)
    rng = random.Random(4)
    config = pir.IndexPirConfig(40, 4, 2, 2, False, "noCompression", False)
    param = pir.MulPir.generateParameter(config, g)
    oparam = opir.generate_parameter(opir.IndexPirConfig(40, 4, 2, 2, False, "noCompression", False), o.n, o.t)
    dbs = [[bytes(rng.randrange(256) for _ in range(4)) for _ in range(40)] for _ in range(2)]
    sk, relin = o.keygen(8)
    key, okeys = load_keys(g, o, sk, relin, param.evaluationKeyConfig.galoisElements)
    server = pir.MulPirServer(param, g, [pir.MulPirServer.process(d, g, param) for d in dbs])
    indices = [7, 33]
    query = opir.generate_query(o, oparam, indices, sk, 77)
    got = server.computeResponse(np.stack(query), key, indicesCount=2)
    expected = opir.compute_response(o, query, 2, okeys, relin, [opir.process_database(o, oparam, d) for d in dbs], oparam)
    for qi in range(2):
        assert np.array_equal(got[qi, 0], expected[qi][0])
        assert opir.decrypt_response(o, oparam, [[got[qi, 0]]], [indices[qi]], sk) == [dbs[qi][indices[qi]]]
    # PirError.invalidBatchSize: 2 databases cannot serve 3 queries (PirUtil.swift:498-500)
    three = opir.generate_query(o, opir.IndexPirParameter(40, 4, oparam.dimensions, 3), [1, 2, 3], sk, 5)
    with pytest.raises(hecuda.HeError) as err:
        server.computeResponse(np.stack(three), key, indicesCount=3)
    assert "invalidBatchSize" in str(err.value)
    # PirError.invalidDatabasePlaintextCount (MulPir.swift:352-358)
    with pytest.raises(pir.PirError):
        pir.MulPirServer(param, g, [pir.ProcessedDatabase(g, np.zeros((3, 16), dtype=np.uint64))])
    # no relinearization key for a two-dimensional database
    bare = hecuda.EvaluationKey(g, None)
    for e, k in okeys.items():
        bare.setGaloisKey(e, k)
    with pytest.raises(hecuda.HeError) as err:
        server.computeResponse(np.stack(query), bare, indicesCount=2)
    assert "missingRelinearizationKey" in str(err.value)
    bare.close()
    key.close()
    g.close()


@pytest.mark.parametrize("n,bits,t,entries,entry_size,compression", [
    (4096, [27, 28, 28], 17, 30000, 1, "hybridCompression"),      # EncryptionParametersConfig.defaultPir (BenchmarkMetricExtensions.swift:60-63)
    (4096, [27, 28, 28], 17, 300, 3000, "noCompression"),         # entries larger than a plaintext: 2 chunks
    (8192, [55, 55, 55, 55], 65537, 5000, 100, "maxCompression"),
])
def test_index_pir_production_sizes(n, bits, t, entries, entry_size, compression):
    g, o = contexts(n, bits, t)
    rng = random.Random(entries)
    config = pir.IndexPirConfig(entries, entry_size, 2, 1, True, compression, False)
    param = pir.MulPir.generateParameter(config, g)
    oparam = opir.generate_parameter(opir.IndexPirConfig(entries, entry_size, 2, 1, True, compression, False), n, t)
    database = [bytes(rng.randrange(256) for _ in range(entry_size)) for _ in range(entries)]
    sk, relin = o.keygen(13)
    key, okeys = load_keys(g, o, sk, relin, param.evaluationKeyConfig.galoisElements)
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
    g.close()


@pytest.mark.parametrize("n,bits,t,entries,entry_size", [(16, TEST_MODULI_BITS, 1153, 60, 7), (4096, [27, 28, 28], 17, 20000, 3),
                                                         (4096, [27, 28, 28], 17, 200, 3000)])
def test_wire_level_response_matches_oracle_and_decrypts(n, bits, t, entries, entry_size):
    """Serialized seeded query in, serialized skipLSBs reply out (one C-ABI call) == oracle composition; the reply still
    decrypts to the database entry after the dropped bits come back as zeros (Bfv+Decrypt.swift:51-110)."""
    from oracle import drbg_oracle as drbg
    g, o = contexts(n, bits, t)
    rng = random.Random(entries)
    config = pir.IndexPirConfig(entries, entry_size, 2, 1, True, "hybridCompression", False)
    param = pir.MulPir.generateParameter(config, g)
    oparam = opir.generate_parameter(opir.IndexPirConfig(entries, entry_size, 2, 1, True, "hybridCompression", False), n, t)
    database = [bytes(rng.randrange(256) for _ in range(entry_size)) for _ in range(entries)]
    sk, relin = o.keygen(13)
    key, okeys = load_keys(g, o, sk, relin, param.evaluationKeyConfig.galoisElements)
    server = pir.MulPirServer(param, g, [pir.MulPirServer.process(database, g, param)])
    index = rng.randrange(entries)
    # a seeded query: c1 comes from the seed, c0 = Delta m + e - c1 s is built with the oracle's secret key
    query = opir.generate_query(o, oparam, [index], sk, 900)
    seeds = np.random.default_rng(5).integers(0, 256, size=(len(query), 32), dtype=np.uint8)
    seeded = []
    for ct, seed in zip(query, seeds):
        a_eval = drbg.random_poly(n, o.q, seed.tobytes())
        c1 = orc.ntt_inverse(n, o.q, a_eval)
        # keep the encrypted message: c0' = c0 + (c1 - c1') s, computed in Eval format with the secret key
        s_eval = np.asarray(sk, dtype=np.uint64)[: o.L]
        delta = orc.poly_op("sub", n, o.q, orc.ntt_forward(n, o.q, ct[1]), a_eval)
        c0 = orc.poly_op("add", n, o.q, ct[0], orc.ntt_inverse(n, o.q, orc.poly_op("mul", n, o.q, delta, s_eval)))
        seeded.append((np.frombuffer(opir.serialize_poly(n, o.q, c0), dtype=np.uint8), seed, np.stack([c0, c1])))
    poly0 = np.stack([s[0] for s in seeded])
    replies, skips = pir.PirWire.computeResponse(server, poly0, seeds, key)
    assert skips == opir.skip_lsbs_for_decryption(n, o.q[0], t)
    expanded = [drbg.expand_seeded_ciphertext(o, s[0].tobytes(), s[1].tobytes()) for s in seeded]
    for e, s in zip(expanded, seeded):
        assert np.array_equal(e, s[2])
    expected = opir.compute_response(o, expanded, 1, okeys, relin, [opir.process_database(o, oparam, database)], oparam)
    recovered = []
    for chunk in range(server.chunkCount):
        ct = expected[0][chunk]
        want = opir.serialize_poly(n, o.q[:1], ct[0], skips[0]) + opir.serialize_poly(n, o.q[:1], ct[1], skips[1])
        assert replies[0, chunk].tobytes() == want, chunk
        half = len(opir.serialize_poly(n, o.q[:1], ct[0], skips[0]))
        recovered.append(np.stack([opir.load_poly(n, o.q[:1], want[:half], skips[0]), opir.load_poly(n, o.q[:1], want[half:], skips[1])]))
    assert opir.decrypt_response(o, oparam, [recovered], [index], sk) == [database[index]]
    key.close()
    g.close()


def test_captured_response_graph_cache_is_bounded(monkeypatch):
    """More (database, key) pairs than HECUDA_PIR_GRAPH_CACHE: old idle captures are evicted, answers stay right."""
    monkeypatch.setenv("HECUDA_PIR_GRAPH_CACHE", "2")
    g, o = contexts(64, [55, 55, 55], 65537)
    rng = random.Random(11)
    config = pir.IndexPirConfig(100, 8, 2, 1, False, "noCompression", False)
    param = pir.MulPir.generateParameter(config, g)
    oparam = opir.generate_parameter(opir.IndexPirConfig(100, 8, 2, 1, False, "noCompression", False), o.n, o.t)
    database = [bytes(rng.randrange(256) for _ in range(8)) for _ in range(100)]
    server = pir.MulPirServer(param, g, [pir.MulPirServer.process(database, g, param)])
    odb = opir.process_database(o, oparam, database)
    clients = []
    for c in range(4):
        sk, relin = o.keygen(60 + c)
        key, okeys = load_keys(g, o, sk, relin, param.evaluationKeyConfig.galoisElements, seed=300 + 10 * c)
        clients.append((sk, relin, key, okeys))
    for round_ in range(2):
        for c, (sk, relin, key, okeys) in enumerate(clients):
            index = (7 * c + round_) % 100
            query = opir.generate_query(o, oparam, [index], sk, 400 + 10 * c + round_)
            expected = opir.compute_response(o, query, 1, okeys, relin, [odb], oparam)
            got = server.computeResponse(np.stack(query), key, indicesCount=1)
            for chunk in range(server.chunkCount):
                assert np.array_equal(got[0, chunk], expected[0][chunk])
    for _, _, key, _ in clients:
        key.close()
    g.close()


def test_captured_response_graph_follows_key_changes_and_concurrent_callers():
    """The host entry point replays the response pipeline of a (database, key, shape) triple as a CUDA graph
    (csrc/pir.cu).  Replacing a Galois key must invalidate the capture (its device pointers are baked into the graph);
    concurrent callers must each get a private instance; results stay bit-exact with the oracle throughout."""
    import threading

    g, o = contexts(64, [55, 55, 55], 65537)
    rng = random.Random(9)
    config = pir.IndexPirConfig(200, 24, 2, 1, True, "hybridCompression", False)
    param = pir.MulPir.generateParameter(config, g)
    oparam = opir.generate_parameter(opir.IndexPirConfig(200, 24, 2, 1, True, "hybridCompression", False), o.n, o.t)
    database = [bytes(rng.randrange(256) for _ in range(24)) for _ in range(200)]
    server = pir.MulPirServer(param, g, [pir.MulPirServer.process(database, g, param)])
    odb = opir.process_database(o, oparam, database)
    elements = param.evaluationKeyConfig.galoisElements

    def check(sk, relin, key, okeys, index, seed):
        query = opir.generate_query(o, oparam, [index], sk, seed)
        expected = opir.compute_response(o, query, 1, okeys, relin, [odb], oparam)
        got = server.computeResponse(np.stack(query), key, indicesCount=1)
        for chunk in range(server.chunkCount):
            assert np.array_equal(got[0, chunk], expected[0][chunk])
        reply = [[got[0, c] for c in range(server.chunkCount)]]
        assert opir.decrypt_response(o, oparam, reply, [index], sk) == [database[index]]

    sk, relin = o.keygen(41)
    key, okeys = load_keys(g, o, sk, relin, elements, seed=700)
    for trial in range(3):  # first call captures, the next ones replay
        check(sk, relin, key, okeys, rng.randrange(200), 900 + trial)
    # a second client's secret key: new relinearization key object, and the FIRST key's Galois keys replaced in place
    sk2, relin2 = o.keygen(43)
    key2, okeys2 = load_keys(g, o, sk2, relin2, elements, seed=800)
    check(sk2, relin2, key2, okeys2, 17, 950)
    for e in elements:  # replace key material on the original handle: the capture made for it is stale now
        okeys[e] = o.galois_keygen(990 + e, sk, e)
        key.setGaloisKey(e, okeys[e])
    check(sk, relin, key, okeys, 23, 960)
    # concurrent callers on the same (database, key, shape); the oracle side is computed up front (single-threaded)
    jobs = []
    for tid in range(4):
        for k in range(3):
            index = (31 * tid + k) % 200
            query = opir.generate_query(o, oparam, [index], sk2, 1000 + 10 * tid + k)
            jobs.append((tid, np.stack(query), opir.compute_response(o, query, 1, okeys2, relin2, [odb], oparam)))
    errors = []

    def worker(tid):
        try:
            for t, query, expected in jobs:
                if t != tid:
                    continue
                got = server.computeResponse(query, key2, indicesCount=1)
                for chunk in range(server.chunkCount):
                    assert np.array_equal(got[0, chunk], expected[0][chunk])
        except Exception as exc:  # noqa: BLE001
            errors.append(exc)

    threads = [threading.Thread(target=worker, args=(i,)) for i in range(4)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors
    key.close()
    key2.close()
    g.close()
