"""Pins of the MulPir restatement (oracle/pir_oracle.py) on the reference's own vectors and properties.

KATs: MulPirTests.evaluationKeyConfig / computeCoordinates / unevenDimensionVectorsTest
(Sources/_TestUtilities/PirUtilities/MulPirTests.swift:39-237), CoefficientPackingTests bytesToCoeffKAT / coeffsToBytesKAT
(Tests/HomomorphicEncryptionTests/CoefficientPackingTests.swift:83-211).  Properties: ExpansionTests
(Sources/_TestUtilities/PirUtilities/ExpansionTests.swift:25-160) and the end-to-end IndexPirTests.indexPirTest
(IndexPirTests.swift:23-128): the decrypted response is the queried database entry.
"""
import random

import numpy as np
import pytest

from oracle import oracle as orc
from oracle import pir_oracle as pir


def reference_test_context():
    """TestUtils.getTestContext for UInt64 (TestUtilities.swift:297-335): N=16, t=1153, moduli 55/52/62/58 bits."""
    n = 16
    moduli = orc.generate_primes([55, 52, 62, 58], False, n)
    return orc.Context(n, moduli, 1153)


def evaluation_key(ctx, sk, elements, seed=100):
    return {e: ctx.galois_keygen(seed + i, sk, e) for i, e in enumerate(elements)}


# ------------------------------------------------------------------------------------------------------ KATs
EVK_KATS = [
    ("noCompression", 2, 4096, [4097]), ("noCompression", 2, 8192, [8193]),
    ("noCompression", 32, 4096, [257, 513, 1025, 2049, 4097]), ("noCompression", 32, 8192, [513, 1025, 2049, 4097, 8193]),
    ("noCompression", 1024, 4096, [9, 17, 33, 65, 129, 257, 513, 1025, 2049, 4097]),
    ("noCompression", 1024, 8192, [17, 33, 65, 129, 257, 513, 1025, 2049, 4097, 8193]),
    ("hybridCompression", 2, 4096, [4097]), ("hybridCompression", 2, 8192, [8193]),
    ("hybridCompression", 32, 4096, [257, 1025]), ("hybridCompression", 32, 8192, [513, 2049]),
    ("hybridCompression", 1024, 4096, [9, 17, 33, 65, 129, 1025]), ("hybridCompression", 1024, 8192, [17, 33, 65, 129, 1025]),
    ("maxCompression", 2, 4096, [4097]), ("maxCompression", 2, 8192, [8193]),
    ("maxCompression", 32, 4096, [257]), ("maxCompression", 32, 8192, [513]),
    ("maxCompression", 1024, 4096, [9, 17, 33, 65, 129]), ("maxCompression", 1024, 8192, [17, 33, 65, 129]),
]


@pytest.mark.parametrize("compression,count,degree,expected", EVK_KATS)
def test_evaluation_key_config_kats(compression, count, degree, expected):
    assert pir.evaluation_key_config(count, degree, compression) == expected


def test_compute_coordinates_kats():
    p2 = pir.IndexPirParameter(100, 16, [10, 10])
    for index, expected in [(0, [0, 0]), (1, [0, 1]), (2, [0, 2]), (10, [1, 0]), (11, [1, 1]), (12, [1, 2]), (98, [9, 8]),
                            (99, [9, 9])]:
        assert pir.compute_coordinates(p2, index, 1) == expected
    p3 = pir.IndexPirParameter(30, 16, [5, 3, 2])
    for index, expected in [(0, [0, 0, 0]), (1, [0, 0, 1]), (2, [0, 1, 0]), (10, [1, 2, 0]), (11, [1, 2, 1]),
                            (12, [2, 0, 0]), (27, [4, 1, 1]), (28, [4, 2, 0]), (29, [4, 2, 1])]:
        assert pir.compute_coordinates(p3, index, 1) == expected


@pytest.mark.parametrize("entries,batch,even,uneven", [(9, 1, [3, 3], [5, 2]), (20, 1, [5, 4], [10, 2]),
                                                       (100, 1, [10, 10], [25, 4]), (72, 1, [9, 8], [24, 3]),
                                                       (100, 3, [10, 10], [13, 8])])
def test_uneven_dimension_vectors(entries, batch, even, uneven):
    # entrySizeInBytes = 21 > bytesPerPlaintext = 20 of the test context, so perChunkPlaintextCount == entryCount
    assert pir.bytes_per_plaintext(16, 1153) == 20
    for flag, expected in ((False, even), (True, uneven)):
        config = pir.IndexPirConfig(entries, 21, 2, batch, flag, "noCompression", False)
        assert pir.generate_parameter(config, 16, 1153).dimensions == expected


BASE = [0, 3, 1, 8, 5, 15, 8, 13, 11, 3, 2, 2, 7, 1]
BYTES_TO_COEFF = [
    ([3, 24, 95, 141, 179, 34, 113], 4, 0, False, BASE),
    ([3, 24, 95, 141, 179, 34, 113], 4, 0, True, BASE),
    ([4, 69, 230, 164, 150, 0], 4, 1, True, [0, 2, 0, 8, 4, 14, 8, 12, 10, 2, 2, 2, 6, 0, 0, 0]),
    ([2, 123, 128, 64], 4, 2, False, [0, 0, 0, 8, 4, 12, 8, 12, 8, 0, 0, 0, 4, 0, 0, 0]),
    ([2, 123, 128, 64], 4, 2, True, [0, 0, 0, 8, 4, 12, 8, 12, 8, 0, 0, 0, 4, 0, 0, 0]),
    ([23, 128], 4, 3, True, [0, 0, 0, 8, 0, 8, 8, 8, 8, 0, 0, 0, 0, 0, 0, 0]),
    (list(range(256)), 8, 0, False, list(range(256))),
]
COEFF_TO_BYTES = [
    (BASE, 4, 0, [3, 24, 95, 141, 179, 34, 113]), (BASE, 4, 1, [4, 69, 230, 164, 150, 0]), (BASE, 4, 2, [2, 123, 128, 64]),
    (BASE, 4, 3, [23, 128]), (BASE, 5, 0, [0, 194, 130, 189, 13, 88, 196, 35, 132]),
    ([19, 16, 21, 4, 0, 1, 15, 3, 10, 3], 5, 1, [152, 162, 0, 113, 81]),
    ([19, 16, 21, 4, 0, 1, 15, 3, 10, 3], 5, 2, [146, 144, 24, 64]), (list(range(256)), 8, 0, list(range(256))),
]


@pytest.mark.parametrize("data,bits,skip,decode,expected", BYTES_TO_COEFF)
def test_bytes_to_coefficients_kats(data, bits, skip, decode, expected):
    assert pir.bytes_to_coefficients(bytes(data), bits, decode, skip).tolist() == expected


@pytest.mark.parametrize("coeffs,bits,skip,expected", COEFF_TO_BYTES)
def test_coefficients_to_bytes_kats(coeffs, bits, skip, expected):
    assert list(pir.coefficients_to_bytes(coeffs, bits, skip)) == expected


def test_packing_roundtrip():
    rng = random.Random(5)
    for bits in (1, 4, 10, 16, 19, 31, 47, 63):
        data = bytes(rng.randrange(256) for _ in range(512))
        coeffs = pir.bytes_to_coefficients(data, bits, decode=False)
        assert int(coeffs.max()) < (1 << bits)
        assert pir.coefficients_to_bytes(coeffs, bits)[: len(data)] == data


# ------------------------------------------------------------------------------------------------------ expansion
@pytest.mark.parametrize("compression", ["noCompression", "hybridCompression", "maxCompression"])
def test_expand_ciphertext_for_one_step(compression):
    """ExpansionTests.expandCiphertextForOneStep: N=32, 4 primes, t=17; p0[k*step] = 2 c[k*step], p1[k*step] = 2 c[k*step+step/2]."""
    n = 32
    moduli = orc.generate_primes([60, 60, 60, 60], False, n)
    ctx = orc.Context(n, moduli, 17)
    rng = random.Random(3)
    sk, _ = ctx.keygen(11, relin=False)
    keys = evaluation_key(ctx, sk, pir.evaluation_key_config(n, n, compression))
    for log_step in range(1, 6):
        step, half = 1 << log_step, 1 << (log_step - 1)
        data = np.array([rng.randrange(17) for _ in range(n)], dtype=np.uint64)
        ct = ctx.encrypt(50 + log_step, sk, data)
        p0, p1 = pir.expand_ciphertext_for_one_step(ctx, ct, log_step, keys)
        d0, d1 = ctx.decrypt(sk, p0), ctx.decrypt(sk, p1)
        for index in range(0, n, step):
            assert int(d0[index]) == int(data[index]) * 2 % 17
            assert int(d1[index]) == int(data[index + half]) * 2 % 17


def test_one_ciphertext_roundtrip():
    """ExpansionTests.oneCiphertextRoundtrip: every inputCount in 1...degree."""
    ctx = reference_test_context()
    n, rng = ctx.n, random.Random(8)
    sk, _ = ctx.keygen(21, relin=False)
    keys = evaluation_key(ctx, sk, [(1 << k) + 1 for k in range(1, pir.log2(n) + 1)])
    for count in range(1, n + 1):
        data = [rng.randrange(2) for _ in range(count)]
        ones = [i for i, v in enumerate(data) if v]
        plain = pir.compress_inputs_for_one_ciphertext(ctx, count, ones)
        ct = ctx.encrypt(1000 + count, sk, plain)
        out = pir.expand_ciphertext(ctx, ct, count, 1, pir.ceil_log2(count), keys)
        assert len(out) == count
        for index in range(count):
            dec = ctx.decrypt(sk, out[index])
            assert int(dec[0]) == data[index] and not dec[1:].any()


def test_multiple_ciphertexts_roundtrip():
    ctx = reference_test_context()
    n, rng = ctx.n, random.Random(9)
    sk, _ = ctx.keygen(22, relin=False)
    keys = evaluation_key(ctx, sk, [(1 << k) + 1 for k in range(1, pir.log2(n) + 1)])
    for count in (1, 5, 16, 17, 23, 32):
        data = [rng.randrange(2) for _ in range(count)]
        ones = [i for i, v in enumerate(data) if v]
        cts = pir.compress_binary_inputs(ctx, count, ones, sk, 2000 + count)
        out = pir.expand(ctx, cts, count, keys)
        assert len(out) == count
        for index in range(count):
            dec = ctx.decrypt(sk, out[index])
            assert int(dec[0]) == data[index] and not dec[1:].any()


# ------------------------------------------------------------------------------------------------------ end to end
CONFIGS = [
    dict(entry_size_in_bytes=1, dimension_count=2, uneven_dimensions=False, key_compression="noCompression"),
    dict(entry_size_in_bytes=8, dimension_count=2, uneven_dimensions=False, key_compression="noCompression"),
    dict(entry_size_in_bytes=24, dimension_count=2, uneven_dimensions=True, key_compression="noCompression"),
    dict(entry_size_in_bytes=24, dimension_count=1, uneven_dimensions=True, key_compression="noCompression"),
    dict(entry_size_in_bytes=24, dimension_count=1, uneven_dimensions=True, key_compression="hybridCompression"),
    dict(entry_size_in_bytes=24, dimension_count=1, uneven_dimensions=True, key_compression="maxCompression"),
]


@pytest.mark.parametrize("encoding_entry_size", [False, True])
@pytest.mark.parametrize("cfg", CONFIGS)
def test_index_pir_end_to_end(cfg, encoding_entry_size):
    """IndexPirTests.indexPirTest: 100 entries, batch 2, test context; the decrypted reply is the database entry."""
    ctx = reference_test_context()
    rng = random.Random(hash((cfg["entry_size_in_bytes"], cfg["dimension_count"], encoding_entry_size)) & 0xFFFF)
    config = pir.IndexPirConfig(entry_count=100, batch_size=2, encoding_entry_size=encoding_entry_size, **cfg)
    param = pir.generate_parameter(config, ctx.n, ctx.t)
    size = param.entry_size_in_bytes
    database = [bytes(rng.randrange(256) for _ in range(rng.randint(1, size) if encoding_entry_size else size))
                for _ in range(param.entry_count)]
    db = pir.process_database(ctx, param, database)
    chunk_count = pir.dividing_ceil(param.encoded_entry_size, pir.bytes_per_plaintext(ctx.n, ctx.t))
    assert db.plaintexts.shape[0] == chunk_count * param.per_chunk_plaintext_count
    sk, relin = ctx.keygen(31)
    keys = evaluation_key(ctx, sk, param.galois_elements)
    for trial, batch in enumerate((2, 1)):
        indices = rng.sample(range(param.entry_count), batch)
        query = pir.generate_query(ctx, param, indices, sk, 400 + 10 * trial)
        response = pir.compute_response(ctx, query, len(indices), keys, relin, [db], param)
        assert len(response) == batch and all(len(r) == chunk_count for r in response)
        assert all(ct.shape == (2, 1, ctx.n) for r in response for ct in r)
        entries = pir.decrypt_response(ctx, param, response, indices, sk)
        assert entries == [database[i] for i in indices]


# ------------------------------------------------------------------------------------------------------ wire format
@pytest.mark.parametrize("poly,moduli,skip,expected", [
    ([1, 2, 3, 4, 6, 7, 8, 9], [521, 541], 0, [1, 2, 3, 4, 6, 7, 8, 9]),
    ([1, 2, 3, 4], [521], 1, [0, 2, 2, 4]),
    ([1, 21, 302, 417], [521], 2, [0, 20, 300, 416]),
])
def test_poly_serialize_roundtrip_kats(poly, moduli, skip, expected):
    """PolyRq+SerializeTests.roundtripKAT (PolyRq+SerializeTests.swift:61-104)."""
    n = len(poly) // len(moduli)
    data = pir.serialize_poly(n, moduli, poly, skip)
    assert len(data) == pir.serialization_byte_count(n, moduli, skip)
    assert pir.load_poly(n, moduli, data, skip).reshape(-1).tolist() == expected
    with pytest.raises(ValueError):   # serializationErrOnWrongBuffer (:21-37)
        pir.load_poly(n, moduli, data + b"\0", skip)


@pytest.mark.parametrize("n,q0,t,expected", [
    (4096, (1 << 27) - 40959, 17, [19, 11]),                    # n_4096_logq_27_28_28_logt_5
    (8192, (1 << 55) - 311295, (1 << 23) + 16385, [28, 19]),    # n_8192_logq_3x55_logt_24
])
def test_skip_lsbs_for_decryption_kats(n, q0, t, expected):
    """EncryptionParametersTests.predefined KAT column skipLSBs (EncryptionParametersTests.swift:218-252) for the two
    parameter sets whose moduli are written out in EncryptionParameters.swift:357-367,401-411."""
    assert pir.skip_lsbs_for_decryption(n, q0, t) == expected
