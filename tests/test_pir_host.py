"""Host logic of hecuda.pir (no GPU): the reference's KATs, and agreement with the independent restatement in
oracle/pir_oracle.py on parameters and on the plaintext rows `MulPirServer.process` uploads."""
import random
from types import SimpleNamespace

import numpy as np
import pytest

from hecuda import pir
from oracle import oracle as orc
from oracle import pir_oracle as opir
from test_oracle_pir import BYTES_TO_COEFF, COEFF_TO_BYTES, EVK_KATS


@pytest.mark.parametrize("compression,count,degree,expected", EVK_KATS)
def test_evaluation_key_config_kats(compression, count, degree, expected):
    config = pir.MulPir.evaluationKeyConfig(count, degree, compression)
    assert config.galoisElements == expected and config.hasRelinearizationKey


@pytest.mark.parametrize("data,bits,skip,decode,expected", BYTES_TO_COEFF)
def test_bytes_to_coefficients_kats(data, bits, skip, decode, expected):
    assert pir.CoefficientPacking.bytesToCoefficients(bytes(data), bits, decode, skip).tolist() == expected


@pytest.mark.parametrize("coeffs,bits,skip,expected", COEFF_TO_BYTES)
def test_coefficients_to_bytes_kats(coeffs, bits, skip, expected):
    assert list(pir.CoefficientPacking.coefficientsToBytes(coeffs, bits, skip)) == expected


def test_invalid_packing_and_dimension_count():
    with pytest.raises(Exception):
        pir.CoefficientPacking.bytesToCoefficients(b"ab", 4, False, 4)
    with pytest.raises(pir.PirError):
        pir.IndexPirConfig(10, 1, 3, 1, False, "noCompression")


@pytest.mark.parametrize("entries,batch,even,uneven", [(9, 1, [3, 3], [5, 2]), (20, 1, [5, 4], [10, 2]),
                                                       (100, 1, [10, 10], [25, 4]), (72, 1, [9, 8], [24, 3]),
                                                       (100, 3, [10, 10], [13, 8])])
def test_uneven_dimension_vectors(entries, batch, even, uneven):
    ctx = SimpleNamespace(degree=16, plaintextModulus=1153)
    for flag, expected in ((False, even), (True, uneven)):
        config = pir.IndexPirConfig(entries, 21, 2, batch, flag, "noCompression", False)
        assert pir.MulPir.generateParameter(config, ctx).dimensions == expected


def test_parameters_agree_with_oracle():
    rng = random.Random(1)
    for _ in range(300):
        degree = rng.choice([16, 4096, 8192])
        t = rng.choice([17, 1153, 65537, 557057])
        kw = dict(entry_count=rng.randint(1, 200000), entry_size_in_bytes=rng.choice([1, 8, 24, 100, 3000, 20000]),
                  dimension_count=rng.choice([1, 2]), batch_size=rng.choice([1, 2, 3, 8]),
                  uneven_dimensions=rng.random() < 0.5,
                  key_compression=rng.choice(["noCompression", "hybridCompression", "maxCompression"]),
                  encoding_entry_size=rng.random() < 0.5)
        expected = opir.generate_parameter(opir.IndexPirConfig(**kw), degree, t)
        got = pir.MulPir.generateParameter(
            pir.IndexPirConfig(kw["entry_count"], kw["entry_size_in_bytes"], kw["dimension_count"], kw["batch_size"],
                               kw["uneven_dimensions"], kw["key_compression"], kw["encoding_entry_size"]),
            SimpleNamespace(degree=degree, plaintextModulus=t))
        assert got.dimensions == expected.dimensions
        assert got.evaluationKeyConfig.galoisElements == expected.galois_elements
        assert got.encodedEntrySize == expected.encoded_entry_size and got.expandedQueryCount == expected.expanded_query_count


@pytest.mark.parametrize("entry_size,dimension_count,encoding", [(1, 2, False), (8, 2, True), (24, 2, False), (24, 1, True),
                                                                  (47, 2, True), (60, 1, False)])
def test_plaintext_rows_agree_with_oracle(entry_size, dimension_count, encoding):
    n, t = 16, 1153
    moduli = orc.generate_primes([55, 52, 62, 58], False, n)
    octx = orc.Context(n, moduli, t)
    rng = random.Random(entry_size * 7 + dimension_count)
    entries = [bytes(rng.randrange(256) for _ in range(rng.randint(0, entry_size) if encoding else entry_size))
               for _ in range(57)]
    entries[3] = bytes(len(entries[3]))  # an all-zero entry (nil plaintext when it fills a plaintext alone)
    config = pir.IndexPirConfig(len(entries), entry_size, dimension_count, 2, True, "noCompression", encoding)
    param = pir.MulPir.generateParameter(config, SimpleNamespace(degree=n, plaintextModulus=t))
    rows, present = pir.MulPirServer.plaintextRows(entries, SimpleNamespace(degree=n, plaintextModulus=t), param)
    oparam = opir.generate_parameter(opir.IndexPirConfig(len(entries), entry_size, dimension_count, 2, True,
                                                         "noCompression", encoding), n, t)
    odb = opir.process_database(octx, oparam, entries)
    assert present.tolist() == odb.present.tolist()
    assert rows.shape[0] == odb.plaintexts.shape[0]
    for i in range(rows.shape[0]):
        if present[i]:
            assert np.array_equal(octx.plaintext_to_eval(rows[i]), odb.plaintexts[i])
        else:
            assert not rows[i].any()


def test_skip_lsbs_for_decryption_matches_oracle_and_kats():
    cases = [(4096, (1 << 27) - 40959, 17, [19, 11]), (8192, (1 << 55) - 311295, (1 << 23) + 16385, [28, 19])]
    for n, q0, t, expected in cases:
        assert pir.skipLSBsForDecryption(SimpleNamespace(degree=n, plaintextModulus=t, coefficientModuli=[q0])) == expected
    rng = random.Random(2)
    for _ in range(300):
        n = 1 << rng.randint(3, 15)
        q0 = rng.randrange(1 << 10, 1 << 62)
        t = rng.randrange(2, q0)
        got = pir.skipLSBsForDecryption(SimpleNamespace(degree=n, plaintextModulus=t, coefficientModuli=[q0]))
        assert got == opir.skip_lsbs_for_decryption(n, q0, t)
