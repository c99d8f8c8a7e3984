"""GPU parity for the PNNS matrix-vector product: PlaintextMatrix.mulTranspose(vector:using:)
through the C ABI, bit-exact against oracle/pnns_oracle.py, plus the reference's property decrypt(M x v^T) == M v mod t."""
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import hecuda
from hecuda import pnns
from oracle import oracle as orc
from oracle import pnns_oracle as opn


def setup(n, t, bits, rows, cols, seed):
    moduli = orc.generate_primes(list(bits), False, n)
    g, o = hecuda.Context(n, moduli, t), orc.Context(n, moduli, t)
    rng = random.Random(seed)
    matrix = [[rng.randrange(t) for _ in range(cols)] for _ in range(rows)]
    sk, _ = o.keygen(9, relin=False)
    elements = [pnns.GaloisElement.rotatingColumns(-1, n)]
    bsgs = pnns.BabyStepGiantStep.forVectorDimension(cols)
    if bsgs.giantStep > 1:
        elements.append(pnns.GaloisElement.rotatingColumns(-bsgs.babyStep, n))
    key, okeys = hecuda.EvaluationKey(g, None), {}
    for i, e in enumerate(dict.fromkeys(elements)):
        okeys[e] = o.galois_keygen(70 + i, sk, e)
        key.setGaloisKey(e, okeys[e])
    return g, o, rng, matrix, sk, key, okeys, bsgs


@pytest.mark.parametrize("n,t,bits,rows,cols", [
    (16, 1153, (55, 52, 62, 58), 10, 4), (16, 1153, (55, 52, 62, 58), 16, 8), (16, 1153, (55, 52, 62, 58), 40, 5),
    (64, 65537, (55, 55, 55), 100, 24), (64, 65537, (55, 55, 55), 64, 32), (16, 1153, (55, 52, 62, 58), 7, 1),
    (4096, 65537, (36, 36, 37), 5000, 128), (8192, 65537, (55, 55, 55, 55), 9000, 512)])
def test_mul_transpose_vector_matches_oracle_and_decrypts(n, t, bits, rows, cols):
    g, o, rng, matrix, sk, key, okeys, bsgs = setup(n, t, bits, rows, cols, rows * 31 + cols)
    flat = [v for row in matrix for v in row]
    device_matrix = pnns.PlaintextMatrix(g, pnns.MatrixDimensions(rows, cols), flat)
    oplain = opn.diagonal_plaintexts(o, rows, cols, opn.BabyStepGiantStep.for_dimension(cols), flat)
    batch = 3 if n <= 64 else 1
    vectors = [[rng.randrange(t) for _ in range(cols)] for _ in range(batch)]
    cts = np.stack([o.encrypt(12 + b, sk, pnns.denseRowVector(g, v)) for b, v in enumerate(vectors)])
    got = device_matrix.mulTranspose(cts, key)
    single = device_matrix.mulTranspose(cts, key, modSwitchDownToSingle=True)
    assert got.shape == (batch, device_matrix.resultCiphertextCount, 2, o.L, n)
    for b in range(batch):
        expected = opn.mul_transpose_vector(o, oplain, rows, opn.BabyStepGiantStep.for_dimension(cols), cts[b], okeys)
        decoded = []
        for r, ct in enumerate(expected):
            assert np.array_equal(got[b, r], ct), (b, r)
            assert np.array_equal(single[b, r], opn.mod_switch_down_to_single(o, ct)), (b, r)
            decoded += opn.decode_simd(o, o.decrypt(sk, single[b, r])).tolist()
        assert decoded[:rows] == [sum(x * y for x, y in zip(row, vectors[b])) % t for row in matrix]
    device_matrix.close()
    key.close()
    g.close()


def test_pnns_errors():
    g, o, rng, matrix, sk, key, okeys, bsgs = setup(16, 1153, (55, 52, 62, 58), 10, 4, 1)
    flat = [v for row in matrix for v in row]
    with pytest.raises(pnns.PnnsError):                       # columnCount > simd column count
        pnns.PlaintextMatrix(g, pnns.MatrixDimensions(4, 9), list(range(36)))
    with pytest.raises(pnns.PnnsError):                       # babyStep < giantStep (MatrixMultiplication.swift:39-40)
        pnns.BabyStepGiantStep(4, 1, 4)
    device_matrix = pnns.PlaintextMatrix(g, pnns.MatrixDimensions(10, 4), flat)
    ct = o.encrypt(3, sk, pnns.denseRowVector(g, [1, 2, 3, 4]))
    bare = hecuda.EvaluationKey(g, None)
    with pytest.raises(hecuda.HeError) as err:                # missing Galois key for rotateColumns(by: -1)
        device_matrix.mulTranspose(ct, bare)
    assert "missingGaloisElement" in str(err.value)
    other = hecuda.Context(16, g.coefficientModuli, 1153)
    with pytest.raises(hecuda.HeError):                       # PnnsError.wrongContext
        hecuda.load_library()  # keep the loader referenced
        pnns.PlaintextMatrix.mulTranspose(device_matrix, ct, hecuda.EvaluationKey(other, None))
    bare.close()
    device_matrix.close()
    key.close()
    g.close()


@pytest.mark.parametrize("n,t,bits,rows,cols,queries", [
    (16, 1153, (55, 52, 62, 58), 4, 2, 3), (16, 1153, (55, 52, 62, 58), 3, 4, 5), (16, 1153, (55, 52, 62, 58), 8, 4, 2),
    (16, 1153, (55, 52, 62, 58), 20, 3, 3), (64, 65537, (55, 55, 55), 10, 8, 9), (64, 65537, (55, 55, 55), 40, 12, 4),
    (64, 65537, (55, 55, 55), 5, 32, 3), (64, 65537, (55, 55, 55), 10, 8, 1),
    (4096, 65537, (36, 36, 37), 300, 128, 20), (8192, 65537, (55, 55, 55, 55), 5000, 384, 3)])
def test_mul_transpose_matrix_matches_oracle_and_decrypts(n, t, bits, rows, cols, queries):
    """PlaintextMatrix.mulTranspose(matrix:using:) incl. extractDenseRow and dense-column packing."""
    moduli = orc.generate_primes(list(bits), False, n)
    g, o = hecuda.Context(n, moduli, t), orc.Context(n, moduli, t)
    rng = random.Random(rows * 131 + cols * 7 + queries)
    matrix = [[rng.randrange(t) for _ in range(cols)] for _ in range(rows)]
    query = [[rng.randrange(t) for _ in range(cols)] for _ in range(queries)]
    sk, _ = o.keygen(9, relin=False)
    key, okeys = hecuda.EvaluationKey(g, None), {}
    for i, e in enumerate(opn.matrix_evaluation_key_elements(n, rows, cols, queries)):
        okeys[e] = o.galois_keygen(70 + i, sk, e)
        key.setGaloisKey(e, okeys[e])
    flat = [v for row in matrix for v in row]
    device_matrix = pnns.PlaintextMatrix(g, pnns.MatrixDimensions(rows, cols), flat)
    bsgs = opn.BabyStepGiantStep.for_dimension(cols)
    oplain = opn.diagonal_plaintexts(o, rows, cols, bsgs, flat)
    query_plain = opn.dense_row_plaintexts(o, queries, cols, [v for row in query for v in row])
    cts = [o.encrypt(200 + i, sk, p) for i, p in enumerate(query_plain)]
    got = device_matrix.mulTransposeMatrix(np.stack(cts), pnns.MatrixDimensions(queries, cols), key)
    expected = opn.mul_transpose_matrix(o, oplain, rows, cols, bsgs, cts, queries, okeys)
    assert got.shape[0] == len(expected)
    for i, ct in enumerate(expected):
        assert np.array_equal(got[i], ct), i
    single = device_matrix.mulTransposeMatrix(np.stack(cts), pnns.MatrixDimensions(queries, cols), key, modSwitchDownToSingle=True)
    decoded = []
    for i, ct in enumerate(expected):
        assert np.array_equal(single[i], opn.mod_switch_down_to_single(o, ct)), i
        decoded.append(opn.decode_simd(o, o.decrypt(sk, single[i])).tolist())
    assert opn.unpack_dense_column(o, decoded, rows, queries) == \
        [sum(a * b for a, b in zip(matrix[r], query[c])) % t for r in range(rows) for c in range(queries)]
    with pytest.raises(pnns.PnnsError):   # PnnsError.invalidMatrixDimensions: column counts differ (:242-244)
        device_matrix.mulTransposeMatrix(np.stack(cts), pnns.MatrixDimensions(queries, cols + 1), key)
    device_matrix.close()
    key.close()
    g.close()
