"""Host logic of hecuda.pnns (no GPU) against the independent restatement in oracle/pnns_oracle.py."""
import random
from types import SimpleNamespace

import numpy as np
import pytest

from hecuda import pnns
from oracle import oracle as orc
from oracle import pnns_oracle as opn


@pytest.mark.parametrize("n,t", [(16, 1153), (64, 65537), (1024, 65537)])
def test_simd_encoder_matches_oracle(n, t):
    ctx = orc.Context(n, orc.generate_primes([55, 55], False, n), t)
    enc = pnns.SimdEncoder(n, t)
    assert enc.psi == orc.min_primitive_root(2 * n, t)
    assert enc.encodingMatrix.tolist() == opn.simd_encoding_matrix(n)
    rng = random.Random(n)
    values = np.array([[rng.randrange(t) for _ in range(n)] for _ in range(3)], dtype=np.uint64)
    plain = enc.encode(values)
    for i in range(3):
        assert np.array_equal(plain[i], opn.encode_simd(ctx, values[i]))
    assert np.array_equal(enc.decode(plain), values)


def test_galois_elements_match_oracle():
    for n in (16, 4096, 8192):
        for step in (-1, 1, 3, -5, n // 2 - 1):
            assert pnns.GaloisElement.rotatingColumns(step, n) == orc.galois_element_rotating_columns(step, n)
        assert pnns.GaloisElement.swappingRows(n) == orc.galois_element_swapping_rows(n)


@pytest.mark.parametrize("dim", [1, 2, 3, 4, 5, 8, 100, 128, 384, 512, 4096])
def test_baby_step_giant_step_matches_oracle(dim):
    a, b = pnns.BabyStepGiantStep.forVectorDimension(dim), opn.BabyStepGiantStep.for_dimension(dim)
    assert (a.vectorDimension, a.babyStep, a.giantStep) == (b.vector_dimension, b.baby_step, b.giant_step)


@pytest.mark.parametrize("n,t,rows,cols", [(16, 1153, 10, 4), (16, 1153, 40, 5), (64, 65537, 100, 24), (64, 65537, 64, 32)])
def test_diagonal_packing_and_dense_row_match_oracle(n, t, rows, cols):
    octx = orc.Context(n, orc.generate_primes([55, 55], False, n), t)
    host = SimpleNamespace(degree=n, plaintextModulus=t)
    rng = random.Random(rows + cols)
    values = [rng.randrange(t) for _ in range(rows * cols)]
    bsgs = pnns.BabyStepGiantStep.forVectorDimension(cols)
    got = pnns.PlaintextMatrix.diagonalPlaintexts(host, pnns.MatrixDimensions(rows, cols), bsgs, values)
    expected = opn.diagonal_plaintexts(octx, rows, cols, opn.BabyStepGiantStep.for_dimension(cols), values)
    assert got.shape[0] == len(expected)
    for i, row in enumerate(expected):
        assert np.array_equal(got[i], row), i
    vector = values[:cols]
    assert np.array_equal(pnns.denseRowVector(host, vector), opn.dense_row_vector(octx, vector))


def test_dense_row_extraction_and_rotation_plans_match_oracle():
    rng = random.Random(5)
    for _ in range(400):
        n = rng.choice([16, 64, 4096])
        cols = rng.randint(1, n // 2)
        rows = rng.randint(2, 3 * n)
        dims = pnns.MatrixDimensions(rows, cols)
        count = pnns.CiphertextMatrix.ciphertextCount(n, dims)
        for r in {0, rows - 1, rng.randrange(rows), rng.randrange(rows)}:
            index, mask, rotations = pnns.CiphertextMatrix.denseRowExtraction(n, dims, count, r)
            expected = opn.dense_row_extraction(n, rows, cols, count, r)
            assert (index, mask, rotations) == (expected.ciphertext_index, expected.mask, expected.rotate_count)
    for n in (16, 64, 8192):
        elements = [orc.galois_element_rotating_columns(s, n) for s in (1, -1, 4)] + [2 * n - 1]
        assert pnns.GaloisElement.stepsFor(elements, n) == opn.steps_for(elements, n)
        for step in range(1, min(n // 2, 40)):
            assert pnns.GaloisElement.rotationSequence(elements, step, n) == opn.rotation_sequence(elements, step, n)
    assert pnns.GaloisElement.planMultiStep([1, 16, 256], 33, 8192) == {16: 2, 1: 1}
    assert pnns.GaloisElement.planMultiStep([2, 4], 3, 64) is None
