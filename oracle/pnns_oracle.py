"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the PNNS server's matrix-vector product.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.

Reference code followed (paths relative to Sources/):

  * SIMD encoding            HomomorphicEncryption/Encoding.swift:194-246 (generateEncodingMatrix, encodeSimd, decodeSimd)
  * BabyStepGiantStep        PrivateNearestNeighborSearch/MatrixMultiplication.swift:26-62
  * diagonal / dense-row packing   PrivateNearestNeighborSearch/PlaintextMatrix.swift:246-283,341-482
  * mulTranspose(vector:)    MatrixMultiplication.swift:131-226
  * rotateColumnsAndSum      _HomomorphicEncryptionExtras/HeScheme.swift:113-134 (rotations whose key is present)
  * response post-processing Server.swift:61-88 (modSwitchDownToSingle)

Pinned by the property the reference's MatrixMultiplicationTests check: the decrypted, SIMD-decoded result is the
matrix-vector product modulo t (tests/test_oracle_pnns.py).  `mulTranspose(matrix:)` -- extractDenseRow and the
repacking of several result columns into one ciphertext -- is not restated: one dense-row ciphertext per query vector.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

from . import oracle as O
from .pir_oracle import dividing_ceil, log2, next_power_of_two

GENERATOR = 3  # GaloisElementGenerator.value (PolyRq/Galois.swift:169-171)


def simd_encoding_matrix(n: int) -> list:
    """HeContext.generateEncodingMatrix (Encoding.swift:196-219)."""
    logn, row, mask = log2(n), n >> 1, 2 * n - 1
    matrix, g = [0] * n, 1
    for i in range(row):
        matrix[i] = O.reverse_bits((g - 1) >> 1, logn)
        matrix[row | i] = O.reverse_bits((mask - g) >> 1, logn)
        g = g * GENERATOR & mask
    return matrix


def encode_simd(ctx: O.Context, values) -> np.ndarray:
    """encodeSimd (Encoding.swift:222-235): scatter into Eval positions, inverse NTT modulo t -> Coeff plaintext."""
    matrix = simd_encoding_matrix(ctx.n)
    ev = np.zeros(ctx.n, dtype=np.uint64)
    for i, v in enumerate(values):
        ev[matrix[i]] = int(v) % ctx.t
    return O.ntt_inverse(ctx.n, [ctx.t], ev)[0]


def decode_simd(ctx: O.Context, plain) -> np.ndarray:
    """decodeSimd (Encoding.swift:237-246)."""
    ev = O.ntt_forward(ctx.n, [ctx.t], np.asarray(plain, dtype=np.uint64))[0]
    return ev[np.array(simd_encoding_matrix(ctx.n))]


@dataclass(frozen=True)
class BabyStepGiantStep:
    """BabyStepGiantStep (MatrixMultiplication.swift:26-62)."""

    vector_dimension: int
    baby_step: int
    giant_step: int

    @staticmethod
    def for_dimension(vector_dimension: int) -> "BabyStepGiantStep":
        dimension = next_power_of_two(vector_dimension)
        baby = math.isqrt(dimension)
        if baby * baby < dimension:
            baby += 1  # Int(Double(dimension).squareRoot().rounded(.up))
        giant = dividing_ceil(dimension, baby)
        assert baby >= giant
        return BabyStepGiantStep(dimension, baby, giant)


def evaluation_key_elements(n: int, column_count: int) -> list:
    """The rotations mulTranspose(vector:) needs (MatrixMultiplication.swift:84-94): by -1 and by -babyStep."""
    bsgs = BabyStepGiantStep.for_dimension(column_count)
    return [O.galois_element_rotating_columns(-1, n), O.galois_element_rotating_columns(-bsgs.baby_step, n)]


def diagonal_plaintexts(ctx: O.Context, row_count: int, column_count: int, bsgs: BabyStepGiantStep, values) -> list:
    """PlaintextMatrix.diagonalPlaintexts (PlaintextMatrix.swift:417-482) -> Coeff plaintexts."""
    n = ctx.n
    assert column_count <= n // 2
    data = np.asarray(values, dtype=np.uint64).reshape(row_count, column_count)
    padded_rows = next_power_of_two(column_count)
    packed = np.zeros((padded_rows, row_count), dtype=np.uint64)
    for r in range(padded_rows):
        for c in range(row_count):
            pc = (c + r) % padded_rows
            if pc < column_count:
                packed[r, c] = data[c, pc]
    per_column = dividing_ceil(row_count, n)
    out = []
    for r in range(padded_rows):
        for chunk_index in range(per_column):
            chunk = np.zeros(n, dtype=np.uint64)
            piece = packed[r, chunk_index * n:(chunk_index + 1) * n]
            chunk[: len(piece)] = piece
            i = (len(out) - chunk_index) // per_column
            step = (i // bsgs.baby_step) * bsgs.baby_step
            if step:
                half = n // 2
                chunk = np.concatenate([np.roll(chunk[:half], step), np.roll(chunk[half:], step)])
            out.append(encode_simd(ctx, chunk))
    assert len(out) == padded_rows * per_column
    return out


def dense_row_vector(ctx: O.Context, vector) -> np.ndarray:
    """PlaintextMatrix.denseRowPlaintexts for ONE row (PlaintextMatrix.swift:341-413): the row, padded to a power of
    two, repeated to fill both SIMD rows."""
    n = ctx.n
    v = [int(x) % ctx.t for x in vector]
    packed = v + [0] * (next_power_of_two(len(v)) - len(v))
    simd_columns = n // 2
    if len(packed) < simd_columns and len(packed) + len(v) > simd_columns:
        packed += [0] * (simd_columns - len(packed))
    offset = len(packed) % simd_columns
    packed += [0] * (0 if offset == 0 else next_power_of_two(offset) - offset)
    repeat = list(packed) if len(packed) <= simd_columns else packed[simd_columns:]
    while len(packed) < n:
        packed += repeat
    return encode_simd(ctx, packed)


def _add(ctx, a, b):
    l = a.shape[-2]
    q = np.array(ctx.q[:l], dtype=np.uint64)[None, :, None]
    return (a + b) % q


def rotate_columns(ctx, ct, step: int, galois_keys: dict):
    """Bfv.rotateColumns (HeScheme.swift:1463-1470) = applyGalois with GaloisElement.rotatingColumns(by:degree:)."""
    element = O.galois_element_rotating_columns(step, ctx.n)
    if element not in galois_keys:
        raise KeyError("missingGaloisElement")
    return ctx.apply_galois(ct, element, galois_keys[element], threads=1)[0]


def rotate_columns_and_sum(ctx, cts: list, step: int, galois_keys: dict):
    """HeScheme.rotateColumnsAndSum (_HomomorphicEncryptionExtras/HeScheme.swift:113-134)."""
    cts = list(cts)
    acc = cts.pop()
    for ct in reversed(cts):
        acc = rotate_columns(ctx, acc, step, galois_keys)
        acc = _add(ctx, acc, ct)
    return acc


def mul_transpose_vector(ctx: O.Context, plaintexts: list, row_count: int, bsgs: BabyStepGiantStep, ct, galois_keys: dict):
    """PlaintextMatrix.mulTranspose(vector:using:) (MatrixMultiplication.swift:131-226).  plaintexts: the diagonal
    packing in Coeff format; ct: dense-row ciphertext (2, L, N) Coeff.  Returns resultCiphertextCount ciphertexts."""
    n, L = ctx.n, ct.shape[-2]
    states, state = [], ct
    for step in range(bsgs.baby_step):
        states.append(state)
        if step != bsgs.baby_step - 1:
            state = rotate_columns(ctx, state, -1, galois_keys)
    rotated = np.stack([np.stack([O.ntt_forward(n, ctx.q[:L], s[p]) for p in range(2)]) for s in states])
    result_count = dividing_ceil(row_count, n)
    out = []
    for r in range(result_count):
        to_add = []
        for g in range(bsgs.giant_step):
            count = min(len(states), bsgs.vector_dimension - bsgs.baby_step * g)
            indices = [result_count * (j + bsgs.baby_step * g) + r for j in range(count)]
            rows = np.stack([ctx.plaintext_to_eval(plaintexts[i], L) for i in indices])
            ip = ctx.inner_product_plain(rotated[:count], rows[None], None, threads=1)[0]
            to_add.append(np.stack([O.ntt_inverse(n, ctx.q[:L], ip[p]) for p in range(2)]))
        out.append(rotate_columns_and_sum(ctx, to_add, -bsgs.baby_step, galois_keys))
    return out


def mod_switch_down_to_single(ctx, ct):
    while ct.shape[-2] > 1:
        ct = ctx.mod_switch_down(ct, threads=1)[0]
    return ct
