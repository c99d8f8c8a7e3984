"""hecuda.pnns -- host-side mirror of the PNNS server's matrix-vector product over libhecuda.

    BabyStepGiantStep, MatrixDimensions        PrivateNearestNeighborSearch/MatrixMultiplication.swift:26-62,
                                               PlaintextMatrix.swift:45-72
    PlaintextMatrix (.diagonal packing)        PlaintextMatrix.swift:417-482
    PlaintextMatrix.mulTranspose(vector:using:) MatrixMultiplication.swift:131-226
    SIMD encoding                              HomomorphicEncryption/Encoding.swift:194-246

The diagonal packing and the SIMD encoding of the database are host work in the reference too (offline
preprocessing); conversion to Eval format, the rotations, inner products and modulus switching run on the device.
`mulTranspose(matrix:)` (extractDenseRow + repacking of several result columns per ciphertext) is not mirrored yet:
pass one dense-row ciphertext per query vector.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass

import numpy as np

from . import Context, EvaluationKey, HeError, _check, _host, _ptr, load_library


class PnnsError(ValueError):
    pass


def _next_power_of_two(v: int) -> int:
    return 1 << max(0, (int(v) - 1).bit_length())


@dataclass(frozen=True)
class BabyStepGiantStep:
    """BabyStepGiantStep (MatrixMultiplication.swift:26-62)."""

    vectorDimension: int
    babyStep: int
    giantStep: int

    def __post_init__(self):
        if self.babyStep < self.giantStep:
            raise PnnsError("babyStep cannot be smaller than giantStep")

    @staticmethod
    def forVectorDimension(vectorDimension: int) -> "BabyStepGiantStep":
        dimension = _next_power_of_two(vectorDimension)
        baby = math.isqrt(dimension - 1) + 1 if dimension > 1 else 1   # ceil(sqrt(dimension))
        return BabyStepGiantStep(dimension, baby, -(-dimension // baby))


@dataclass(frozen=True)
class MatrixDimensions:
    """MatrixDimensions (PlaintextMatrix.swift:45-72)."""

    rowCount: int
    columnCount: int

    def __post_init__(self):
        if self.rowCount <= 0 or self.columnCount <= 0:
            raise PnnsError(f"invalidMatrixDimensions(rowCount: {self.rowCount}, columnCount: {self.columnCount})")


class GaloisElement:
    """GaloisElement.rotatingColumns / swappingRows (PolyRq/Galois.swift:174-212)."""

    @staticmethod
    def rotatingColumns(step: int, degree: int) -> int:
        positive = abs(step)
        if not 0 < positive < degree >> 1:
            raise HeError(-1, f"invalidRotationStep(step: {step}, degree: {degree})")
        if step > 0:
            positive = (degree >> 1) - positive
        return pow(3, positive, 2 * degree)

    @staticmethod
    def swappingRows(degree: int) -> int:
        return 2 * degree - 1


class SimdEncoder:
    """Context.encode(values:format: .simd) / decode (Encoding.swift:194-246) for a plaintext modulus t = 1 mod 2N,
    with the reference's choice of the minimal primitive 2N-th root of unity (PolyContext / NTT tables)."""

    def __init__(self, degree: int, plaintextModulus: int):
        n, t = int(degree), int(plaintextModulus)
        if (t - 1) % (2 * n) or t >= 1 << 32:
            raise HeError(-2, "simdEncodingNotSupported: t must be an NTT-friendly prime below 2^32")
        self.n, self.t, self.logn = n, t, n.bit_length() - 1
        self.psi = self._minimal_root()
        rev = self._bit_reverse(np.arange(n), self.logn)
        powers = np.ones(n, dtype=np.uint64)
        for i in range(1, n):
            powers[i] = int(powers[i - 1]) * self.psi % t
        self.roots = powers[rev]                                   # psi^bitrev(i), the merged-twiddle table
        inverse = pow(self.psi, -1, t)
        ipowers = np.ones(n, dtype=np.uint64)
        for i in range(1, n):
            ipowers[i] = int(ipowers[i - 1]) * inverse % t
        self.inverse_roots = ipowers[rev]
        half, mask = n >> 1, 2 * n - 1
        g, matrix = 1, np.zeros(n, dtype=np.int64)
        for i in range(half):
            matrix[i] = self._bit_reverse(np.array([(g - 1) >> 1]), self.logn)[0]
            matrix[half | i] = self._bit_reverse(np.array([(mask - g) >> 1]), self.logn)[0]
            g = g * 3 & mask
        self.encodingMatrix = matrix

    @staticmethod
    def _bit_reverse(x, bits):
        x = np.asarray(x, dtype=np.int64)
        out = np.zeros_like(x)
        for b in range(bits):
            out |= ((x >> b) & 1) << (bits - 1 - b)
        return out

    def _minimal_root(self) -> int:
        n, t = self.n, self.t
        root = next(r for r in (pow(x, (t - 1) // (2 * n), t) for x in range(2, t)) if pow(r, n, t) == t - 1)
        best, cur, sq = root, root, root * root % t
        for _ in range(n - 1):          # all primitive 2N-th roots are the odd powers of one of them
            cur = cur * sq % t
            best = min(best, cur)
        return best

    def forwardNtt(self, coeffs: np.ndarray) -> np.ndarray:
        """rows x N, natural order in -> bit-reversed Eval out (same convention as PolyRq.forwardNtt)."""
        a = np.array(coeffs, dtype=np.uint64).reshape(-1, self.n)
        t = np.uint64(self.t)
        m, span = 1, self.n >> 1
        while m < self.n:
            a = a.reshape(a.shape[0], m, 2, span)
            w = self.roots[m:2 * m].reshape(1, m, 1)
            v = a[:, :, 1, :] * w % t
            u = a[:, :, 0, :]
            a = np.stack([(u + v) % t, (u + t - v) % t], axis=2)
            m, span = m * 2, span >> 1
        return a.reshape(-1, self.n)

    def inverseNtt(self, evals: np.ndarray) -> np.ndarray:
        a = np.array(evals, dtype=np.uint64).reshape(-1, self.n)
        t = np.uint64(self.t)
        m, span = self.n >> 1, 1
        while m >= 1:
            a = a.reshape(a.shape[0], m, 2, span)
            w = self.inverse_roots[m:2 * m].reshape(1, m, 1)
            u, v = a[:, :, 0, :], a[:, :, 1, :]
            a = np.stack([(u + v) % t, (u + t - v) % t * w % t], axis=2)
            m, span = m >> 1, span * 2
        return a.reshape(-1, self.n) * np.uint64(pow(self.n, -1, self.t)) % t

    def encode(self, values: np.ndarray) -> np.ndarray:
        """rows x (<= N) SIMD values -> rows x N coefficient plaintexts."""
        v = np.asarray(values, dtype=np.uint64)
        v = v.reshape(-1, v.shape[-1])
        ev = np.zeros((v.shape[0], self.n), dtype=np.uint64)
        ev[:, self.encodingMatrix[: v.shape[1]]] = v % np.uint64(self.t)
        return self.inverseNtt(ev)

    def decode(self, plaintexts: np.ndarray) -> np.ndarray:
        return self.forwardNtt(plaintexts)[:, self.encodingMatrix]


class PlaintextMatrix:
    """PlaintextMatrix<Bfv<UInt64>, Eval> in .diagonal packing, resident in HBM."""

    def __init__(self, context: Context, dimensions: MatrixDimensions, values, babyStepGiantStep: BabyStepGiantStep = None,
                 plaintexts=None, evalFormat: bool = False):
        """values: the matrix in row-major order (packed here), or plaintexts: the already packed diagonal plaintexts
        (count x N coefficient rows, or count x L x N with evalFormat) as PlaintextMatrix.init(dimensions:packing:plaintexts:)."""
        self.context, self.dimensions = context, dimensions
        self.babyStepGiantStep = babyStepGiantStep or BabyStepGiantStep.forVectorDimension(dimensions.columnCount)
        if plaintexts is None:
            rows = PlaintextMatrix.diagonalPlaintexts(context, dimensions, self.babyStepGiantStep, values)
        else:
            rows = _host(plaintexts)
            expected = self.babyStepGiantStep.vectorDimension * -(-dimensions.rowCount // context.degree)
            if rows.size != expected * context.degree * (context.L if evalFormat else 1):
                raise PnnsError(f"wrongPlaintextCount(got: {rows.size // context.degree}, expected: {expected})")
        h = C.c_void_p()
        _check(load_library().hecuda_pnns_matrix_create(context._h, _ptr(rows), 1 if evalFormat else 0, dimensions.rowCount, dimensions.columnCount,
                                                       self.babyStepGiantStep.babyStep, self.babyStepGiantStep.giantStep,
                                                       C.byref(h)))
        self._h = h
        self.resultCiphertextCount = -(-dimensions.rowCount // context.degree)

    @staticmethod
    def diagonalPlaintexts(context, dimensions: MatrixDimensions, bsgs: BabyStepGiantStep, values) -> np.ndarray:
        """PlaintextMatrix.diagonalPlaintexts (PlaintextMatrix.swift:417-482) as coefficient rows (count x N).
        `context` only needs `degree` and `plaintextModulus`."""
        n, t = context.degree, context.plaintextModulus
        rows, cols = dimensions.rowCount, dimensions.columnCount
        if cols > n // 2:
            raise PnnsError(f"invalidMatrixDimensions(rowCount: {rows}, columnCount: {cols})")
        data = np.asarray(values, dtype=np.uint64).reshape(rows, cols)
        padded = _next_power_of_two(cols)
        wide = np.zeros((rows, padded), dtype=np.uint64)
        wide[:, :cols] = data
        # diagonal d holds data[c][(c + d) mod padded] for every database row c
        c = np.arange(rows)
        diagonals = np.stack([wide[c, (c + d) % padded] for d in range(padded)])
        per_column = -(-rows // n)
        full = np.zeros((padded, per_column * n), dtype=np.uint64)
        full[:, :rows] = diagonals
        chunks = full.reshape(padded, per_column, n)
        half = n // 2
        for d in range(padded):
            step = d // bsgs.babyStep * bsgs.babyStep
            if step:
                chunks[d] = np.concatenate([np.roll(chunks[d][:, :half], step, axis=1),
                                            np.roll(chunks[d][:, half:], step, axis=1)], axis=1)
        return SimdEncoder(n, t).encode(chunks.reshape(padded * per_column, n))

    def mulTranspose(self, vector, evaluationKey: EvaluationKey, modSwitchDownToSingle: bool = False) -> np.ndarray:
        """mulTranspose(vector:using:) for one (2, L, N) or a batch (batch, 2, L, N) of dense-row query ciphertexts.
        Returns (batch, resultCiphertextCount, 2, L or 1, N)."""
        ctx = self.context
        cts = _host(vector)
        words = 2 * ctx.L * ctx.degree
        if cts.size % words:
            raise HeError(-1, "invalidCiphertext: query vectors must be ciphertexts of 2 x L x N")
        batch = cts.size // words
        out = np.empty((batch, self.resultCiphertextCount, 2, 1 if modSwitchDownToSingle else ctx.L, ctx.degree), dtype=np.uint64)
        _check(load_library().hecuda_pnns_mul_transpose_vector(ctx._h, evaluationKey._h, self._h, _ptr(cts), batch,
                                                              1 if modSwitchDownToSingle else 0, _ptr(out)))
        return out

    def close(self):
        if getattr(self, "_h", None) is not None:
            load_library().hecuda_pnns_matrix_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def denseRowVector(context, vector) -> np.ndarray:
    """The SIMD values of a one-row `.denseRow` matrix (PlaintextMatrix.swift:341-413): the row padded to a power of
    two and repeated to fill both SIMD rows.  Returns the coefficient plaintext (N,)."""
    n, t = context.degree, context.plaintextModulus
    v = [int(x) % t for x in vector]
    packed = v + [0] * (_next_power_of_two(len(v)) - len(v))
    columns = n // 2
    if len(packed) < columns < len(packed) + len(v):
        packed += [0] * (columns - len(packed))
    offset = len(packed) % columns
    if offset:
        packed += [0] * (_next_power_of_two(offset) - offset)
    repeat = list(packed) if len(packed) <= columns else packed[columns:]
    while len(packed) < n:
        packed += repeat
    return SimdEncoder(n, t).encode(np.array(packed[:n], dtype=np.uint64))[0]
