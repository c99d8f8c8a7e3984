"""hecuda.pnns -- host-side mirror of the PNNS server's matrix-vector product over libhecuda.

    BabyStepGiantStep, MatrixDimensions        PrivateNearestNeighborSearch/MatrixMultiplication.swift:26-62,
                                               PlaintextMatrix.swift:45-72
    PlaintextMatrix (.diagonal packing)        PlaintextMatrix.swift:417-482
    PlaintextMatrix.mulTranspose(vector:using:) MatrixMultiplication.swift:131-226
    SIMD encoding                              HomomorphicEncryption/Encoding.swift:194-246

The diagonal packing and the SIMD encoding of the database are host work in the reference too (offline
preprocessing); conversion to Eval format, the rotations, inner products and modulus switching run on the device.
`mulTransposeMatrix` runs the whole mulTranspose(matrix:) -- extractDenseRow, products, dense-column packing -- in one
C-ABI call; the shape logic extractDenseRow derives per row (mask, ciphertext index, replication count) is computed here.
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

    @staticmethod
    def stepsFor(elements, degree: int) -> dict:
        """GaloisElement.stepsFor (PolyRq/Galois.swift:239-258): 3^k mod 2N  <->  rotation by N/2 - k."""
        wanted, out, g = set(int(e) for e in elements), {}, 1
        for k in range(degree // 2 + 1):
            if g in wanted and g not in out:
                out[g] = degree // 2 - k
            g = g * 3 % (2 * degree)
        return {e: out.get(e) for e in wanted}

    @staticmethod
    def planMultiStep(supportedSteps, step: int, degree: int):
        """GaloisElement._planMultiStep (PolyRq/Galois.swift:272-319): greedy decomposition, by steps or by their
        complements to N/2, whichever needs fewer rotations."""
        if abs(step) >= degree:
            raise HeError(-1, f"invalidRotationStep(step: {step}, degree: {degree})")
        if step in supportedSteps:
            return {step: 1}

        def greedy(order, weight):
            left, plan = weight(step), {}
            for s in order:
                w = weight(s)
                if left // w:
                    plan[s] = plan.get(s, 0) + left // w
                left %= w
            return plan if left == 0 else None

        down = sorted(supportedSteps, reverse=True)
        forward = greedy(down, lambda s: s)
        backward = greedy(down[::-1], lambda s: (degree >> 1) - s)
        if forward is None or backward is None:
            return forward if backward is None else backward
        return forward if sum(forward.values()) <= sum(backward.values()) else backward

    @staticmethod
    def rotationSequence(galoisElements, step: int, degree: int) -> list:
        """The single rotations of rotateColumnsMultiStep(by: step) (_HomomorphicEncryptionExtras/HeScheme.swift:65-104),
        larger steps first (the reference walks a Dictionary, i.e. in unspecified order)."""
        if step == 0:
            return []
        if GaloisElement.rotatingColumns(step, degree) in galoisElements:
            return [step]
        steps = [s for s in GaloisElement.stepsFor(galoisElements, degree).values() if s is not None]
        plan = GaloisElement.planMultiStep(steps, step + degree // 2 if step < 0 else step, degree)
        if plan is None:
            raise HeError(-1, f"invalidRotationStep(step: {step}, degree: {degree})")
        return [s for s in sorted(plan, reverse=True) for _ in range(plan[s])]


class CiphertextMatrix:
    """The shape logic of CiphertextMatrix in `.denseRow` packing (CiphertextMatrix.swift, PlaintextMatrix.swift:262-268)."""

    @staticmethod
    def ciphertextCount(degree: int, dimensions: MatrixDimensions) -> int:
        per_ciphertext = 2 * ((degree // 2) // _next_power_of_two(dimensions.columnCount))
        return -(-dimensions.rowCount // per_ciphertext)

    @staticmethod
    def denseRowExtraction(degree: int, dimensions: MatrixDimensions, ciphertextCount: int, rowIndex: int):
        """What extractDenseRow derives for one row (CiphertextMatrix.swift:254-320): (index of the ciphertext holding
        the row, SIMD values of the plaintext mask, number of rotate-and-add replication steps)."""
        columns = degree // 2
        width = _next_power_of_two(dimensions.columnCount)
        per_ciphertext = 2 * (columns // width)
        ciphertext_index = rowIndex // per_ciphertext
        is_last = ciphertext_index == ciphertextCount - 1

        def slots(row):
            lo = (row % per_ciphertext) * width
            hi = lo + width
            if lo <= columns < hi:                       # the batch would straddle the two SIMD rows
                lo, hi = columns, columns + width
            elif hi > columns:                           # second SIMD row: skip the first row's padding
                lo, hi = lo + columns % width, hi + columns % width
            if is_last:                                  # the last ciphertext repeats its rows to the end
                hi = -(-hi // columns) * columns
            return lo, hi

        lo, hi = slots(rowIndex)
        after = rowIndex + 1
        while after < dimensions.rowCount and slots(after)[1] == hi:
            after += 1
        before = max(rowIndex - 1, 0)
        while before > 0 and slots(before)[1] == hi:
            before -= 1
        period = _next_power_of_two(width * (after - before))
        mask = [0] * lo
        copies = 0
        while len(mask) < hi:
            mask += [1] * width + [0] * (period - width)
            copies += 1
        return ciphertext_index, mask[:degree], columns // (copies * width) - 1


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

    def mulTransposeMatrix(self, ciphertexts, queryDimensions: MatrixDimensions, evaluationKey: EvaluationKey,
                           modSwitchDownToSingle: bool = False) -> np.ndarray:
        """mulTranspose(matrix:using:) (MatrixMultiplication.swift:236-298): `ciphertexts` is the dense-row packed query
        CiphertextMatrix (count, 2, L, N); returns the dense-column packed result ciphertexts (count', 2, L or 1, N)."""
        ctx = self.context
        n = ctx.degree
        if queryDimensions.columnCount != self.dimensions.columnCount:
            raise PnnsError(f"invalidMatrixDimensions(rowCount: {queryDimensions.rowCount}, columnCount: {queryDimensions.columnCount})")
        cts = _host(ciphertexts)
        words = 2 * ctx.L * n
        count = cts.size // words
        rows = queryDimensions.rowCount
        expected = CiphertextMatrix.ciphertextCount(n, queryDimensions)
        if cts.size % words or count != expected:
            raise PnnsError(f"wrongCiphertextCount(got: {count}, expected: {expected})")
        index = (C.c_int32 * rows)()
        rotate = (C.c_int32 * rows)()
        masks = np.zeros((rows, n), dtype=np.uint64)
        step = _next_power_of_two(queryDimensions.columnCount)
        if rows > 1:
            encoder = SimdEncoder(n, ctx.plaintextModulus)
            values = np.zeros((rows, n), dtype=np.uint64)
            for r in range(rows):
                index[r], mask, rotate[r] = CiphertextMatrix.denseRowExtraction(n, queryDimensions, count, r)
                values[r, : len(mask)] = mask
            masks = encoder.encode(values)
        per_simd_row = (n // 2) // self.dimensions.rowCount
        sequence = []
        if per_simd_row > 1 and rows > 1:
            sequence = GaloisElement.rotationSequence(evaluationKey.galoisElements, self.dimensions.rowCount, n)
        plan = (C.c_int32 * max(1, len(sequence)))(*sequence)
        capacity = rows * self.resultCiphertextCount
        out = np.empty((capacity, 2, 1 if modSwitchDownToSingle else ctx.L, n), dtype=np.uint64)
        produced = C.c_int64(0)
        _check(load_library().hecuda_pnns_mul_transpose_matrix(
            ctx._h, evaluationKey._h, self._h, _ptr(cts), count, rows, index, _ptr(masks), rotate, step, plan, len(sequence),
            1 if modSwitchDownToSingle else 0, _ptr(out), capacity, C.byref(produced)))
        return out[: produced.value]

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
