"""hecuda.pir -- host-side mirror of the reference's MulPir index-PIR server over libhecuda.

Names and argument meaning follow Sources/PrivateInformationRetrieval/IndexPir:

    IndexPirConfig, IndexPirParameter            IndexPirProtocol.swift:44-230
    MulPir.generateParameter / evaluationKeyConfig  MulPir.swift:37-109
    CoefficientPacking.bytesToCoefficients / coefficientsToBytes   HomomorphicEncryption/CoefficientPacking.swift
    MulPirServer.process / computeResponse       MulPir.swift:412-556, PirUtil.swift:490-568
    PirUtil.expand                               PirUtil.swift:321-355

Only the server side lives here (the client's encrypt / decrypt are SURVEY.md 8f rank 4).  The database stays resident in
HBM (`ProcessedDatabase`); `computeResponse` is one C-ABI call per query.  No CPU fallback: everything that touches
ciphertexts or plaintext polynomials runs in libhecuda.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from enum import Enum
from typing import List, Optional, Sequence

import numpy as np

from . import Context, EvaluationKey, HeError, _check, _host, _ptr, load_library


class PirKeyCompressionStrategy(str, Enum):
    """PirKeyCompressionStrategy (IndexPirProtocol.swift:32-41)."""

    hybridCompression = "hybridCompression"
    maxCompression = "maxCompression"
    noCompression = "noCompression"


class PirError(ValueError):
    pass


def _ceil_log2(value: int) -> int:
    return max(0, (int(value) - 1).bit_length())


def _entry_size_encoding_width(entry_size: int) -> int:
    for width, limit in ((1, 1 << 8), (2, 1 << 16), (4, 1 << 32)):
        if entry_size < limit:
            return width
    return 8


class CoefficientPacking:
    """enum CoefficientPacking (CoefficientPacking.swift): big-endian bit-stream <-> fixed-width coefficients."""

    @staticmethod
    def bytesToCoefficients(data: bytes, bitsPerCoeff: int, decode: bool, skipLSBs: int = 0) -> np.ndarray:
        width = bitsPerCoeff - skipLSBs
        if not (bitsPerCoeff > 0 and width > 0 and skipLSBs >= 0):
            raise HeError(-1, f"invalidCoefficientPacking(bitsPerCoeff: {bitsPerCoeff}, skipLSBs: {skipLSBs})")
        total_bits = 8 * len(data)
        count = total_bits // width if decode else -(-total_bits // width)
        stream = int.from_bytes(bytes(data), "big")
        padded = count * width
        stream = stream << (padded - total_bits) if padded >= total_bits else stream >> (total_bits - padded)
        mask = (1 << width) - 1
        out = np.empty(count, dtype=np.uint64)
        for i in range(count - 1, -1, -1):
            out[i] = (stream & mask) << skipLSBs
            stream >>= width
        return out

    @staticmethod
    def coefficientsToBytes(coeffs: Sequence[int], bitsPerCoeff: int, skipLSBs: int = 0) -> bytes:
        width = bitsPerCoeff - skipLSBs
        if not (bitsPerCoeff > 0 and width > 0 and skipLSBs >= 0):
            raise HeError(-1, f"invalidCoefficientPacking(bitsPerCoeff: {bitsPerCoeff}, skipLSBs: {skipLSBs})")
        stream = 0
        for value in coeffs:
            stream = (stream << width) | ((int(value) >> skipLSBs) & ((1 << width) - 1))
        bits = len(coeffs) * width
        byte_count = -(-bits // 8)
        return (stream << (8 * byte_count - bits)).to_bytes(byte_count, "big")


def _pack_rows(pieces: List[Optional[bytes]], bits: int, degree: int):
    """Many byte strings -> (len(pieces), N) coefficient rows + presence flags (vectorised bytesToCoefficients)."""
    rows = np.zeros((len(pieces), degree), dtype=np.uint64)
    present = np.zeros(len(pieces), dtype=np.uint8)
    weights = (np.uint64(1) << np.arange(bits - 1, -1, -1, dtype=np.uint64))
    for i, piece in enumerate(pieces):
        if not piece:
            continue
        stream = np.unpackbits(np.frombuffer(piece, dtype=np.uint8))
        count = -(-stream.size // bits)
        if count > degree:
            raise PirError("plaintext bytes exceed bytesPerPlaintext")
        if count * bits != stream.size:
            stream = np.concatenate([stream, np.zeros(count * bits - stream.size, dtype=np.uint8)])
        values = stream.reshape(count, bits).astype(np.uint64) @ weights
        if values.any():
            rows[i, :count] = values
            present[i] = 1
    return rows, present


@dataclass
class IndexPirConfig:
    """IndexPirConfig (IndexPirProtocol.swift:44-104)."""

    entryCount: int
    entrySizeInBytes: int
    dimensionCount: int
    batchSize: int
    unevenDimensions: bool
    keyCompression: PirKeyCompressionStrategy
    encodingEntrySize: bool = False

    def __post_init__(self):
        if self.dimensionCount not in (1, 2):
            raise PirError(f"invalidDimensionCount(dimensionCount: {self.dimensionCount}, expected: [1, 2])")
        self.keyCompression = PirKeyCompressionStrategy(self.keyCompression)

    @property
    def entrySizeEncodingWidth(self) -> int:
        return _entry_size_encoding_width(self.entrySizeInBytes) if self.encodingEntrySize else 0

    @property
    def encodedEntrySize(self) -> int:
        return self.entrySizeEncodingWidth + self.entrySizeInBytes


@dataclass
class EvaluationKeyConfig:
    """EvaluationKeyConfig (Keys.swift:222): which Galois keys and whether a relinearization key are needed."""

    galoisElements: List[int] = field(default_factory=list)
    hasRelinearizationKey: bool = False


@dataclass
class IndexPirParameter:
    """IndexPirParameter (IndexPirProtocol.swift:160-230)."""

    entryCount: int
    entrySizeInBytes: int
    dimensions: List[int]
    batchSize: int
    evaluationKeyConfig: EvaluationKeyConfig
    encodingEntrySize: bool = False

    @property
    def entrySizeEncodingWidth(self) -> int:
        return _entry_size_encoding_width(self.entrySizeInBytes) if self.encodingEntrySize else 0

    @property
    def encodedEntrySize(self) -> int:
        return self.entrySizeEncodingWidth + self.entrySizeInBytes

    @property
    def dimensionCount(self) -> int:
        return len(self.dimensions)

    @property
    def expandedQueryCount(self) -> int:
        return int(sum(self.dimensions))


def bytesPerPlaintext(context) -> int:
    """Context.bytesPerPlaintext (EncryptionParameters.swift:103-110)."""
    return context.degree * (context.plaintextModulus.bit_length() - 1) // 8


class MulPir:
    """enum MulPir<Bfv<UInt64>>: IndexPirProtocol (MulPir.swift:24-115)."""

    @staticmethod
    def evaluationKeyConfig(expandedQueryCount: int, degree: int,
                            keyCompression: PirKeyCompressionStrategy) -> EvaluationKeyConfig:
        log_degree = degree.bit_length() - 1
        depth = _ceil_log2(min(expandedQueryCount, degree))
        smallest = log_degree - depth + 1
        compression = PirKeyCompressionStrategy(keyCompression)
        largest = log_degree if compression is PirKeyCompressionStrategy.noCompression else max(smallest, (log_degree + 2) // 2)
        powers = list(range(smallest, largest + 1))
        if compression is PirKeyCompressionStrategy.hybridCompression:
            extra = max(largest, (log_degree + largest + 1) // 2)
            if extra not in powers:
                powers.append(extra)
        return EvaluationKeyConfig([(1 << k) + 1 for k in powers], True)

    @staticmethod
    def generateParameter(config: IndexPirConfig, context) -> IndexPirParameter:
        per_plaintext_bytes = bytesPerPlaintext(context)
        encoded = config.encodedEntrySize
        if encoded <= per_plaintext_bytes:
            plaintexts = -(-config.entryCount // (per_plaintext_bytes // encoded))
        else:
            plaintexts = config.entryCount
        side = plaintexts
        if config.dimensionCount == 2:
            side = int(np.floor(np.sqrt(float(plaintexts))))
            while side * side > plaintexts:      # guard the floating-point root the way floor(root(x, 2)) behaves
                side -= 1
            while (side + 1) * (side + 1) <= plaintexts:
                side += 1
        dims = [side] * config.dimensionCount
        for i in range(len(dims)):
            if int(np.prod(dims, dtype=np.int64)) >= plaintexts:
                break
            dims[i] += 1
        if config.unevenDimensions and config.dimensionCount == 2:   # BFV only (MulPir.swift:56-71)
            def pow2(v):
                return 1 << _ceil_log2(v)
            limit = pow2(sum(dims) * config.batchSize)
            trial = list(dims)
            while pow2(sum(trial) * config.batchSize) <= limit:
                dims = list(trial)
                if trial[1] == 1:
                    break
                trial[1] -= 1
                trial[0] = -(-plaintexts // trial[1])
        evk = MulPir.evaluationKeyConfig(sum(dims) * config.batchSize, context.degree, config.keyCompression)
        return IndexPirParameter(config.entryCount, config.entrySizeInBytes, dims, config.batchSize, evk,
                                 config.encodingEntrySize)


def skipLSBsForDecryption(context) -> List[int]:
    """Bfv.skipLSBsForDecryption(for:) of a single-modulus ciphertext (Bfv+Decrypt.swift:51-110): how many low bits of
    poly 0 / poly 1 a reply may drop.  `context` needs degree, plaintextModulus, coefficientModuli."""
    q0, t = int(context.coefficientModuli[0]), int(context.plaintextModulus)
    l_prime = (q0 // t).bit_length() - 1 - 3 if q0 >= 2 * t else 0
    spread = int(8.0 * (2.0 * context.degree / 9.0) ** 0.5)
    poly0, poly1 = max(l_prime, 0), l_prime - (_ceil_log2(spread) if spread else 0)
    if poly1 <= 1:
        poly0, poly1 = max(l_prime + 1, 0), 0
    return [poly0, poly1]


class ProcessedDatabase:
    """ProcessedDatabase<Bfv<UInt64>> resident in HBM (IndexPirDatabase.swift): `count` optional Eval plaintexts."""

    def __init__(self, context: Context, plaintexts, present=None, evalFormat: bool = False):
        self.context = context
        rows = _host(plaintexts)
        words = context.degree * (context.L if evalFormat else 1)
        if rows.size % words:
            raise PirError("plaintext buffer has the wrong shape")
        self.count = rows.size // words
        flags = None
        if present is not None:
            flags = np.ascontiguousarray(np.asarray(present, dtype=np.uint8))
            if flags.size != self.count:
                raise PirError("presence flags have the wrong length")
        h = C.c_void_p()
        _check(load_library().hecuda_pir_database_create(context._h, _ptr(rows), 1 if evalFormat else 0,
                                                        flags.ctypes.data_as(C.c_void_p) if flags is not None else None,
                                                        self.count, C.byref(h)))
        self._h = h

    def close(self):
        if getattr(self, "_h", None) is not None:
            load_library().hecuda_pir_database_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class PirWire:
    """Request / reply bytes of the index-PIR server (the payloads of the reference's protobuf messages)."""

    @staticmethod
    def computeResponse(server: "MulPirServer", queryPoly0, querySeeds, evaluationKey: EvaluationKey, indicesCount: int = 1):
        """Serialized seeded query ciphertexts in, serialized (skipLSBsForDecryption) reply ciphertexts out, one C-ABI call.
        queryPoly0: (count, byteCount(L rows)) uint8; querySeeds: (count, 32) uint8.  Returns (replies, skipLSBs) with
        replies of shape (indicesCount, chunkCount, bytes(poly0) + bytes(poly1))."""
        from . import Bfv
        ctx = server.context
        seeds = np.ascontiguousarray(np.asarray(querySeeds, dtype=np.uint8)).reshape(-1, 32)
        poly0 = np.ascontiguousarray(np.asarray(queryPoly0, dtype=np.uint8)).reshape(seeds.shape[0], -1)
        if poly0.shape[1] != Bfv.serializationByteCount(ctx, ctx.L):
            raise HeError(-1, "serializedBufferSizeMismatch")
        skips = skipLSBsForDecryption(ctx)
        sizes = [Bfv.serializationByteCount(ctx, 1, s) for s in skips]
        out = np.empty((indicesCount, server.chunkCount, sum(sizes)), dtype=np.uint8)
        handles = (C.c_void_p * len(server.databases))(*[db._h for db in server.databases])
        dims = (C.c_int32 * len(server.parameter.dimensions))(*server.parameter.dimensions)
        _check(load_library().hecuda_mulpir_compute_response_wire(
            ctx._h, evaluationKey._h, handles, len(server.databases), dims, len(server.parameter.dimensions), server.chunkCount,
            poly0.ctypes.data_as(C.c_void_p), seeds.ctypes.data_as(C.c_void_p), seeds.shape[0], indicesCount, skips[0], skips[1],
            out.ctypes.data_as(C.c_void_p)))
        return out, skips


class PirUtil:
    """enum PirUtil<Bfv<UInt64>> (PirUtil.swift:573): expansion and response computation on the device."""

    @staticmethod
    def expand(context: Context, ciphertexts, outputCount: int, evaluationKey: EvaluationKey) -> np.ndarray:
        cts = _host(ciphertexts)
        words = 2 * context.L * context.degree
        count = cts.size // words
        out = np.empty((outputCount, 2, context.L, context.degree), dtype=np.uint64)
        _check(load_library().hecuda_mulpir_expand(context._h, evaluationKey._h, _ptr(cts), count, outputCount, _ptr(out)))
        return out


class MulPirServer:
    """MulPirServer<PirUtil<Bfv<UInt64>>> (MulPir.swift:292-426)."""

    def __init__(self, parameter: IndexPirParameter, context: Context, databases: Sequence[ProcessedDatabase]):
        self.parameter, self.context, self.databases = parameter, context, list(databases)
        expected = self.chunkCount * int(np.prod(parameter.dimensions, dtype=np.int64))
        for db in self.databases:
            if db.count != expected:
                raise PirError(f"invalidDatabasePlaintextCount(plaintextCount: {db.count}, expected: {expected})")

    @property
    def chunkCount(self) -> int:
        return -(-self.parameter.encodedEntrySize // bytesPerPlaintext(self.context))

    @staticmethod
    def process(database: Sequence[bytes], context: Context, parameter: IndexPirParameter) -> ProcessedDatabase:
        """MulPirServer.process (MulPir.swift:433-556): bytes -> coefficient plaintexts (host), Eval conversion (device)."""
        rows, present = MulPirServer.plaintextRows(database, context, parameter)
        return ProcessedDatabase(context, rows, present, evalFormat=False)

    @staticmethod
    def plaintextRows(database: Sequence[bytes], context, parameter: IndexPirParameter):
        """The host half of `process`: the coefficient vector of every plaintext slot, in database order, and which slots
        are non-nil.  `context` only needs `degree` and `plaintextModulus`."""
        if len(database) != parameter.entryCount:
            raise PirError(f"invalidDatabaseEntryCount(entryCount: {len(database)}, expected: {parameter.entryCount})")
        longest = max((len(e) for e in database), default=0)
        if longest > parameter.entrySizeInBytes:
            raise PirError(f"invalidDatabaseEntrySize(maximumEntrySize: {longest}, expected: {parameter.entrySizeInBytes})")
        capacity = bytesPerPlaintext(context)
        bits = context.plaintextModulus.bit_length() - 1
        encoded, width = parameter.encodedEntrySize, parameter.entrySizeEncodingWidth
        chunks = -(-encoded // capacity)
        per_chunk = int(np.prod(parameter.dimensions, dtype=np.int64))
        columns = per_chunk // parameter.dimensions[0]
        if chunks > 1:                                     # processSplitLargeEntries
            grid = [[None] * chunks for _ in range(per_chunk)]
            for row, entry in enumerate(database):
                entry = bytes(entry)
                blob = (len(entry).to_bytes(width, "little") if width else b"") + entry
                for chunk in range(chunks):
                    lo = chunk * capacity
                    # the reference cuts [lo - width, lo - width + capacity) out of the entry, prefixing the size to chunk 0
                    hi = min(lo - width + capacity, len(entry)) + width
                    if lo - width < hi - width:
                        grid[row][chunk] = blob[lo:hi] if chunk else blob[:hi]
            order = [(row, chunk) for chunk in range(chunks) for skip in range(columns)
                     for row in range(skip, per_chunk, columns)]
            pieces = [grid[row][chunk] for row, chunk in order]
        else:                                              # processPackEntries
            stride = (capacity // encoded) * encoded
            flat = bytearray()
            for entry in database:
                entry = bytes(entry)
                body = (len(entry).to_bytes(width, "little") if width else b"") + entry
                flat += body.ljust(encoded, b"\0")
            packed = [bytes(flat[i:i + stride]) for i in range(0, len(flat), stride)]
            packed += [None] * (per_chunk - len(packed))
            pieces = [packed[row] for skip in range(columns) for row in range(skip, per_chunk, columns)]
        return _pack_rows(pieces, bits, context.degree)

    def computeResponse(self, query, evaluationKey: EvaluationKey, indicesCount: int = 1) -> np.ndarray:
        """computeResponse(to:using:) -> Response.ciphertexts as (indicesCount, chunkCount, 2, 1, N) (Coeff, modulus q_0).

        query: Query.ciphertexts stacked, (queryCiphertextCount, 2, L, N) Coeff."""
        ctx = self.context
        cts = _host(query)
        words = 2 * ctx.L * ctx.degree
        if cts.size % words:
            raise HeError(-1, "invalidCiphertext: query must be ciphertexts of 2 x L x N")
        handles = (C.c_void_p * len(self.databases))(*[db._h for db in self.databases])
        dims = (C.c_int32 * len(self.parameter.dimensions))(*self.parameter.dimensions)
        out = np.empty((indicesCount, self.chunkCount, 2, 1, ctx.degree), dtype=np.uint64)
        _check(load_library().hecuda_mulpir_compute_response(
            ctx._h, evaluationKey._h, handles, len(self.databases), dims, len(self.parameter.dimensions), self.chunkCount,
            _ptr(cts), cts.size // words, indicesCount, _ptr(out)))
        return out
