"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's MulPir index-PIR protocol.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module; the
product (swift-homomorphic-encryption_b200/) never does.

Every function names the reference code it follows (paths relative to Sources/):

  * parameters        PrivateInformationRetrieval/IndexPir/MulPir.swift:37-109, IndexPirProtocol.swift:44-200
  * byte packing      HomomorphicEncryption/CoefficientPacking.swift:33-217
  * database layout   PrivateInformationRetrieval/IndexPir/MulPir.swift:433-556
  * query generation  MulPir.swift:181-223, PirUtil.swift:361-404
  * expansion         PirUtil.swift:204-355
  * response          PirUtil.swift:408-568 (computeResponseForOneChunk + computeResponse)
  * decryption        MulPir.swift:229-275

The homomorphic operations underneath are the C oracle's (oracle/he_oracle.c through oracle/oracle.py), which is pinned
on the reference's known-answer tests; this layer is pinned on the reference's MulPir KATs (evaluation-key
configurations, coordinates, uneven dimensions, CoefficientPacking vectors: tests/test_oracle_pir.py) and on the same
end-to-end property the reference tests use: the decrypted response equals the database entry.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

from . import oracle as O


# ------------------------------------------------------------------------------------------------ small helpers
def ceil_log2(x: int) -> int:
    return 0 if x <= 1 else (x - 1).bit_length()


def log2(x: int) -> int:
    return x.bit_length() - 1


def next_power_of_two(x: int) -> int:
    return 1 if x <= 1 else 1 << (x - 1).bit_length()


def dividing_ceil(a: int, b: int) -> int:
    return -(-a // b)


# ------------------------------------------------------------------------------------------------ CoefficientPacking
def bytes_to_coefficients_count(byte_count: int, bits_per_coeff: int, decode: bool, skip_lsbs: int = 0) -> int:
    """CoefficientPacking.bytesToCoefficientsCoeffCount (CoefficientPacking.swift:33-43)."""
    serialized = bits_per_coeff - skip_lsbs
    return 8 * byte_count // serialized if decode else dividing_ceil(8 * byte_count, serialized)


def bytes_to_coefficients(data, bits_per_coeff: int, decode: bool, skip_lsbs: int = 0) -> np.ndarray:
    """CoefficientPacking.bytesToCoefficients (:59-139): the bytes are one big-endian bit stream cut into fields of
    (bitsPerCoeff - skipLSBs) bits; a trailing partial field is padded with zero bits on the right."""
    raw = np.frombuffer(bytes(data), dtype=np.uint8)
    serialized = bits_per_coeff - skip_lsbs
    count = bytes_to_coefficients_count(len(raw), bits_per_coeff, decode, skip_lsbs)
    bits = np.unpackbits(raw)
    need = count * serialized
    if need > bits.size:
        bits = np.concatenate([bits, np.zeros(need - bits.size, dtype=np.uint8)])
    fields = bits[:need].reshape(count, serialized).astype(np.uint64)
    weights = (np.uint64(1) << np.arange(serialized - 1, -1, -1, dtype=np.uint64)).astype(np.uint64)
    return ((fields * weights).sum(axis=1, dtype=np.uint64) << np.uint64(skip_lsbs)).astype(np.uint64)


def coefficients_to_bytes(coeffs, bits_per_coeff: int, skip_lsbs: int = 0) -> bytes:
    """CoefficientPacking.coefficientsToBytes (:141-217)."""
    c = np.asarray(coeffs, dtype=np.uint64) >> np.uint64(skip_lsbs)
    serialized = bits_per_coeff - skip_lsbs
    shifts = np.arange(serialized - 1, -1, -1, dtype=np.uint64)
    bits = ((c[:, None] >> shifts[None, :]) & np.uint64(1)).astype(np.uint8).reshape(-1)
    byte_count = dividing_ceil(len(c) * serialized, 8)
    pad = byte_count * 8 - bits.size
    if pad:
        bits = np.concatenate([bits, np.zeros(pad, dtype=np.uint8)])
    return np.packbits(bits).tobytes()


# ------------------------------------------------------------------------------------------------ parameters
def entry_size_encoding_width(entry_size: int) -> int:
    """IndexPirConfig.entrySizeEncodingWidth(for:) (IndexPirProtocol.swift:107-118)."""
    if entry_size <= 0xFF:
        return 1
    if entry_size <= 0xFFFF:
        return 2
    if entry_size <= 0xFFFFFFFF:
        return 4
    return 8


@dataclass
class IndexPirConfig:
    """IndexPirConfig (IndexPirProtocol.swift:44-104)."""

    entry_count: int
    entry_size_in_bytes: int
    dimension_count: int = 2
    batch_size: int = 1
    uneven_dimensions: bool = False
    key_compression: str = "noCompression"  # | "hybridCompression" | "maxCompression"
    encoding_entry_size: bool = False

    def __post_init__(self):
        if self.dimension_count not in (1, 2):
            raise ValueError("invalidDimensionCount")

    @property
    def encoded_entry_size(self) -> int:
        width = entry_size_encoding_width(self.entry_size_in_bytes) if self.encoding_entry_size else 0
        return width + self.entry_size_in_bytes


@dataclass
class IndexPirParameter:
    """IndexPirParameter (IndexPirProtocol.swift:160-230)."""

    entry_count: int
    entry_size_in_bytes: int
    dimensions: list
    batch_size: int = 1
    galois_elements: list = field(default_factory=list)
    encoding_entry_size: bool = False

    @property
    def entry_size_encoding_width(self) -> int:
        return entry_size_encoding_width(self.entry_size_in_bytes) if self.encoding_entry_size else 0

    @property
    def encoded_entry_size(self) -> int:
        return self.entry_size_encoding_width + self.entry_size_in_bytes

    @property
    def expanded_query_count(self) -> int:
        return sum(self.dimensions)

    @property
    def per_chunk_plaintext_count(self) -> int:
        return math.prod(self.dimensions)


def evaluation_key_config(expanded_query_count: int, degree: int, key_compression: str) -> list:
    """MulPir.evaluationKeyConfig (MulPir.swift:86-109): the Galois elements 2^k + 1 the expansion needs."""
    max_depth = ceil_log2(min(expanded_query_count, degree))
    smallest = log2(degree) - max_depth + 1
    if key_compression == "noCompression":
        largest = log2(degree)
    else:
        largest = max(smallest, dividing_ceil(log2(degree) + 1, 2))
    elements = [(1 << level) + 1 for level in range(smallest, largest + 1)]
    if key_compression == "hybridCompression":
        extra = (1 << max(largest, (log2(degree) + largest + 1) // 2)) + 1
        if extra not in elements:
            elements.append(extra)
    return elements


def bytes_per_plaintext(degree: int, t: int) -> int:
    """EncryptionParameters.bytesPerPlaintext (EncryptionParameters.swift:103-110)."""
    return degree * log2(t) // 8


def generate_parameter(config: IndexPirConfig, degree: int, t: int) -> IndexPirParameter:
    """MulPir.generateParameter (MulPir.swift:37-83) for the BFV scheme."""
    bpp = bytes_per_plaintext(degree, t)
    encoded = config.encoded_entry_size
    per_chunk = dividing_ceil(config.entry_count, bpp // encoded) if encoded <= bpp else config.entry_count
    if config.dimension_count == 1:
        size = per_chunk
    else:
        size = math.isqrt(per_chunk)  # floor(root(x, 2))
    dims = [size] * config.dimension_count
    for i in range(len(dims)):
        if math.prod(dims) < per_chunk:
            dims[i] += 1
        else:
            break
    if config.uneven_dimensions and config.dimension_count == 2:
        limit = next_power_of_two(sum(dims) * config.batch_size)
        new = list(dims)
        while next_power_of_two(sum(new) * config.batch_size) <= limit:
            dims = list(new)
            if new[1] == 1:
                break
            new[1] -= 1
            new[0] = dividing_ceil(per_chunk, new[1])
    elements = evaluation_key_config(sum(dims) * config.batch_size, degree, config.key_compression)
    return IndexPirParameter(config.entry_count, config.entry_size_in_bytes, dims, config.batch_size, elements,
                             config.encoding_entry_size)


# ------------------------------------------------------------------------------------------------ database
@dataclass
class ProcessedDatabase:
    """ProcessedDatabase: `count` optional Eval plaintexts; plaintexts[i] is all-zero where present[i] == 0."""

    plaintexts: np.ndarray  # (count, L, N) uint64, Eval format
    present: np.ndarray     # (count,) uint8


def _encode_entry_size(size: int, width: int) -> bytes:
    return int(size).to_bytes(width, "little")  # IndexPirConfig.encodeEntrySize (IndexPirProtocol.swift:122-157)


def _plaintext_rows(ctx: O.Context, entries_bytes: list) -> list:
    """bytes of each plaintext -> coefficient vectors (None where every coefficient is zero)."""
    bits = log2(ctx.t)
    rows = []
    for chunk in entries_bytes:
        if chunk is None:
            rows.append(None)
            continue
        coeffs = bytes_to_coefficients(chunk, bits, decode=False)
        rows.append(None if not coeffs.any() else coeffs)
    return rows


def _to_eval(ctx: O.Context, rows: list):
    count = len(rows)
    pts = np.zeros((count, ctx.L, ctx.n), dtype=np.uint64)
    present = np.zeros(count, dtype=np.uint8)
    for i, coeffs in enumerate(rows):
        if coeffs is None:
            continue
        plain = np.zeros(ctx.n, dtype=np.uint64)
        plain[: len(coeffs)] = coeffs
        pts[i] = ctx.plaintext_to_eval(plain)
        present[i] = 1
    return ProcessedDatabase(pts, present)


def process_database(ctx: O.Context, param: IndexPirParameter, database: list) -> ProcessedDatabase:
    """MulPirServer.process (MulPir.swift:433-556): pack (or split) the entries into plaintexts, pad to the product of
    the dimensions and reorder so that each first-dimension column is contiguous."""
    assert len(database) == param.entry_count
    assert max((len(e) for e in database), default=0) <= param.entry_size_in_bytes
    bpp = bytes_per_plaintext(ctx.n, ctx.t)
    encoded = param.encoded_entry_size
    width = param.entry_size_encoding_width
    chunk_count = dividing_ceil(encoded, bpp)
    per_chunk = param.per_chunk_plaintext_count
    remaining = per_chunk // param.dimensions[0]
    if chunk_count > 1:  # processSplitLargeEntries (:453-504)
        table = []
        for entry in database:
            pieces = []
            for start in range(0, encoded, bpp):
                entry_start = start - width
                end = min(entry_start + bpp, len(entry))
                if not entry_start < end:
                    pieces.append(None)
                elif start == 0 and param.encoding_entry_size:
                    pieces.append(_encode_entry_size(len(entry), width) + bytes(entry[0:end]))
                else:
                    pieces.append(bytes(entry[entry_start:end]))
            table.append(_plaintext_rows(ctx, pieces))
        while len(table) < per_chunk:
            table.append([None] * chunk_count)
        flat = []
        for chunk in range(chunk_count):
            for skip in range(remaining):
                for row in range(skip, len(table), remaining):
                    flat.append(table[row][chunk])
        assert len(flat) == chunk_count * per_chunk
        return _to_eval(ctx, flat)
    # processPackEntries (:506-556)
    flat_bytes = bytearray()
    for entry in database:
        e = bytes(entry)
        if param.encoding_entry_size:
            e = _encode_entry_size(len(entry), width) + e
        flat_bytes += e + bytes(encoded - len(e))
    per_plaintext = (bpp // encoded) * encoded
    pieces = [bytes(flat_bytes[s:s + per_plaintext]) for s in range(0, len(flat_bytes), per_plaintext)]
    rows = _plaintext_rows(ctx, pieces)
    while len(rows) < per_chunk:
        rows.append(None)
    reordered = []
    for skip in range(remaining):
        for row in range(skip, len(rows), remaining):
            reordered.append(rows[row])
    return _to_eval(ctx, reordered)


# ------------------------------------------------------------------------------------------------ client: query
def entry_chunks_per_plaintext(ctx: O.Context, param: IndexPirParameter) -> int:
    bpp = bytes_per_plaintext(ctx.n, ctx.t)
    return bpp // param.encoded_entry_size if bpp >= param.encoded_entry_size else 1


def compute_coordinates(param: IndexPirParameter, index: int, chunks_per_plaintext: int) -> list:
    """MulPirClient.computeCoordinates (MulPir.swift:181-193)."""
    if not 0 <= index < param.entry_count:
        raise IndexError("invalidIndex")
    plaintext_index = index // chunks_per_plaintext
    product = math.prod(param.dimensions)
    coords = []
    for size in param.dimensions:
        product //= size
        coord = plaintext_index // product
        plaintext_index -= coord * product
        coords.append(coord)
    return coords


def compress_inputs_for_one_ciphertext(ctx: O.Context, total_input_count: int, one_indices: list) -> np.ndarray:
    """PirUtil.compressInputsForOneCiphertext (PirUtil.swift:361-377): a Coeff plaintext with 2^-ceilLog2(count) at
    the queried positions (the expansion doubles every level)."""
    assert total_input_count <= ctx.n
    raw = np.zeros(ctx.n, dtype=np.uint64)
    inverse = pow(pow(2, ceil_log2(total_input_count), ctx.t), -1, ctx.t)
    for index in one_indices:
        raw[index] = inverse
    return raw


def compress_binary_inputs(ctx: O.Context, total_input_count: int, one_indices: list, sk, seed: int) -> list:
    """PirUtil.compressBinaryInputs (PirUtil.swift:381-404)."""
    remaining, processed, cts = total_input_count, 0, []
    while remaining > 0:
        count = min(remaining, ctx.n)
        local = [x - processed for x in one_indices if processed <= x < processed + count]
        plain = compress_inputs_for_one_ciphertext(ctx, count, local)
        cts.append(ctx.encrypt(seed + len(cts), sk, plain))
        processed += count
        remaining -= count
    return cts


def generate_query(ctx: O.Context, param: IndexPirParameter, indices: list, sk, seed: int) -> list:
    """MulPirClient.generateQuery (MulPir.swift:201-218): list of Coeff ciphertexts (2, L, N)."""
    per_plaintext = entry_chunks_per_plaintext(ctx, param)
    accumulated, positions = 0, []
    for index in indices:
        coords = compute_coordinates(param, index, per_plaintext)
        for dim, size in enumerate(param.dimensions):
            positions.append(accumulated + coords[dim])
            accumulated += size
    return compress_binary_inputs(ctx, param.expanded_query_count * len(indices), positions, sk, seed)


# ------------------------------------------------------------------------------------------------ server: expansion
def _add(ctx, a, b, l):
    q = np.array(ctx.q[:l], dtype=np.uint64)[None, :, None]
    return (a + b) % q


def _sub(ctx, a, b, l):
    q = np.array(ctx.q[:l], dtype=np.uint64)[None, :, None]
    return (a + q - b) % q


def expand_ciphertext_for_one_step(ctx: O.Context, ct: np.ndarray, log_step: int, galois_keys: dict):
    """PirUtil.expandCiphertextForOneStep (PirUtil.swift:204-236)."""
    n, l = ctx.n, ct.shape[-2]
    assert log_step <= log2(n)
    shifting_power = 1 << (log_step - 1)
    target = (1 << (log2(n) - log_step + 1)) + 1
    candidates = [e for e in galois_keys if e <= target]
    if not candidates:
        raise KeyError("missingGaloisKey")
    element = max(candidates)
    count = 1 << (log2(target - 1) - log2(element - 1))
    c1, current = ct, 1
    for _ in range(count):
        c1 = ctx.apply_galois(c1, element, galois_keys[element], threads=1)[0]
        current = current * element % (2 * n)
    assert current == target
    difference = _sub(ctx, ct, c1, l)
    shifted = np.stack([O.multiply_power_of_x(n, ctx.q[:l], -shifting_power, difference[p]) for p in range(2)])
    return _add(ctx, c1, ct, l), shifted


def expand_ciphertext(ctx, ct, output_count: int, log_step: int, expected_height: int, galois_keys: dict) -> list:
    """PirUtil.expandCiphertext (PirUtil.swift:249-304)."""
    assert 0 <= output_count <= ctx.n
    l = ct.shape[-2]
    if output_count == 1:
        return [ct] if log_step > expected_height else [_add(ctx, ct, ct, l)]
    second_count = output_count >> 1
    first_count = output_count - second_count
    p0, p1 = expand_ciphertext_for_one_step(ctx, ct, log_step, galois_keys)
    first = expand_ciphertext(ctx, p0, first_count, log_step + 1, expected_height, galois_keys)
    second = expand_ciphertext(ctx, p1, second_count, log_step + 1, expected_height, galois_keys)
    out = []
    for a, b in zip(first[:second_count], second):
        out += [a, b]
    return out + first[second_count:]


def expand(ctx, ciphertexts: list, output_count: int, galois_keys: dict) -> list:
    """PirUtil.expand (PirUtil.swift:321-355)."""
    assert (len(ciphertexts) - 1) * ctx.n < output_count <= len(ciphertexts) * ctx.n
    remaining, out = output_count, []
    for ct in ciphertexts:
        count = min(remaining, ctx.n)
        remaining -= count
        out += expand_ciphertext(ctx, ct, count, 1, ceil_log2(count), galois_keys)
    return out


# ------------------------------------------------------------------------------------------------ server: response
def compute_response_for_one_chunk(ctx, dim0_query_eval: np.ndarray, remaining_query: list, chunk: ProcessedDatabase,
                                   relin_key, param: IndexPirParameter) -> np.ndarray:
    """PirUtil.computeResponseForOneChunk (PirUtil.swift:408-486) -> Coeff ciphertext (2, 1, N)."""
    n, L = ctx.n, ctx.L
    dim0 = dim0_query_eval.shape[0]
    columns = param.per_chunk_plaintext_count // param.dimensions[0]
    assert columns == 1 or columns == len(remaining_query)
    pts = chunk.plaintexts.reshape(columns, dim0, L, n)
    present = chunk.present.reshape(columns, dim0)
    sums = ctx.inner_product_plain(dim0_query_eval, pts, present)            # (columns, 2, L, N) Eval
    results = [np.stack([O.ntt_inverse(n, ctx.q, s[p]) for p in range(2)]) for s in sums]  # convertToCanonicalFormat
    start = 0
    for size in param.dimensions[1:]:
        vector0 = np.stack(remaining_query[start:start + size])
        nxt = []
        for group in range(0, len(results), size):
            vector1 = np.stack(results[group:group + size])
            product = ctx.inner_product(vector0[None], vector1[None], threads=1)[0]
            nxt.append(ctx.relinearize(product, relin_key, threads=1)[0])
        results = nxt
        start += size
    assert len(results) == 1
    ct = results[0]
    while ct.shape[-2] > 1:                                                  # modSwitchDownToSingle (HeScheme.swift:1481)
        ct = ctx.mod_switch_down(ct, threads=1)[0]
    return ct


def compute_response(ctx, query: list, indices_count: int, galois_keys: dict, relin_key, databases: list,
                     param: IndexPirParameter) -> list:
    """PirUtil.computeResponse (PirUtil.swift:490-568): response[query][chunk] = Coeff ciphertext (2, 1, N)."""
    if not (len(databases) == 1 or len(databases) >= indices_count):
        raise ValueError("invalidBatchSize")
    expanded = expand(ctx, query, param.expanded_query_count * indices_count, galois_keys)
    chunk_count = dividing_ceil(param.encoded_entry_size, bytes_per_plaintext(ctx.n, ctx.t))
    eqc = param.expanded_query_count
    out = []
    for qi in range(indices_count):
        cts = expanded[qi * eqc:(qi + 1) * eqc]
        db = databases[0 if len(databases) == 1 else qi]
        dim0 = param.dimensions[0]
        first = np.stack([np.stack([O.ntt_forward(ctx.n, ctx.q, ct[p]) for p in range(2)]) for ct in cts[:dim0]])
        rest = cts[dim0:]
        count = db.plaintexts.shape[0]
        per_chunk = count // chunk_count
        reply = []
        for start in range(0, count, per_chunk):
            chunk = ProcessedDatabase(db.plaintexts[start:start + per_chunk], db.present[start:start + per_chunk])
            reply.append(compute_response_for_one_chunk(ctx, first, rest, chunk, relin_key, param))
        out.append(reply)
    return out


# ------------------------------------------------------------------------------------------------ client: decrypt
def decrypt_response(ctx, param: IndexPirParameter, response: list, indices: list, sk) -> list:
    """MulPirClient.decrypt (MulPir.swift:245-275)."""
    bits = log2(ctx.t)
    per_plaintext = entry_chunks_per_plaintext(ctx, param)
    encoded = param.encoded_entry_size
    out = []
    for reply, index in zip(response, indices):
        data = b"".join(coefficients_to_bytes(ctx.decrypt(sk, ct), bits) for ct in reply)
        position = index % per_plaintext
        entry = data[position * encoded:(position + 1) * encoded]
        if param.encoding_entry_size:
            width = param.entry_size_encoding_width
            size = int.from_bytes(entry[:width], "little")
            entry = entry[width:][:size]
        out.append(entry)
    return out


# ------------------------------------------------------------------------------------------------ PolyRq wire format
def serialization_byte_count(n: int, moduli, skip_lsbs: int = 0) -> int:
    """PolyContext.serializationByteCount (PolyRq+Serialize.swift:86-96)."""
    return sum(dividing_ceil(n * (ceil_log2(int(q)) - skip_lsbs), 8) for q in moduli)


def serialize_poly(n: int, moduli, data, skip_lsbs: int = 0) -> bytes:
    """PolyRq.serialize(skipLSBs:) (PolyRq+Serialize.swift:67-84): rows packed at ceilLog2(q_i) bits, concatenated."""
    rows = np.asarray(data, dtype=np.uint64).reshape(len(moduli), n)
    return b"".join(coefficients_to_bytes(rows[i], ceil_log2(int(q)), skip_lsbs) for i, q in enumerate(moduli))


def load_poly(n: int, moduli, buffer: bytes, skip_lsbs: int = 0) -> np.ndarray:
    """PolyRq.load(from:skipLSBs:) (PolyRq+Serialize.swift:28-61)."""
    if len(buffer) != serialization_byte_count(n, moduli, skip_lsbs):
        raise ValueError("serializedBufferSizeMismatch")
    out, offset = np.zeros((len(moduli), n), dtype=np.uint64), 0
    for i, q in enumerate(moduli):
        bits = ceil_log2(int(q))
        count = dividing_ceil(n * (bits - skip_lsbs), 8)
        out[i] = bytes_to_coefficients(buffer[offset:offset + count], bits, decode=True, skip_lsbs=skip_lsbs)[:n]
        offset += count
    return out


def skip_lsbs_for_decryption(n: int, q0: int, t: int) -> list:
    """Bfv.skipLSBsForDecryption(for:) of a single-modulus ciphertext (Bfv+Decrypt.swift:51-110)."""
    l_prime = log2(q0 // t) - 3 if q0 >= 2 * t else 0
    tmp = int(8.0 * math.sqrt(2.0 * n / 9.0))
    poly0 = max(l_prime, 0)
    poly1 = l_prime - (0 if tmp == 0 else ceil_log2(tmp))
    if poly1 <= 1:
        poly0, poly1 = max(l_prime + 1, 0), 0
    return [poly0, poly1]
