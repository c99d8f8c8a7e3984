// pnns.cu -- PNNS server: encrypted-vector x plaintext-matrix product on one device.
//
//   PlaintextMatrix.mulTranspose(vector:using:)    PrivateNearestNeighborSearch/MatrixMultiplication.swift:131-226
//   BabyStepGiantStep                              MatrixMultiplication.swift:26-62
//   rotateColumnsAndSum                            _HomomorphicEncryptionExtras/HeScheme.swift:113-134
//   Server.computeResponse post-processing         PrivateNearestNeighborSearch/Server.swift:61-88
//
// The matrix stays in HBM in Eval format, re-ordered once so that every (result ciphertext, giant step) pair owns
// `babyStep` consecutive plaintexts (absent ones flagged), which makes step 2 of the algorithm a single launch of the
// streaming ct x pt inner-product kernel per query vector.  Query vectors that share an evaluation key are batched
// through the rotation chains (babyStep - 1 rotations by -1, giantStep - 1 rotations by -babyStep).
#include <algorithm>

#include "capi_internal.hpp"

using namespace hecuda;
using namespace hecuda::api;

struct hecuda_pnns_matrix {
    const hecuda_context *owner = nullptr;
    u64 *d_plain = nullptr;              // [result][giant][baby] x L x N, Eval
    unsigned char *d_present = nullptr;  // [result][giant][baby]
    int64_t row_count = 0, column_count = 0, result_count = 0;
    int baby = 0, giant = 0, dimension = 0;
};

namespace {

struct RowConsts {
    int rows;
    u64 p[kMaxRows];
};

// acc[item] (+)= src[item * src_stride]   over ciphertexts of 2 x rows x N
// modes (optional, per item): 0 = leave acc alone, 1 = copy, 2 = add; without modes every item uses `add`
__global__ void __launch_bounds__(256) accumulate_kernel(u64 *__restrict__ acc, const u64 *__restrict__ src,
                                                        int64_t src_item_stride, const __grid_constant__ RowConsts c,
                                                        int n, int add, const signed char *__restrict__ modes) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    if (modes) {
        const int mode = modes[blockIdx.z];
        if (mode == 0) return;
        add = mode == 2;
    }
    const int pr = blockIdx.y;
    const int64_t ct_words = (int64_t)2 * c.rows * n;
    const int64_t off = (int64_t)pr * n + e;
    const u64 v = src[(int64_t)blockIdx.z * src_item_stride + off];
    u64 *dst = acc + (int64_t)blockIdx.z * ct_words + off;
    if (add) {
        const u64 p = c.p[pr % c.rows];
        const u64 s = *dst + v;
        *dst = s >= p ? s - p : s;
    } else {
        *dst = v;
    }
}

cudaError_t launch_accumulate(const Context &c, int l, u64 *acc, const u64 *src, int64_t src_item_stride, int64_t items,
                              bool add, cudaStream_t s, const signed char *modes = nullptr) {
    RowConsts rc;
    const NttRowMap map = c.map_q(l);
    rc.rows = l;
    for (int r = 0; r < l; ++r) rc.p[r] = c.slots[map.slot[r]].dev.p;
    const int threads = c.n >= 256 ? 256 : (c.n < 32 ? 32 : (int)c.n);
    const int64_t ct_words = (int64_t)2 * l * c.n;
    for (int64_t done = 0; done < items;) {
        const int64_t chunk = std::min<int64_t>(items - done, 65535);
        dim3 grid((unsigned)((c.n + threads - 1) / threads), (unsigned)(2 * l), (unsigned)chunk);
        ++g_kernel_launches;
        accumulate_kernel<<<grid, threads, 0, s>>>(acc + done * ct_words, src + done * src_item_stride, src_item_stride, rc,
                                                   (int)c.n, add ? 1 : 0, modes ? modes + done : nullptr);
        done += chunk;
    }
    return cudaGetLastError();
}

// GaloisElement.rotatingColumns(by:degree:) (PolyRq/Galois.swift:195-212)
unsigned rotating_columns(int step, int64_t degree) {
    unsigned positive = (unsigned)(step < 0 ? -step : step);
    if (step > 0) positive = (unsigned)(degree >> 1) - positive;
    unsigned long long g = 1, base = 3, mod = 2ull * (unsigned long long)degree;
    for (unsigned e = positive; e; e >>= 1) {
        if (e & 1) g = g * base % mod;
        base = base * base % mod;
    }
    return (unsigned)g;
}

int32_t find_key(const hecuda_evk *k, unsigned element, const u64 **key) {
    hecuda_evk *km = const_cast<hecuda_evk *>(k);
    std::lock_guard<std::mutex> g(km->mu);
    auto it = km->galois.find(element);
    if (it == km->galois.end()) return fail(HECUDA_ERR_MISSING_KEY, "missingGaloisElement: " + std::to_string(element));
    *key = it->second;
    return HECUDA_OK;
}

struct Tmp {
    cudaStream_t s;
    std::vector<void *> ptrs;
    explicit Tmp(cudaStream_t stream) : s(stream) {}
    ~Tmp() {
        for (void *p : ptrs) cudaFreeAsync(p, s);
    }
    cudaError_t alloc(u64 **out, size_t words) {
        cudaError_t e = cudaMallocAsync((void **)out, std::max<size_t>(words, 1) * sizeof(u64), s);
        if (e == cudaSuccess) ptrs.push_back(*out);
        return e;
    }
};

int32_t mul_transpose_device(const hecuda_context *h, const hecuda_evk *k, const hecuda_pnns_matrix *m, const u64 *d_vec,
                             int64_t batch, bool to_single, u64 *d_out, cudaStream_t s) {
    const Context &c = *h->ctx;
    const int L = c.L;
    const int64_t n = c.n;
    const size_t ct_words = (size_t)2 * L * n;
    const int baby = m->baby, giant = m->giant;
    const int64_t results = m->result_count;
    const u64 *key1 = nullptr, *keyb = nullptr;
    const unsigned e1 = n > 2 ? rotating_columns(-1, n) : 0, eb = rotating_columns(-baby, n);
    int32_t rc;
    if (baby > 1 && (rc = find_key(k, e1, &key1))) return rc;
    if (giant > 1 && (rc = find_key(k, eb, &keyb))) return rc;
    Tmp tmp(s);
    u64 *states = nullptr, *rotated = nullptr, *ip = nullptr, *acc[2] = {nullptr, nullptr}, *scratch = nullptr, *ms = nullptr;
    const int64_t gal_items = std::max<int64_t>(batch, batch * results);
    const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(h->chunk, gal_items));
    CK(tmp.alloc(&states, ct_words * baby * batch));
    CK(tmp.alloc(&rotated, ct_words * baby * batch));
    CK(tmp.alloc(&ip, ct_words * results * giant * batch));
    CK(tmp.alloc(&acc[0], ct_words * results * batch));
    CK(tmp.alloc(&acc[1], ct_words * results * batch));
    CK(tmp.alloc(&scratch, galois_scratch_words(c, L) * (size_t)chunk));
    CK(tmp.alloc(&ms, ct_words * results * batch));
    cudaError_t e;
    // 1) v_j = theta^j(v): states[j][b]                                    (MatrixMultiplication.swift:180-189)
    CK(cudaMemcpyAsync(states, d_vec, ct_words * batch * sizeof(u64), cudaMemcpyDeviceToDevice, s));
    for (int j = 1; j < baby; ++j)
        for (int64_t done = 0; done < batch; done += chunk) {
            const int64_t items = std::min<int64_t>(chunk, batch - done);
            if ((e = apply_galois_chunk(c, scratch, key1, states + ct_words * ((j - 1) * batch + done), L, e1,
                                        states + ct_words * (j * batch + done), items, s)) != cudaSuccess)
                return cuda_fail(e, "rotateColumns");
        }
    // convertToEvalFormat (:190-193), then [j][b] -> [b][j] so that each vector's states are consecutive
    const NttRowMap map = c.map_q(L);
    if ((e = launch_ntt_forward(c, map, states, states, (int64_t)baby * batch * 2 * L, s)) != cudaSuccess) return cuda_fail(e, "ntt");
    if (batch == 1) {
        std::swap(states, rotated);
    } else {
        for (int j = 0; j < baby; ++j)
            CK(cudaMemcpy2DAsync(rotated + ct_words * j, ct_words * baby * sizeof(u64), states + ct_words * j * batch,
                                 ct_words * sizeof(u64), ct_words * sizeof(u64), (size_t)batch, cudaMemcpyDeviceToDevice, s));
    }
    // 2) w_k: one inner product per (result ciphertext, giant step)                         (:197-216)
    for (int64_t b = 0; b < batch; ++b)
        if ((e = launch_inner_product_plain(c, rotated + ct_words * baby * b, 2, L, baby, m->d_plain, m->d_present,
                                            ip + ct_words * results * giant * b, results * giant, s)) != cudaSuccess)
            return cuda_fail(e, "innerProduct(ciphertexts:plaintexts:)");
    if ((e = launch_ntt_inverse(c, map, ip, ip, batch * results * giant * 2 * L, kScalePlain, s)) != cudaSuccess)
        return cuda_fail(e, "ntt");
    // 3) rotateColumnsAndSum(by: -babyStep): Horner over the giant steps, all (vector, result) pairs at once (:218-226)
    const int64_t items = batch * results;
    int cur = 0;
    if ((e = launch_accumulate(c, L, acc[cur], ip + ct_words * (giant - 1), (int64_t)ct_words * giant, items, false, s)) != cudaSuccess)
        return cuda_fail(e, "sum");
    for (int g = giant - 2; g >= 0; --g) {
        for (int64_t done = 0; done < items; done += chunk) {
            const int64_t part = std::min<int64_t>(chunk, items - done);
            if ((e = apply_galois_chunk(c, scratch, keyb, acc[cur] + ct_words * done, L, eb, acc[cur ^ 1] + ct_words * done,
                                        part, s)) != cudaSuccess)
                return cuda_fail(e, "rotateColumns");
        }
        cur ^= 1;
        if ((e = launch_accumulate(c, L, acc[cur], ip + ct_words * g, (int64_t)ct_words * giant, items, true, s)) != cudaSuccess)
            return cuda_fail(e, "sum");
    }
    // Server.computeResponse: modSwitchDownToSingle (Server.swift:79-80)
    if (!to_single || L == 1) {
        CK(cudaMemcpyAsync(d_out, acc[cur], ct_words * items * sizeof(u64), cudaMemcpyDeviceToDevice, s));
        return HECUDA_OK;
    }
    const u64 *src = acc[cur];
    for (int l = L; l > 1; --l) {
        u64 *dst = l == 2 ? d_out : (src == ms ? acc[cur] : ms);
        if ((e = launch_mod_switch(c, src, l, dst, items * 2, s)) != cudaSuccess) return cuda_fail(e, "modSwitchDown");
        src = dst;
    }
    return HECUDA_OK;
}

// batched applyGalois over `items` contiguous ciphertexts, chunked by the scratch size
int32_t galois_batch(const Context &c, u64 *scratch, int64_t chunk, const u64 *key, unsigned element, const u64 *in, u64 *out,
                     int64_t items, cudaStream_t s) {
    const size_t ct_words = (size_t)2 * c.L * c.n;
    for (int64_t done = 0; done < items; done += chunk) {
        const int64_t part = std::min<int64_t>(chunk, items - done);
        cudaError_t e = apply_galois_chunk(c, scratch, key, in + ct_words * done, c.L, element, out + ct_words * done, part, s);
        if (e != cudaSuccess) return cuda_fail(e, "applyGalois");
    }
    return HECUDA_OK;
}

struct MatrixQuery {
    int32_t rows;                      // ciphertextMatrix.rowCount
    const int32_t *ciphertext_index;   // per row
    const u64 *host_masks;             // rows x N coefficient plaintexts
    const int32_t *rotate_count;       // per row
    int32_t column_step;
    const int32_t *pack_rotations;     // single rotations composing rotateColumnsMultiStep(by: matrix.rowCount)
    int32_t pack_rotation_count;
};

// PlaintextMatrix.mulTranspose(matrix:using:) (MatrixMultiplication.swift:236-298) with CiphertextMatrix.extractDenseRow
// (CiphertextMatrix.swift:245-352) batched over all query rows.
int32_t mul_transpose_matrix_device(const hecuda_context *h, const hecuda_evk *k, const hecuda_pnns_matrix *m,
                                    const u64 *d_cts, const MatrixQuery &q, bool to_single, u64 *d_out, int64_t out_capacity,
                                    int64_t *out_count, cudaStream_t s) {
    const Context &c = *h->ctx;
    const int L = c.L;
    const int64_t n = c.n, R = q.rows, results = m->result_count;
    const size_t ct_words = (size_t)2 * L * n;
    const int64_t per_simd_row = (n / 2) / m->row_count;
    const int64_t S = per_simd_row, G = S > 0 ? (R + S - 1) / S : 0;
    const int64_t outputs = S > 0 ? (G + 1) / 2 : R * results;
    *out_count = outputs;
    if (outputs > out_capacity) return fail(HECUDA_ERR_INVALID_ARGUMENT, "output buffer too small: needs " + std::to_string(outputs) + " ciphertexts");
    Tmp tmp(s);
    const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(h->chunk, R));
    u64 *x = nullptr, *y = nullptr, *z = nullptr, *scratch = nullptr, *inner = nullptr, *masks = nullptr, *mask_eval = nullptr;
    signed char *d_modes = nullptr;
    CK(tmp.alloc(&x, ct_words * R));
    CK(tmp.alloc(&y, ct_words * R));
    CK(tmp.alloc(&z, ct_words * R));
    CK(tmp.alloc(&scratch, galois_scratch_words(c, L) * (size_t)chunk));
    CK(tmp.alloc(&inner, ct_words * R * results));
    const NttRowMap map = c.map_q(L);
    const unsigned swap_element = (unsigned)(2 * n - 1);
    cudaError_t e;
    int32_t rc;
    // mode tables: replication steps (extractDenseRow) then packing positions
    int32_t max_rot = 0;
    for (int64_t r = 0; r < R && R > 1; ++r) max_rot = std::max(max_rot, q.rotate_count[r]);
    std::vector<signed char> modes;
    for (int32_t t = 1; t <= max_rot; ++t)
        for (int64_t r = 0; r < R; ++r) modes.push_back(t <= q.rotate_count[r] ? 2 : 0);
    const size_t pack_modes_offset = modes.size();
    const int64_t top = std::min<int64_t>(S, R) - 1;  // the longest group: positions above it hold nothing
    for (int64_t p = top; p >= 0 && S > 0; --p)
        for (int64_t g = 0; g < G; ++g) {
            const int64_t size = std::min<int64_t>(S, R - g * S);
            modes.push_back(p == size - 1 ? 1 : (p < size - 1 ? 2 : 0));
        }
    CK(cudaMallocAsync((void **)&d_modes, std::max<size_t>(modes.size(), 8), s));
    tmp.ptrs.push_back(d_modes);
    if (!modes.empty()) CK(cudaMemcpyAsync(d_modes, modes.data(), modes.size(), cudaMemcpyHostToDevice, s));
    if (R == 1) {  // extractDenseRow is the identity for a single row (:263-265)
        CK(cudaMemcpyAsync(y, d_cts, ct_words * sizeof(u64), cudaMemcpyDeviceToDevice, s));
    } else {
        const u64 *key_step = nullptr, *key_swap = nullptr;
        const unsigned step_element = rotating_columns(q.column_step, n);
        if (max_rot > 0 && (rc = find_key(k, step_element, &key_step))) return rc;
        if ((rc = find_key(k, swap_element, &key_swap))) return rc;
        CK(tmp.alloc(&masks, (size_t)n * R));
        CK(tmp.alloc(&mask_eval, (size_t)L * n * R));
        CK(cudaMemcpyAsync(masks, q.host_masks, (size_t)n * R * sizeof(u64), cudaMemcpyHostToDevice, s));
        for (int64_t r = 0; r < R; ++r)
            CK(cudaMemcpyAsync(x + ct_words * r, d_cts + ct_words * q.ciphertext_index[r], ct_words * sizeof(u64),
                               cudaMemcpyDeviceToDevice, s));
        // ciphertextEval *= plaintextMask (:322-324)
        if ((e = launch_ntt_forward(c, map, x, x, R * 2 * L, s)) != cudaSuccess) return cuda_fail(e, "ntt");
        if ((e = launch_plaintext_to_eval(c, masks, L, mask_eval, R, s)) != cudaSuccess) return cuda_fail(e, "plaintext_to_eval");
        for (int64_t r = 0; r < R; ++r)
            if ((e = launch_inner_product_plain(c, x + ct_words * r, 2, L, 1, mask_eval + (size_t)L * n * r, nullptr,
                                                y + ct_words * r, 1, s)) != cudaSuccess)
                return cuda_fail(e, "multiply by mask");
        if ((e = launch_ntt_inverse(c, map, y, y, R * 2 * L, kScalePlain, s)) != cudaSuccess) return cuda_fail(e, "ntt");
        // replicate across one SIMD row: rotate the copy, add where the row still needs copies (:331-336)
        const u64 *copy = y;
        u64 *ping[2] = {x, z};
        for (int32_t t = 1; t <= max_rot; ++t) {
            u64 *dst = ping[t & 1];
            if ((rc = galois_batch(c, scratch, chunk, key_step, step_element, copy, dst, R, s))) return rc;
            copy = dst;
            if ((e = launch_accumulate(c, L, y, copy, (int64_t)ct_words, R, true, s, d_modes + (size_t)(t - 1) * R)) != cudaSuccess)
                return cuda_fail(e, "sum");
        }
        // both SIMD rows: ciphertext += swapRows(ciphertext) (:342-345)
        if ((rc = galois_batch(c, scratch, chunk, key_swap, swap_element, y, x, R, s))) return rc;
        if ((e = launch_accumulate(c, L, y, x, (int64_t)ct_words, R, true, s)) != cudaSuccess) return cuda_fail(e, "sum");
    }
    if ((rc = mul_transpose_device(h, k, m, y, R, false, inner, s))) return rc;
    u64 *final_cts = inner;
    if (S > 0) {  // pack the result columns (:262-283); here results == 1
        u64 *acc[2] = {x, y};
        int cur = 0;
        CK(cudaMemsetAsync(acc[0], 0, ct_words * G * sizeof(u64), s));
        std::vector<std::pair<unsigned, const u64 *>> rotations;
        for (int32_t i = 0; i < q.pack_rotation_count; ++i) {
            const unsigned element = rotating_columns(q.pack_rotations[i], n);
            const u64 *key = nullptr;
            if (S > 1 && (rc = find_key(k, element, &key))) return rc;
            rotations.push_back({element, key});
        }
        const int64_t gchunk = std::max<int64_t>(1, std::min<int64_t>(chunk, G));
        size_t mode_row = pack_modes_offset;
        for (int64_t p = top; p >= 0; --p, mode_row += (size_t)G) {
            if (p < top)
                for (const auto &rot : rotations) {  // rotateColumnsMultiStep(by: dimensions.rowCount)
                    if ((rc = galois_batch(c, scratch, gchunk, rot.second, rot.first, acc[cur], acc[cur ^ 1], G, s))) return rc;
                    cur ^= 1;
                }
            if ((e = launch_accumulate(c, L, acc[cur], inner + ct_words * p, (int64_t)ct_words * S, G, true, s,
                                       d_modes + mode_row)) != cudaSuccess)
                return cuda_fail(e, "sum");
        }
        // swapRowsAndAdd(swapping: packedRows[1], addingTo: packedRows[0]) for every full pair (:277-281)
        const int64_t pairs = G / 2;
        u64 *packed = z;
        if ((e = launch_accumulate(c, L, packed, acc[cur], (int64_t)ct_words * 2, outputs, false, s)) != cudaSuccess)
            return cuda_fail(e, "copy");
        if (pairs > 0) {
            const u64 *key_swap = nullptr;
            if ((rc = find_key(k, swap_element, &key_swap))) return rc;
            u64 *odd = acc[cur ^ 1];
            if ((e = launch_accumulate(c, L, odd, acc[cur] + ct_words, (int64_t)ct_words * 2, pairs, false, s)) != cudaSuccess)
                return cuda_fail(e, "copy");
            if ((rc = galois_batch(c, scratch, std::max<int64_t>(1, std::min<int64_t>(chunk, pairs)), key_swap, swap_element, odd,
                                   inner, pairs, s)))
                return rc;
            if ((e = launch_accumulate(c, L, packed, inner, (int64_t)ct_words, pairs, true, s)) != cudaSuccess) return cuda_fail(e, "sum");
        }
        final_cts = packed;
    }
    if (!to_single || L == 1) {
        CK(cudaMemcpyAsync(d_out, final_cts, ct_words * outputs * sizeof(u64), cudaMemcpyDeviceToDevice, s));
    } else {
        const u64 *src = final_cts;
        u64 *spare[2] = {nullptr, nullptr};  // sized for the outputs (more than R ciphertexts when rowCount > N)
        CK(tmp.alloc(&spare[0], ct_words * outputs));
        CK(tmp.alloc(&spare[1], ct_words * outputs));
        int which = 0;
        for (int l = L; l > 1; --l) {
            u64 *dst = l == 2 ? d_out : spare[which];
            which ^= 1;
            if ((e = launch_mod_switch(c, src, l, dst, outputs * 2, s)) != cudaSuccess) return cuda_fail(e, "modSwitchDown");
            src = dst;
        }
    }
    CK(cudaStreamSynchronize(s));  // `modes` and the caller's descriptor arrays are host temporaries
    return HECUDA_OK;
}

int32_t check_args(const hecuda_context *h, const hecuda_evk *k, const hecuda_pnns_matrix *m, const void *vectors,
                   int64_t batch, const void *out) {
    int32_t rc = check_ctx(h);
    if (rc) return rc;
    if (!m || m->owner != h) return fail(HECUDA_ERR_INVALID_ARGUMENT, "wrongContext: plaintext matrix belongs to another context");
    if (!k) return fail(HECUDA_ERR_MISSING_KEY, "missingGaloisKey");
    if (k->owner != h) return fail(HECUDA_ERR_INVALID_ARGUMENT, "invalidContext: evaluation key belongs to another context");
    if (batch < 0 || (batch && (!vectors || !out))) return fail(HECUDA_ERR_INVALID_ARGUMENT, "invalidCiphertext: null buffer");
    return HECUDA_OK;
}

}  // namespace

extern "C" {

int32_t hecuda_pnns_matrix_create(const hecuda_context *h, const uint64_t *plaintexts, int32_t eval_format,
                                  int64_t row_count, int64_t column_count, int32_t baby_step, int32_t giant_step,
                                  hecuda_pnns_matrix **out) {
    int32_t rc = check_ctx(h);
    if (rc) return rc;
    if (!out || !plaintexts) return fail(HECUDA_ERR_INVALID_ARGUMENT, "null argument");
    *out = nullptr;
    const Context &c = *h->ctx;
    const int64_t n = c.n;
    if (row_count < 1 || column_count < 1 || column_count > n / 2)  // PnnsError.invalidMatrixDimensions (PlaintextMatrix.swift:429-431)
        return fail(HECUDA_ERR_INVALID_ARGUMENT, "invalidMatrixDimensions");
    int64_t dimension = 1;
    while (dimension < column_count) dimension <<= 1;
    if (baby_step < 1 || giant_step < 1 || baby_step < giant_step || (int64_t)baby_step * giant_step < dimension ||
        (int64_t)baby_step * (giant_step - 1) >= dimension || baby_step >= n / 2)
        return fail(HECUDA_ERR_INVALID_ARGUMENT, "wrongMatrixPacking: babyStep / giantStep do not cover the padded dimension");
    const int64_t results = (row_count + n - 1) / n;  // plaintextsPerColumnCount / resultCiphertextCount
    const int64_t count = dimension * results;       // PlaintextMatrix.plaintextCount, .diagonal (:269-273)
    const int64_t slots = results * giant_step * baby_step;
    const size_t row_words = (size_t)c.L * n, in_words = eval_format ? row_words : (size_t)n;
    hecuda_pnns_matrix *m = new (std::nothrow) hecuda_pnns_matrix();
    if (!m) return fail(HECUDA_ERR_CUDA, "out of host memory");
    m->owner = h;
    m->row_count = row_count;
    m->column_count = column_count;
    m->result_count = results;
    m->baby = baby_step;
    m->giant = giant_step;
    m->dimension = (int)dimension;
    std::vector<unsigned char> present((size_t)slots, 0);
    u64 *d_in = nullptr;
    cudaError_t e = cudaMalloc(&m->d_plain, row_words * slots * sizeof(u64));
    if (e == cudaSuccess) e = fill(m->d_plain, 0, row_words * slots * sizeof(u64));
    if (e == cudaSuccess) e = cudaMalloc(&m->d_present, (size_t)slots);
    if (e == cudaSuccess) e = cudaMalloc(&d_in, in_words * count * sizeof(u64));
    if (e == cudaSuccess) e = upload(d_in, plaintexts, in_words * count * sizeof(u64));
    // plaintext index resultCount * (j + babyStep * g) + r  ->  slot (r, g, j)      (MatrixMultiplication.swift:203-205)
    for (int64_t r = 0; e == cudaSuccess && r < results; ++r)
        for (int g = 0; e == cudaSuccess && g < giant_step; ++g) {
            const int64_t terms = std::min<int64_t>(baby_step, dimension - (int64_t)baby_step * g);
            const int64_t slot = (r * giant_step + g) * baby_step;
            const u64 *src = d_in + in_words * (results * ((int64_t)baby_step * g) + r);
            if (eval_format) {
                e = cudaMemcpy2D(m->d_plain + row_words * slot, row_words * sizeof(u64), src, row_words * results * sizeof(u64),
                                 row_words * sizeof(u64), (size_t)terms, cudaMemcpyDeviceToDevice);
            } else {  // Plaintext.convertToEvalFormat per row (:206-208); gather the strided rows first
                u64 *d_rows = nullptr;
                e = cudaMalloc(&d_rows, (size_t)n * terms * sizeof(u64));
                if (e == cudaSuccess)
                    e = cudaMemcpy2D(d_rows, n * sizeof(u64), src, n * results * sizeof(u64), n * sizeof(u64), (size_t)terms,
                                     cudaMemcpyDeviceToDevice);
                if (e == cudaSuccess) e = launch_plaintext_to_eval(c, d_rows, c.L, m->d_plain + row_words * slot, terms, nullptr);
                if (e == cudaSuccess) e = cudaStreamSynchronize(nullptr);
                cudaFree(d_rows);
            }
            for (int64_t j = 0; j < terms; ++j) present[(size_t)(slot + j)] = 1;
        }
    if (e == cudaSuccess) e = upload(m->d_present, present.data(), (size_t)slots);
    cudaFree(d_in);
    if (e != cudaSuccess) {
        hecuda_pnns_matrix_destroy(m);
        return cuda_fail(e, "pnns matrix upload");
    }
    *out = m;
    return HECUDA_OK;
}

int32_t hecuda_pnns_matrix_destroy(hecuda_pnns_matrix *m) {
    if (!m) return HECUDA_OK;
    if (m->d_plain) cudaFree(m->d_plain);
    if (m->d_present) cudaFree(m->d_present);
    delete m;
    return HECUDA_OK;
}

int32_t hecuda_pnns_matrix_result_count(const hecuda_pnns_matrix *m, int64_t *count) {
    if (!m || !count) return fail(HECUDA_ERR_INVALID_ARGUMENT, "null argument");
    *count = m->result_count;
    return HECUDA_OK;
}

int32_t hecuda_pnns_mul_transpose_vector_device(const hecuda_context *h, const hecuda_evk *k, const hecuda_pnns_matrix *m,
                                                const uint64_t *vectors, int64_t batch, int32_t mod_switch_to_single,
                                                uint64_t *out, void *stream) {
    int32_t rc = check_args(h, k, m, vectors, batch, out);
    if (rc || batch == 0) return rc;
    return mul_transpose_device(h, k, m, (const u64 *)vectors, batch, mod_switch_to_single != 0, (u64 *)out, (cudaStream_t)stream);
}

int32_t hecuda_pnns_mul_transpose_vector(const hecuda_context *h, const hecuda_evk *k, const hecuda_pnns_matrix *m,
                                         const uint64_t *vectors, int64_t batch, int32_t mod_switch_to_single,
                                         uint64_t *out) {
    int32_t rc = check_args(h, k, m, vectors, batch, out);
    if (rc || batch == 0) return rc;
    WsGuard g(h);
    if (!g.w) return fail(HECUDA_ERR_CUDA, "could not create a CUDA stream / workspace");
    const Context &c = *h->ctx;
    const size_t ct_words = (size_t)2 * c.L * c.n;
    const size_t out_words = (size_t)2 * (mod_switch_to_single ? 1 : c.L) * c.n * m->result_count * batch;
    cudaStream_t s = g.w->stream;
    Tmp tmp(s);
    u64 *d_in = nullptr, *d_out = nullptr;
    CK(tmp.alloc(&d_in, ct_words * batch));
    CK(tmp.alloc(&d_out, ct_words * m->result_count * batch));
    CK(cudaMemcpyAsync(d_in, vectors, ct_words * batch * sizeof(u64), cudaMemcpyHostToDevice, s));
    rc = mul_transpose_device(h, k, m, d_in, batch, mod_switch_to_single != 0, d_out, s);
    if (rc) {
        cudaStreamSynchronize(s);
        return rc;
    }
    CK(cudaMemcpyAsync(out, d_out, out_words * sizeof(u64), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return HECUDA_OK;
}

int32_t hecuda_pnns_mul_transpose_matrix(const hecuda_context *h, const hecuda_evk *k, const hecuda_pnns_matrix *m,
                                         const uint64_t *ciphertexts, int32_t ciphertext_count, int32_t query_row_count,
                                         const int32_t *row_ciphertext_index, const uint64_t *row_masks,
                                         const int32_t *row_rotate_count, int32_t column_step, const int32_t *pack_rotations,
                                         int32_t pack_rotation_count, int32_t mod_switch_to_single, uint64_t *out,
                                         int64_t out_capacity, int64_t *out_count) {
    int32_t rc = check_args(h, k, m, ciphertexts, ciphertext_count, out);
    if (rc) return rc;
    if (!out_count || ciphertext_count < 1 || query_row_count < 1 || pack_rotation_count < 0 ||
        (pack_rotation_count && !pack_rotations))
        return fail(HECUDA_ERR_INVALID_ARGUMENT, "invalidMatrixDimensions");
    const Context &c = *h->ctx;
    if (query_row_count > 1) {
        if (!row_ciphertext_index || !row_masks || !row_rotate_count) return fail(HECUDA_ERR_INVALID_ARGUMENT, "null row descriptors");
        if (column_step < 1 || column_step > c.n / 2) return fail(HECUDA_ERR_INVALID_ARGUMENT, "invalidMatrixDimensions");
        for (int32_t r = 0; r < query_row_count; ++r)
            if (row_ciphertext_index[r] < 0 || row_ciphertext_index[r] >= ciphertext_count || row_rotate_count[r] < 0)
                return fail(HECUDA_ERR_INVALID_ARGUMENT, "wrongCiphertextCount: row descriptor out of range");
    }
    WsGuard g(h);
    if (!g.w) return fail(HECUDA_ERR_CUDA, "could not create a CUDA stream / workspace");
    const size_t ct_words = (size_t)2 * c.L * c.n;
    cudaStream_t s = g.w->stream;
    Tmp tmp(s);
    u64 *d_in = nullptr, *d_out = nullptr;
    CK(tmp.alloc(&d_in, ct_words * ciphertext_count));
    CK(tmp.alloc(&d_out, ct_words * (size_t)std::max<int64_t>(out_capacity, 1)));
    CK(cudaMemcpyAsync(d_in, ciphertexts, ct_words * ciphertext_count * sizeof(u64), cudaMemcpyHostToDevice, s));
    const MatrixQuery q{query_row_count, row_ciphertext_index, (const u64 *)row_masks, row_rotate_count, column_step,
                        pack_rotations, pack_rotation_count};
    rc = mul_transpose_matrix_device(h, k, m, d_in, q, mod_switch_to_single != 0, d_out, out_capacity, out_count, s);
    if (rc) {
        cudaStreamSynchronize(s);
        return rc;
    }
    const size_t out_words = (size_t)2 * (mod_switch_to_single ? 1 : c.L) * c.n * (size_t)*out_count;
    CK(cudaMemcpyAsync(out, d_out, out_words * sizeof(u64), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return HECUDA_OK;
}

}  // extern "C"
