// host_mirror_test.cpp -- exercises the C++ mirror of the reference interface (host/HeScheme.hpp) end to end on a GPU:
// reads a golden case (written by the Python test from tests/golden/mul_n64.npz), runs mulAssign / relinearize /
// modSwitchDown and the error paths, and compares with the expected residues.
//   usage: host_mirror_test <case.bin>
#include <cstdio>
#include <cstring>
#include <fstream>

#include "../../swift-homomorphic-encryption_b200/host/HeScheme.hpp"

static std::vector<uint64_t> read_vec(std::ifstream &f) {
    uint64_t n = 0;
    f.read((char *)&n, 8);
    std::vector<uint64_t> v(n);
    f.read((char *)v.data(), 8 * n);
    return v;
}

int main(int argc, char **argv) {
    if (argc < 2) return 2;
    std::ifstream f(argv[1], std::ios::binary);
    if (!f) return 3;
    auto hdr = read_vec(f);  // n, t, batch
    auto moduli = read_vec(f), a = read_vec(f), b = read_vec(f), key = read_vec(f);
    auto product = read_vec(f), relin = read_vec(f), switched = read_vec(f);
    const int64_t n = (int64_t)hdr[0];
    const int batch = (int)hdr[2];
    auto ctx = std::make_shared<const he::Context>(n, moduli, hdr[1]);
    const int L = ctx->ciphertextModuliCount();
    const size_t pw = (size_t)L * n;
    he::EvaluationKey evk(ctx, key);
    int failures = 0;
    std::vector<he::Ciphertext> lhs, rhs;
    for (int i = 0; i < batch; ++i) {
        he::Ciphertext x(ctx, 2, L), y(ctx, 2, L);
        std::copy(a.begin() + i * 2 * pw, a.begin() + (i + 1) * 2 * pw, x.data.begin());
        std::copy(b.begin() + i * 2 * pw, b.begin() + (i + 1) * 2 * pw, y.data.begin());
        lhs.push_back(x);
        rhs.push_back(y);
    }
    // fused paths: multiply+relinearize(+modSwitchDown) over the batch, relinearize+modSwitchDown on a product
    {
        auto l1 = lhs, l2 = lhs;
        he::Bfv::mulRelinearizeAssign(l1, rhs, evk, false);
        he::Bfv::mulRelinearizeAssign(l2, rhs, evk, true);
        for (int i = 0; i < batch; ++i) {
            failures += l1[i].polyCount != 2 || l1[i].moduliCount != L || l1[i].data.size() != 2 * pw ||
                        std::memcmp(l1[i].data.data(), relin.data() + i * 2 * pw, 2 * pw * 8) != 0;
            failures += l2[i].moduliCount != L - 1 || l2[i].data.size() != 2 * (pw - n) ||
                        std::memcmp(l2[i].data.data(), switched.data() + i * 2 * (pw - n), 2 * (pw - n) * 8) != 0;
        }
        he::Ciphertext prod(ctx, 3, L);
        std::copy(product.begin(), product.begin() + 3 * pw, prod.data.begin());
        he::Bfv::relinearizeModSwitchDown(prod, evk);
        failures += prod.polyCount != 2 || prod.moduliCount != L - 1 ||
                    std::memcmp(prod.data.data(), switched.data(), 2 * (pw - n) * 8) != 0;
        try { he::Ciphertext two(ctx, 2, L); he::Bfv::relinearizeModSwitchDown(two, evk); ++failures; } catch (const he::HeError &e) { failures += e.kind != he::HeError::invalidCiphertext; }
    }
    // single-ciphertext path
    he::Ciphertext one = lhs[0];
    he::Bfv::mulAssign(one, rhs[0]);
    failures += std::memcmp(one.data.data(), product.data(), 3 * pw * 8) != 0;
    // batched path
    he::Bfv::mulAssign(lhs, rhs);
    for (int i = 0; i < batch; ++i) failures += std::memcmp(lhs[i].data.data(), product.data() + i * 3 * pw, 3 * pw * 8) != 0;
    for (int i = 0; i < batch; ++i) {
        he::Bfv::relinearize(lhs[i], evk);
        failures += lhs[i].polyCount != 2 || std::memcmp(lhs[i].data.data(), relin.data() + i * 2 * pw, 2 * pw * 8) != 0;
        he::Bfv::modSwitchDown(lhs[i]);
        failures += lhs[i].moduliCount != L - 1 ||
                    std::memcmp(lhs[i].data.data(), switched.data() + i * 2 * (pw - n), 2 * (pw - n) * 8) != 0;
    }
    // error behaviour (Bfv+Multiply.swift:66-76, Bfv.swift:205-207)
    try { he::Bfv::mulAssign(lhs[0], rhs[0]); ++failures; } catch (const he::HeError &e) { failures += e.kind != he::HeError::incompatibleCiphertexts; }
    try { he::Ciphertext three(ctx, 3, L); he::Bfv::mulAssign(three, rhs[0]); ++failures; } catch (const he::HeError &e) { failures += e.kind != he::HeError::invalidCiphertext; }
    try { he::Ciphertext two(ctx, 2, L); he::Bfv::relinearize(two, evk); ++failures; } catch (const he::HeError &e) { failures += e.kind != he::HeError::invalidCiphertext; }
    // NTT round trip
    he::PolyRq poly(ctx, L);
    std::copy(a.begin(), a.begin() + pw, poly.data.begin());
    he::Bfv::forwardNtt(poly);
    he::Bfv::inverseNtt(poly);
    failures += std::memcmp(poly.data.data(), a.data(), pw * 8) != 0;
    // optional second case: Galois / coefficient-wise / inner-product surface
    if (argc >= 3) {
        std::ifstream f2(argv[2], std::ios::binary);
        if (!f2) return 4;
        auto h2 = read_vec(f2);  // element, rotation step, terms
        auto ct = read_vec(f2), other = read_vec(f2), gkey = read_vec(f2), rkey = read_vec(f2);
        auto sum = read_vec(f2), diff = read_vec(f2), neg = read_vec(f2), galois = read_vec(f2), rotated = read_vec(f2);
        auto single = read_vec(f2), ipCts = read_vec(f2), ipPts = read_vec(f2), ipPresent = read_vec(f2), ipOut = read_vec(f2);
        he::EvaluationKey gal(ctx);
        gal.setGaloisKey((uint32_t)h2[0], gkey);
        gal.setGaloisKey(he::Bfv::rotatingColumnsElement((int)(int64_t)h2[1], n), rkey);
        auto load = [&](const std::vector<uint64_t> &v, int polys, int rows) {
            he::Ciphertext c(ctx, polys, rows);
            std::copy(v.begin(), v.begin() + c.data.size(), c.data.begin());
            return c;
        };
        auto same = [&](const he::Ciphertext &c, const std::vector<uint64_t> &v) {
            return c.data.size() == v.size() && std::memcmp(c.data.data(), v.data(), v.size() * 8) == 0;
        };
        he::Ciphertext x = load(ct, 2, L), y = load(other, 2, L);
        he::Ciphertext t1 = x; he::Bfv::addAssign(t1, y); failures += !same(t1, sum);
        he::Ciphertext t2 = x; he::Bfv::subAssign(t2, y); failures += !same(t2, diff);
        he::Ciphertext t3 = x; he::Bfv::negAssign(t3); failures += !same(t3, neg);
        he::Ciphertext t4 = x; he::Bfv::applyGalois(t4, (uint32_t)h2[0], gal); failures += !same(t4, galois);
        he::Ciphertext t5 = x; he::Bfv::rotateColumns(t5, (int)(int64_t)h2[1], gal); failures += !same(t5, rotated);
        he::Ciphertext t6 = x; he::Bfv::modSwitchDownToSingle(t6); failures += t6.moduliCount != 1 || !same(t6, single);
        const int terms = (int)h2[2];
        std::vector<he::Ciphertext> cts;
        std::vector<he::PolyRq> pts;
        std::vector<const he::PolyRq *> ptrs;
        for (int k = 0; k < terms; ++k) {
            he::Ciphertext c(ctx, 2, L);
            std::copy(ipCts.begin() + k * 2 * pw, ipCts.begin() + (k + 1) * 2 * pw, c.data.begin());
            cts.push_back(c);
            he::PolyRq p(ctx, L);
            std::copy(ipPts.begin() + k * pw, ipPts.begin() + (k + 1) * pw, p.data.begin());
            pts.push_back(p);
        }
        for (int k = 0; k < terms; ++k) ptrs.push_back(ipPresent[k] ? &pts[k] : nullptr);
        failures += !same(he::Bfv::innerProduct(cts, ptrs), ipOut);
        try { he::Bfv::applyGalois(x, 5, gal); ++failures; } catch (const he::HeError &e) { failures += e.kind != he::HeError::missingGaloisKey; }
        try { he::Ciphertext three(ctx, 3, L); he::Bfv::addAssign(three, x); ++failures; } catch (const he::HeError &e) { failures += e.kind != he::HeError::incompatibleCiphertexts; }
    }
    std::printf("host mirror: %d failure(s)\n", failures);
    return failures ? 1 : 0;
}
