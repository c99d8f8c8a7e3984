"""The C++ host-side mirror of the reference interface (swift-homomorphic-encryption_b200/host/HeScheme.hpp):
compiles and links against libhecuda.so on CPU; on a GPU it replays the golden multiply -> relinearize ->
modSwitchDown case and the reference's error paths through the mirror."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "host_mirror_test.cpp")
BIN = os.path.join(ROOT, "tests", "cpp", "host_mirror_test")
LIBDIR = os.path.join(ROOT, "swift-homomorphic-encryption_b200")


def build():
    deps = [SRC, os.path.join(LIBDIR, "host", "HeScheme.hpp"), os.path.join(ROOT, "include", "hecuda.h")]
    if not os.path.exists(BIN) or os.path.getmtime(BIN) < max(os.path.getmtime(d) for d in deps):
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-o", BIN, SRC, "-L" + LIBDIR, "-lhecuda",
                               "-Wl,-rpath," + LIBDIR])
    return BIN


def test_host_mirror_compiles_and_links():
    assert os.path.exists(os.path.join(LIBDIR, "libhecuda.so")), "build libhecuda.so first"
    build()


@pytest.mark.gpu
def test_host_mirror_golden_case(tmp_path):
    z = np.load(os.path.join(ROOT, "tests", "golden", "mul_n64.npz"))
    path = tmp_path / "case.bin"
    with open(path, "wb") as f:
        def put(arr):
            a = np.ascontiguousarray(np.asarray(arr, dtype=np.uint64)).ravel()
            f.write(np.uint64(a.size).tobytes())
            f.write(a.tobytes())
        put([int(z["n"]), int(z["t"]), z["a"].shape[0]])
        for k in ("moduli", "a", "b", "relin_key", "product", "relinearized", "switched"):
            put(z[k])
    # second case: the Galois / coefficient-wise / inner-product surface, expected values from the oracle
    from oracle import oracle as orc
    n, t, moduli = int(z["n"]), int(z["t"]), [int(m) for m in z["moduli"]]
    o = orc.Context(n, moduli, t)
    L = o.L
    sk, _ = o.keygen(3, relin=False)
    element, step, terms = 3, -2, 4
    x = np.asarray(z["a"][0], dtype=np.uint64).reshape(2, L, n)
    y = np.asarray(z["b"][0], dtype=np.uint64).reshape(2, L, n)
    gk = o.galois_keygen(11, sk, element)
    rk = o.galois_keygen(12, sk, orc.galois_element_rotating_columns(step, n))
    q = np.array(moduli[:L], dtype=np.uint64)[None, :, None]
    single = x
    while single.shape[-2] > 1:
        single = o.mod_switch_down(single)[0]
    cts = np.stack([np.stack([orc.ntt_forward(n, moduli[:L], c[p]) for p in range(2)])
                    for c in orc.fill_uniform(9, moduli[:L], n, terms * 2 * L).reshape(terms, 2, L, n)])
    pts = orc.fill_uniform(10, moduli[:L], n, terms * L).reshape(terms, L, n)
    present = np.array([1, 0, 1, 1], dtype=np.uint8)
    pts[1] = 0
    path2 = tmp_path / "case2.bin"
    with open(path2, "wb") as f:
        def put(arr):
            a = np.ascontiguousarray(np.asarray(arr, dtype=np.uint64)).ravel()
            f.write(np.uint64(a.size).tobytes())
            f.write(a.tobytes())
        put(np.array([element, step, terms], dtype=np.int64).view(np.uint64))
        for arr in (x, y, gk, rk, (x + y) % q, (x + q - y) % q, (q - x) % q, o.apply_galois(x, element, gk)[0],
                    o.apply_galois(x, orc.galois_element_rotating_columns(step, n), rk)[0], single, cts, pts, present,
                    o.inner_product_plain(cts, pts[None], present[None])[0]):
            put(arr)
    out = subprocess.run([build(), str(path), str(path2)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "0 failure(s)" in out.stdout
