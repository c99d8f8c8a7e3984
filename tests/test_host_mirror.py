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
    out = subprocess.run([build(), str(path)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "0 failure(s)" in out.stdout
