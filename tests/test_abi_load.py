"""CPU-side checks of the drop-in boundary: libhecuda.so loads, exports every symbol include/hecuda.h declares,
and refuses to work without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "hecuda.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hecuda_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    import hecuda

    assert declared_symbols() == sorted(hecuda.SYMBOLS)


def test_library_exports_every_declared_symbol():
    import hecuda

    assert os.path.exists(hecuda.LIB_PATH), "build libhecuda.so first: python -c 'import __graft_entry__ as g; g.build()'"
    lib = C.CDLL(hecuda.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} missing from libhecuda.so"
    assert hecuda.load_library().hecuda_version() >= 100


def test_header_cites_reference_for_each_op():
    text = open(HEADER).read()
    for needle in ("Bfv+Multiply.swift:18-21", "Bfv.swift:201-219", "Bfv.swift:163-171", "PolyRq+Ntt.swift:230",
                   "PolyRq+Ntt.swift:329-347", "Context.swift:94-143", "Sources/CUtil"):
        assert needle in text


def test_no_cpu_fallback_without_gpu():
    import hecuda

    if hecuda.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(hecuda.HeError) as ei:
        hecuda.Context(8192, [36028797018652673, 36028797017571329], 557057)
    assert ei.value.code == -4  # HECUDA_ERR_NO_DEVICE


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "swift-homomorphic-encryption_b200")
    for dirpath, _, files in os.walk(pkg):
        if "build" in dirpath.split(os.sep):
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "he_oracle" not in src and "from oracle" not in src and "import oracle" not in src, f


def _build_c_example(tmp_path, name="multiply_relinearize"):
    import shutil
    import subprocess
    gcc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else shutil.which("gcc")
    libdir = os.path.join(ROOT, "swift-homomorphic-encryption_b200")
    out = str(tmp_path / name)
    subprocess.check_call([gcc, "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-O1", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", name + ".c"), "-L" + libdir, "-lhecuda",
                           "-Wl,-rpath," + libdir, "-o", out])
    return out


def test_header_is_valid_c99_and_example_links(tmp_path):
    """include/hecuda.h is a C header (the reference's FFI is C: Sources/CUtil): a strict C99 client compiles and links."""
    import subprocess
    exe = _build_c_example(tmp_path)
    import torch
    if not torch.cuda.is_available():
        run = subprocess.run([exe], capture_output=True, text=True)
        assert run.returncode != 0 and "no CUDA device" in run.stderr   # fails loudly: no CPU fallback


@pytest.mark.gpu
def test_c_example_runs_on_gpu(tmp_path):
    import subprocess
    run = subprocess.run([_build_c_example(tmp_path)], capture_output=True, text=True)
    assert run.returncode == 0, run.stderr
    assert "16 products" in run.stdout


def test_key_broadcast_example_links(tmp_path):
    """examples/evk_broadcast.c (multi-GPU setup through the C ABI: NCCL communicator + evaluation-key broadcast) is
    strict C99 and links; without a GPU it fails loudly."""
    import subprocess
    exe = _build_c_example(tmp_path, "evk_broadcast")
    import torch
    if not torch.cuda.is_available():
        run = subprocess.run([exe, "0", "1", str(tmp_path / "id")], capture_output=True, text=True)
        assert run.returncode != 0


def _checksums(outputs):
    import re
    return [int(re.search(r"checksum (\d+)", o).group(1)) for o in outputs]


@pytest.mark.gpu
def test_key_broadcast_example_single_process(tmp_path):
    import subprocess
    run = subprocess.run([_build_c_example(tmp_path, "evk_broadcast"), "0", "1", str(tmp_path / "id")], capture_output=True, text=True)
    assert run.returncode == 0, run.stderr
    assert "ciphertexts [0, 8)" in run.stdout


@pytest.mark.gpu
def test_key_broadcast_over_nccl_two_processes(tmp_path):
    """Two processes, two GPUs, no torch: rank 0 creates the relinearization and Galois keys, hecuda_evk_broadcast moves
    them over NCCL, each rank relinearizes and rotates its half of the batch; the halves' checksums add up to the
    single-process checksum (mod 2^64)."""
    import subprocess
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    exe = _build_c_example(tmp_path, "evk_broadcast")
    single = subprocess.run([exe, "0", "1", str(tmp_path / "id1")], capture_output=True, text=True)
    assert single.returncode == 0, single.stderr
    idfile = str(tmp_path / "id2")
    procs = [subprocess.Popen([exe, str(r), "2", idfile], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True) for r in range(2)]
    outs = [p.communicate(timeout=300) for p in procs]
    assert all(p.returncode == 0 for p in procs), [o[1] for o in outs]
    halves = _checksums([o[0] for o in outs])
    assert sum(halves) % (1 << 64) == _checksums([single.stdout])[0]
