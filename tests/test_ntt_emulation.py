"""CPU-side verification of the register-tiled NTT kernels' index maps and lazy-reduction schedule.

tests/emu/ntt_emulate.cu instantiates the very same __host__ __device__ pass functions the sm_100a kernels call
(csrc/ntt_fast.cuh) and replays them thread by thread on the host; the result must equal the oracle's NTT.
Also checks that the chosen pass plans are shared-memory bank-conflict free in the TMA 128-byte swizzle."""
import os
import shutil
import subprocess
from collections import Counter

import numpy as np
import pytest

from oracle import oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU_SRC = os.path.join(ROOT, "tests", "emu", "ntt_emulate.cu")
EMU_BIN = os.path.join(ROOT, "tests", "emu", "ntt_emulate")
NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"

PLANS = {10: (3, 3, 4), 11: (4, 3, 4), 12: (4, 4, 4), 13: (3, 3, 3, 4), 14: (4, 3, 3, 4)}


@pytest.fixture(scope="module")
def emu():
    if not os.path.exists(NVCC):
        pytest.skip("nvcc not available")
    hdr = os.path.join(ROOT, "swift-homomorphic-encryption_b200", "csrc", "ntt_fast.cuh")
    if not os.path.exists(EMU_BIN) or os.path.getmtime(EMU_BIN) < max(os.path.getmtime(EMU_SRC), os.path.getmtime(hdr)):
        subprocess.check_call([NVCC, "-O1", "-std=c++17", "-Wno-deprecated-gpu-targets", "-o", EMU_BIN, EMU_SRC])
    return EMU_BIN


def run_emu(binary, logn, p, direction, t, data):
    text = "\n".join(str(int(v)) for v in data) + "\n"
    out = subprocess.run([binary, str(logn), str(p), direction, str(t)], input=text, capture_output=True, text=True,
                         check=True).stdout
    return np.array([int(v) for v in out.split()], dtype=np.uint64)


@pytest.mark.parametrize("logn", [10, 11, 12, 13, 14])
@pytest.mark.parametrize("bits", [20, 27, 30, 31, 55, 56, 57, 61, 62])
def test_emulated_kernels_match_oracle(emu, logn, bits):
    n = 1 << logn
    p = orc.generate_primes([bits], False, n)[0]
    x = orc.fill_uniform(logn * 100 + bits, [p], n, 1)[0]
    x[:3] = [0, p - 1, 1]
    fwd = run_emu(emu, logn, p, "fwd", 0, x)
    assert np.array_equal(fwd, orc.ntt_forward(n, [p], [x])[0])
    inv = run_emu(emu, logn, p, "inv", 0, fwd)
    assert np.array_equal(inv, x)
    # worst case for the lazy bounds: all residues p - 1
    worst = np.full(n, p - 1, dtype=np.uint64)
    assert np.array_equal(run_emu(emu, logn, p, "fwd", 0, worst), orc.ntt_forward(n, [p], [worst])[0])
    assert np.array_equal(run_emu(emu, logn, p, "inv", 0, worst), orc.ntt_inverse(n, [p], [worst])[0])
    # inverse with the BFV `* t` folded in (Bfv+Multiply.swift:40)
    t = 557057
    scaled = run_emu(emu, logn, p, "inv", t, fwd)
    assert [int(v) for v in scaled[:64]] == [int(v) * t % p for v in x[:64]]


def _passes(logn, inverse):
    plan = PLANS[logn][::-1] if inverse else PLANS[logn]
    out, acc = [], 0
    for c in plan:
        out.append((acc, c) if inverse else (logn - acc - c, c))
        acc += c
    return out


@pytest.mark.parametrize("logn", sorted(PLANS))
@pytest.mark.parametrize("inverse", [False, True])
def test_plans_cover_all_elements_and_are_conflict_free(logn, inverse):
    """Replays the index map of ntt_fast.cuh (registers = stage bits x low passenger bits) and the 128-byte TMA swizzle
    phys(e) = e ^ (((e >> 4) & 7) << 1): every pass touches each element exactly once, and every shared-memory access is
    bank-conflict free -- 128-bit accesses are served per quarter warp (8 lanes x 16 bytes must hit 8 distinct
    16-byte bank groups), 64-bit accesses per half warp (16 distinct 8-byte bank pairs)."""
    n, T = 1 << logn, (1 << logn) // 16
    assert sum(PLANS[logn]) == logn
    for LB, C in _passes(logn, inverse):
        E = 0 if LB == 0 else 4 - C
        assert LB == 0 or LB >= E
        vec = E >= 1 or LB == 0
        seen = set()
        for r in range(0, 16, 2 if vec else 1):
            lanes = 8 if vec else 16
            for w0 in range(0, T, lanes):
                slots = []
                for tau in range(w0, min(w0 + lanes, T)):
                    lo, hi = tau & ((1 << (LB - E)) - 1), tau >> (LB - E)
                    e = (hi << (LB + C)) | ((r >> E) << LB) | (lo << E) | (r & ((1 << E) - 1))
                    seen.add(e)
                    if vec:
                        assert e % 2 == 0
                        seen.add(e + 1)
                    phys = e ^ (((e >> 4) & 7) << 1)  # words (128-byte TMA swizzle)
                    slots.append((phys // 2) % 8 if vec else phys % 16)
                assert max(Counter(slots).values()) == 1, (logn, inverse, LB, C, r, w0)
        assert seen == set(range(n))


@pytest.mark.parametrize("logn", [10, 12, 13, 14])
def test_emulated_narrow_h_class(emu, logn):
    """Primes h 2^32 + 1 below 2^55 (the auxiliary primes of the multiply) take the NARROW-H butterfly."""
    n = 1 << logn
    p = (1 << 54) + 1
    while not orc.is_prime(p):
        p += 1 << 32
    assert p < 1 << 55 and p % (1 << 32) == 1
    x = orc.fill_uniform(logn, [p], n, 1)[0]
    x[:3] = [0, p - 1, 1]
    fwd = run_emu(emu, logn, p, "fwd", 0, x)
    assert np.array_equal(fwd, orc.ntt_forward(n, [p], [x])[0])
    assert np.array_equal(run_emu(emu, logn, p, "inv", 0, fwd), x)
    worst = np.full(n, p - 1, dtype=np.uint64)
    assert np.array_equal(run_emu(emu, logn, p, "fwd", 0, worst), orc.ntt_forward(n, [p], [worst])[0])
    assert np.array_equal(run_emu(emu, logn, p, "inv", 0, worst), orc.ntt_inverse(n, [p], [worst])[0])
