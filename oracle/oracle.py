"""ctypes front-end of the CPU oracle (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import this.
The product (swift-homomorphic-encryption_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libhe_oracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "he_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


_lib = None
u64p = C.POINTER(C.c_uint64)


def _p(a: np.ndarray):
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(u64p)


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.orc_pow_mod.restype = C.c_uint64
        L.orc_pow_mod.argtypes = [C.c_uint64] * 3
        L.orc_inverse_mod.restype = C.c_uint64
        L.orc_inverse_mod.argtypes = [C.c_uint64] * 2
        L.orc_is_prime.restype = C.c_int
        L.orc_is_prime.argtypes = [C.c_uint64]
        L.orc_generate_primes.restype = C.c_int
        L.orc_generate_primes.argtypes = [C.POINTER(C.c_int32), C.c_int32, C.c_int32, C.c_int64, u64p]
        L.orc_min_primitive_root.restype = C.c_uint64
        L.orc_min_primitive_root.argtypes = [C.c_int64, C.c_uint64]
        L.orc_reverse_bits.restype = C.c_uint32
        L.orc_reverse_bits.argtypes = [C.c_uint32, C.c_int32]
        for name in ("orc_barrett_reduce_single",):
            getattr(L, name).restype = C.c_uint64
            getattr(L, name).argtypes = [C.c_uint64] * 2
        for name in ("orc_barrett_reduce_double", "orc_barrett_reduce_product", "orc_shoup_mul", "orc_shoup_mul_lazy"):
            getattr(L, name).restype = C.c_uint64
            getattr(L, name).argtypes = [C.c_uint64] * 3
        L.orc_ntt_forward.argtypes = [C.c_int64, u64p, C.c_int32, u64p, C.c_int64]
        L.orc_ntt_inverse.argtypes = [C.c_int64, u64p, C.c_int32, u64p, C.c_int64]
        L.orc_ntt_forward_threads.argtypes = [C.c_int64, u64p, C.c_int32, u64p, C.c_int64, C.c_int32]
        L.orc_ntt_tables.argtypes = [C.c_int64, C.c_uint64, u64p, u64p, u64p, u64p]
        L.orc_divide_round_qlast.argtypes = [C.c_int64, u64p, C.c_int32, u64p]
        for name in ("orc_poly_add", "orc_poly_sub", "orc_poly_mul"):
            getattr(L, name).restype = None
            getattr(L, name).argtypes = [C.c_int64, u64p, C.c_int32, u64p, u64p]
        L.orc_rnstool_create.restype = C.c_void_p
        L.orc_rnstool_create.argtypes = [C.c_int64, u64p, C.c_int32, C.c_uint64]
        L.orc_rnstool_create_w.restype = C.c_void_p
        L.orc_rnstool_create_w.argtypes = [C.c_int64, u64p, C.c_int32, C.c_uint64, C.c_int32]
        L.orc_rnstool_destroy.argtypes = [C.c_void_p]
        L.orc_rnstool_bsk.argtypes = [C.c_void_p, u64p]
        L.orc_rnstool_small_montgomery_reduce.argtypes = [C.c_void_p, u64p]
        for name in ("orc_rnstool_lift", "orc_rnstool_approximate_floor", "orc_rnstool_bsk_to_q", "orc_rnstool_floor",
                     "orc_rnstool_convert_bsk_mtilde"):
            getattr(L, name).restype = None
            getattr(L, name).argtypes = [C.c_void_p, u64p, u64p]
        L.orc_rnstool_scale_and_round.restype = None
        L.orc_rnstool_scale_and_round.argtypes = [C.c_void_p, u64p, C.c_uint64, u64p]
        L.orc_convert_approximate.restype = None
        L.orc_convert_approximate.argtypes = [C.c_int64, u64p, C.c_int32, u64p, C.c_int32, u64p, u64p]
        L.orc_context_create.restype = C.c_void_p
        L.orc_context_create.argtypes = [C.c_int64, u64p, C.c_int32, C.c_uint64]
        L.orc_context_create_w.restype = C.c_void_p
        L.orc_context_create_w.argtypes = [C.c_int64, u64p, C.c_int32, C.c_uint64, C.c_int32]
        L.orc_context_destroy.argtypes = [C.c_void_p]
        L.orc_context_L.argtypes = [C.c_void_p]
        L.orc_context_bsk.argtypes = [C.c_void_p, u64p]
        L.orc_bfv_mul.argtypes = [C.c_void_p, u64p, u64p, u64p, C.c_int64, C.c_int32]
        L.orc_bfv_lift_ntt.argtypes = [C.c_void_p, u64p, u64p, C.c_int64]
        L.orc_keygen.argtypes = [C.c_void_p, C.c_uint64, u64p, u64p]
        L.orc_gen_keyswitch_key.argtypes = [C.c_void_p, C.c_uint64, u64p, u64p, u64p]
        L.orc_keyswitch_update.argtypes = [C.c_void_p, u64p, C.c_int32, u64p, u64p]
        L.orc_bfv_relinearize.argtypes = [C.c_void_p, u64p, C.c_int32, u64p, u64p, C.c_int64, C.c_int32]
        L.orc_bfv_mod_switch_down.argtypes = [C.c_void_p, u64p, C.c_int32, C.c_int32, u64p, C.c_int64, C.c_int32]
        L.orc_encrypt.argtypes = [C.c_void_p, C.c_uint64, u64p, u64p, u64p]
        L.orc_decrypt.argtypes = [C.c_void_p, u64p, u64p, C.c_int32, C.c_int32, u64p]
        L.orc_galois_coeff.restype = None
        L.orc_galois_coeff.argtypes = [C.c_int64, u64p, C.c_int32, C.c_int64, u64p, u64p]
        L.orc_galois_eval.restype = None
        L.orc_galois_eval.argtypes = [C.c_int64, C.c_int32, C.c_int64, u64p, u64p]
        L.orc_galois_element_rotating_columns.restype = C.c_int64
        L.orc_galois_element_rotating_columns.argtypes = [C.c_int64, C.c_int64]
        L.orc_galois_element_swapping_rows.restype = C.c_int64
        L.orc_galois_element_swapping_rows.argtypes = [C.c_int64]
        L.orc_gen_galois_key.argtypes = [C.c_void_p, C.c_uint64, u64p, C.c_int64, u64p]
        L.orc_bfv_apply_galois.argtypes = [C.c_void_p, u64p, C.c_int32, C.c_int64, u64p, u64p, C.c_int64, C.c_int32]
        L.orc_plaintext_to_eval.argtypes = [C.c_void_p, u64p, C.c_int32, u64p]
        L.orc_inner_product_plain.argtypes = [C.c_void_p, u64p, C.c_int32, C.c_int32, C.c_int64, u64p,
                                              C.POINTER(C.c_uint8), u64p, C.c_int64, C.c_int32]
        L.orc_bfv_inner_product.argtypes = [C.c_void_p, u64p, u64p, u64p, C.c_int64, C.c_int64, C.c_int32]
        L.orc_multiply_power_of_x.restype = None
        L.orc_multiply_power_of_x.argtypes = [C.c_int64, u64p, C.c_int32, C.c_int64, u64p, u64p]
        L.orc_fill_uniform.restype = None
        L.orc_fill_uniform.argtypes = [C.c_uint64, u64p, C.c_int32, C.c_int64, u64p, C.c_int64]
        L.orc_num_threads.restype = C.c_int
        _lib = L
    return _lib


def _arr(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.uint64))


# ---------------------------------------------------------------- scalar helpers
def generate_primes(bit_counts, prefer_small: bool, ntt_degree: int = 1):
    bc = (C.c_int32 * len(bit_counts))(*bit_counts)
    out = np.zeros(len(bit_counts), dtype=np.uint64)
    n = lib().orc_generate_primes(bc, len(bit_counts), int(prefer_small), ntt_degree, _p(out))
    if n != len(bit_counts):
        raise ValueError("notEnoughPrimes")
    return [int(v) for v in out]


def min_primitive_root(degree: int, p: int) -> int:
    return int(lib().orc_min_primitive_root(degree, p))


def is_prime(p: int) -> bool:
    return bool(lib().orc_is_prime(p))


def inverse_mod(a: int, p: int) -> int:
    return int(lib().orc_inverse_mod(a, p))


def reverse_bits(x: int, bits: int) -> int:
    return int(lib().orc_reverse_bits(x, bits))


# ---------------------------------------------------------------- ring ops
def ntt_forward(n: int, moduli, data):
    d = _arr(data).copy().reshape(-1, n)
    m = _arr(moduli)
    if lib().orc_ntt_forward(n, _p(m), len(m), _p(d), d.shape[0]) != 0:
        raise ValueError("invalidNttModulus")
    return d


def ntt_forward_inplace(n: int, moduli, data, threads: int = 1):
    """In place over the rows of a C-contiguous uint64 array, on `threads` OpenMP threads (no copy: timing use)."""
    m = _arr(moduli)
    assert data.dtype == np.uint64 and data.flags.c_contiguous
    if lib().orc_ntt_forward_threads(n, _p(m), len(m), _p(data), data.size // n, threads) != 0:
        raise ValueError("invalidNttModulus")
    return data


def ntt_inverse(n: int, moduli, data):
    d = _arr(data).copy().reshape(-1, n)
    m = _arr(moduli)
    if lib().orc_ntt_inverse(n, _p(m), len(m), _p(d), d.shape[0]) != 0:
        raise ValueError("invalidNttModulus")
    return d


def ntt_tables(n: int, p: int):
    roots = np.zeros(n, dtype=np.uint64)
    inv = np.zeros(n, dtype=np.uint64)
    a = np.zeros(1, dtype=np.uint64)
    b = np.zeros(1, dtype=np.uint64)
    if lib().orc_ntt_tables(n, p, _p(roots), _p(inv), _p(a), _p(b)) != 0:
        raise ValueError("invalidNttModulus")
    return roots, inv, int(a[0]), int(b[0])


def divide_round_qlast(n: int, moduli, data):
    d = _arr(data).copy().reshape(len(moduli), n)
    m = _arr(moduli)
    rc = lib().orc_divide_round_qlast(n, _p(m), len(m), _p(d))
    if rc != 0:
        raise ValueError(f"divide_round_qlast rc={rc}")
    return d[: len(moduli) - 1].copy()


def poly_op(op: str, n: int, moduli, lhs, rhs):
    a = _arr(lhs).copy().reshape(len(moduli), n)
    b = _arr(rhs).reshape(len(moduli), n)
    m = _arr(moduli)
    getattr(lib(), f"orc_poly_{op}")(n, _p(m), len(m), _p(a), _p(b))
    return a


def convert_approximate(n, q, tmod, data):
    q = _arr(q)
    tm = _arr(tmod)
    d = _arr(data).reshape(len(q), n)
    out = np.zeros((len(tm), n), dtype=np.uint64)
    lib().orc_convert_approximate(n, _p(q), len(q), _p(tm), len(tm), _p(d), _p(out))
    return out


def galois_coeff(n: int, moduli, element: int, data):
    d = _arr(data).reshape(len(moduli), n)
    out = np.zeros_like(d)
    m = _arr(moduli)
    lib().orc_galois_coeff(n, _p(m), len(m), element, _p(d), _p(out))
    return out


def galois_eval(n: int, nmod: int, element: int, data):
    d = _arr(data).reshape(nmod, n)
    out = np.zeros_like(d)
    lib().orc_galois_eval(n, nmod, element, _p(d), _p(out))
    return out


def galois_element_rotating_columns(step: int, degree: int) -> int:
    e = int(lib().orc_galois_element_rotating_columns(step, degree))
    if e == 0:
        raise ValueError("invalidRotationStep")
    return e


def galois_element_swapping_rows(degree: int) -> int:
    return int(lib().orc_galois_element_swapping_rows(degree))


def multiply_power_of_x(n: int, moduli, power: int, data):
    d = _arr(data).reshape(len(moduli), n)
    out = np.zeros_like(d)
    m = _arr(moduli)
    lib().orc_multiply_power_of_x(n, _p(m), len(m), power, _p(d), _p(out))
    return out


def fill_uniform(seed: int, moduli, n: int, rows: int):
    m = _arr(moduli)
    out = np.zeros((rows, n), dtype=np.uint64)
    lib().orc_fill_uniform(seed, _p(m), len(m), n, _p(out), rows)
    return out


class RnsTool:
    """_RnsTool at the top level for base q (RnsTool.swift:18-121)."""

    def __init__(self, n: int, q, t: int, word_bits: int = 64):
        """word_bits = 32: the reference's Bfv<UInt32> constants (m~ = 2^16, gamma = 2^30 - 20405, 29-bit Bsk)."""
        self.n, self.q, self.t, self.word_bits = n, [int(v) for v in q], t, word_bits
        qa = _arr(q)
        self.h = lib().orc_rnstool_create_w(n, _p(qa), len(qa), t, word_bits)
        if not self.h:
            raise ValueError("rnstool_create failed")
        out = np.zeros(len(q) + 1, dtype=np.uint64)
        lib().orc_rnstool_bsk(self.h, _p(out))
        self.bsk = [int(v) for v in out]

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_rnstool_destroy(self.h)
            self.h = None

    @property
    def nq(self):
        return len(self.q)

    def small_montgomery_reduce(self, data):
        d = _arr(data).copy().reshape(self.nq + 2, self.n)
        lib().orc_rnstool_small_montgomery_reduce(self.h, _p(d))
        return d[: self.nq + 1].copy()

    def _call(self, name, data, rows_in, rows_out):
        d = _arr(data).reshape(rows_in, self.n)
        out = np.zeros((rows_out, self.n), dtype=np.uint64)
        getattr(lib(), name)(self.h, _p(d), _p(out))
        return out

    def lift(self, data):
        return self._call("orc_rnstool_lift", data, self.nq, 2 * self.nq + 1)

    def convert_bsk_mtilde(self, data):
        return self._call("orc_rnstool_convert_bsk_mtilde", data, self.nq, self.nq + 2)

    def approximate_floor(self, data):
        return self._call("orc_rnstool_approximate_floor", data, 2 * self.nq + 1, self.nq + 1)

    def bsk_to_q(self, data):
        return self._call("orc_rnstool_bsk_to_q", data, self.nq + 1, self.nq)

    def floor(self, data):
        return self._call("orc_rnstool_floor", data, 2 * self.nq + 1, self.nq)

    def scale_and_round(self, data, scaling_factor=1):
        d = _arr(data).reshape(self.nq, self.n)
        out = np.zeros(self.n, dtype=np.uint64)
        lib().orc_rnstool_scale_and_round(self.h, _p(d), scaling_factor, _p(out))
        return out


class Context:
    """Context<Bfv<UInt64>> (Context.swift:94-143): coeff_moduli = [q_0..q_{L-1}, q_ks]."""

    def __init__(self, n: int, coeff_moduli, t: int, word_bits: int = 64):
        """word_bits = 32: Context<Bfv<UInt32>> -- the same residues with Bfv<UInt32>'s m~ / gamma / Bsk."""
        self.n, self.moduli, self.t, self.word_bits = n, [int(v) for v in coeff_moduli], int(t), word_bits
        m = _arr(coeff_moduli)
        self.h = lib().orc_context_create_w(n, _p(m), len(m), t, word_bits)
        if not self.h:
            raise ValueError("orc_context_create failed")
        self.L = len(self.moduli) - 1
        out = np.zeros(self.L + 1, dtype=np.uint64)
        lib().orc_context_bsk(self.h, _p(out))
        self.bsk = [int(v) for v in out]

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_context_destroy(self.h)
            self.h = None

    @property
    def q(self):
        return self.moduli[: self.L]

    def mul(self, a, b, threads: int = 0):
        a = _arr(a).reshape(-1, 2, self.L, self.n)
        b = _arr(b).reshape(-1, 2, self.L, self.n)
        out = np.zeros((a.shape[0], 3, self.L, self.n), dtype=np.uint64)
        rc = lib().orc_bfv_mul(self.h, _p(a), _p(b), _p(out), a.shape[0], threads)
        assert rc == 0
        return out

    def lift_ntt(self, polys):
        p = _arr(polys).reshape(-1, self.L, self.n)
        out = np.zeros((p.shape[0], 2 * self.L + 1, self.n), dtype=np.uint64)
        lib().orc_bfv_lift_ntt(self.h, _p(p), _p(out), p.shape[0])
        return out

    def keygen(self, seed: int, relin: bool = True):
        K = self.L + 1
        sk = np.zeros((K, self.n), dtype=np.uint64)
        rk = np.zeros((self.L, 2, K, self.n), dtype=np.uint64) if relin else None
        rc = lib().orc_keygen(self.h, seed, _p(sk), _p(rk) if relin else None)
        assert rc == 0
        return sk, rk

    def keyswitch_update(self, target, ksk):
        tgt = _arr(target)
        l = tgt.shape[0]
        out = np.zeros((2, l, self.n), dtype=np.uint64)
        rc = lib().orc_keyswitch_update(self.h, _p(tgt), l, _p(_arr(ksk)), _p(out))
        assert rc == 0
        return out

    def galois_keygen(self, seed: int, sk, element: int):
        K = self.L + 1
        gk = np.zeros((self.L, 2, K, self.n), dtype=np.uint64)
        rc = lib().orc_gen_galois_key(self.h, seed, _p(_arr(sk)), element, _p(gk))
        assert rc == 0
        return gk

    def apply_galois(self, ct, element: int, galois_key, threads: int = 0):
        c = _arr(ct)
        l = c.shape[-2]
        c = c.reshape(-1, 2, l, self.n)
        out = np.zeros_like(c)
        rc = lib().orc_bfv_apply_galois(self.h, _p(c), l, element, _p(_arr(galois_key)), _p(out), c.shape[0], threads)
        assert rc == 0
        return out

    def plaintext_to_eval(self, plain, l: int = 0):
        l = l or self.L
        out = np.zeros((l, self.n), dtype=np.uint64)
        rc = lib().orc_plaintext_to_eval(self.h, _p(_arr(plain)), l, _p(out))
        assert rc == 0
        return out

    def inner_product_plain(self, cts, pts, present=None, threads: int = 0):
        c = _arr(cts)  # (terms, npoly, l, n)
        terms, npoly, l = c.shape[0], c.shape[1], c.shape[2]
        p = _arr(pts).reshape(-1, terms, l, self.n)
        out = np.zeros((p.shape[0], npoly, l, self.n), dtype=np.uint64)
        pres = None
        if present is not None:
            pa = np.ascontiguousarray(np.asarray(present, dtype=np.uint8)).reshape(p.shape[0], terms)
            pres = pa.ctypes.data_as(C.POINTER(C.c_uint8))
        rc = lib().orc_inner_product_plain(self.h, _p(c), npoly, l, terms, _p(p), pres, _p(out), p.shape[0], threads)
        assert rc == 0
        return out

    def inner_product(self, lhs, rhs, threads: int = 0):
        a, b = _arr(lhs), _arr(rhs)  # (groups, pairs, 2, L, n)
        out = np.zeros((a.shape[0], 3, self.L, self.n), dtype=np.uint64)
        rc = lib().orc_bfv_inner_product(self.h, _p(a), _p(b), _p(out), a.shape[1], a.shape[0], threads)
        assert rc == 0
        return out

    def relinearize(self, ct3, relin_key, threads: int = 0):
        c = _arr(ct3)
        l = c.shape[-2]
        c = c.reshape(-1, 3, l, self.n)
        out = np.zeros((c.shape[0], 2, l, self.n), dtype=np.uint64)
        rc = lib().orc_bfv_relinearize(self.h, _p(c), l, _p(_arr(relin_key)), _p(out), c.shape[0], threads)
        assert rc == 0
        return out

    def mod_switch_down(self, ct, threads: int = 0):
        c = _arr(ct)
        npoly, l = c.shape[-3], c.shape[-2]
        c = c.reshape(-1, npoly, l, self.n)
        out = np.zeros((c.shape[0], npoly, l - 1, self.n), dtype=np.uint64)
        rc = lib().orc_bfv_mod_switch_down(self.h, _p(c), npoly, l, _p(out), c.shape[0], threads)
        assert rc == 0
        return out

    def encrypt(self, seed: int, sk, plain):
        out = np.zeros((2, self.L, self.n), dtype=np.uint64)
        rc = lib().orc_encrypt(self.h, seed, _p(_arr(sk)), _p(_arr(plain)), _p(out))
        assert rc == 0
        return out

    def decrypt(self, sk, ct):
        c = _arr(ct)
        npoly, l = c.shape[0], c.shape[1]
        out = np.zeros(self.n, dtype=np.uint64)
        rc = lib().orc_decrypt(self.h, _p(_arr(sk)), _p(c), npoly, l, _p(out))
        assert rc == 0
        return out


def num_threads() -> int:
    return int(lib().orc_num_threads())
