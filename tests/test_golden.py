"""The committed golden fixtures still agree with the oracle (guards against silent oracle drift)."""
import json
import os

import numpy as np

from oracle import oracle as orc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_ntt_kats_fixture():
    for kat in json.load(open(os.path.join(GOLDEN, "ntt_kats.json"))):
        assert orc.ntt_forward(kat["n"], [kat["modulus"]], [kat["coeff"]])[0].tolist() == kat["eval"]
        assert orc.ntt_inverse(kat["n"], [kat["modulus"]], [kat["eval"]])[0].tolist() == kat["coeff"]


def test_mul_fixture():
    z = np.load(os.path.join(GOLDEN, "mul_n64.npz"))
    ctx = orc.Context(int(z["n"]), [int(v) for v in z["moduli"]], int(z["t"]))
    assert np.array_equal(ctx.mul(z["a"], z["b"]), z["product"])
    assert np.array_equal(ctx.relinearize(z["product"], z["relin_key"]), z["relinearized"])
    assert np.array_equal(ctx.mod_switch_down(z["relinearized"]), z["switched"])
    # decryptability of the encrypted pairs
    for k in range(2):
        m = ctx.decrypt(z["secret_key"], z["switched"][k])
        assert m.shape == (int(z["n"]),)
