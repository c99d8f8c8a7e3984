"""Generates the committed golden fixtures from the KAT-pinned CPU oracle.  Run from the repo root:
    python tests/golden/make_golden.py
ntt_kats.json holds the reference's own NTT vectors (Tests/HomomorphicEncryptionTests/NttTests.swift:73-191)
verbatim; mul_n64.npz holds oracle-generated full-pipeline ciphertexts (the reference pins none)."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oracle as orc  # noqa: E402

KATS = [
    ("ntt2_a", 97, [1, 2], [45, 54]),
    ("ntt2_b", 113, [3, 4], [63, 56]),
    ("ntt4_a", 97, [1, 2, 3, 4], [30, 7, 64, 0]),
    ("ntt4_b", 113, [5, 6, 7, 8], [108, 31, 103, 4]),
    ("ntt8", 4_194_353, [1, 2, 3, 4, 5, 6, 7, 8],
     [3_372_683, 765_982, 387_853, 2_657_954, 2_013_665, 1_280_882, 2_457_874, 3_840_527]),
    ("ntt16", 536_870_849,
     [477_051_601, 421_524_611, 456_257_859, 247_136_825, 128_775_020, 76_785_070, 49_764_016, 525_812_772,
      325_605_371, 88_935_943, 255_470_762, 39_507_048, 404_978_219, 379_383_003, 244_420_585, 346_826_612],
     [230_846_094, 480_599_401, 157_364_576, 360_442_736, 531_052_463, 294_311_347, 432_899_854, 219_721_533,
      286_807_067, 260_650_843, 362_842_688, 315_862_017, 493_042_020, 520_739_674, 167_758_416, 370_401_491]),
    ("ntt32", 769,
     [401, 203, 221, 352, 487, 151, 405, 356, 343, 424, 635, 757, 457, 280, 624, 353,
      496, 353, 624, 280, 457, 757, 635, 424, 343, 356, 405, 151, 487, 352, 221, 203], list(range(1, 33))),
    ("ntt4096_delta", 557_057, [1] + [0] * 4095, [1] * 4096),
]


def main():
    kats = [{"name": k[0], "n": len(k[2]), "modulus": k[1], "coeff": k[2], "eval": k[3]} for k in KATS]
    for kat in kats:  # sanity: the oracle reproduces them
        got = orc.ntt_forward(kat["n"], [kat["modulus"]], [kat["coeff"]])[0].tolist()
        assert got == kat["eval"], kat["name"]
    with open(os.path.join(HERE, "ntt_kats.json"), "w") as f:
        json.dump(kats, f)

    n = 64
    moduli = orc.generate_primes([55] * 4, False, n)
    t = orc.generate_primes([17], True, n)[0]
    ctx = orc.Context(n, moduli, t)
    L = ctx.L
    sk, rk = ctx.keygen(2024)
    rng = np.random.default_rng(7)
    batch = 3
    a = np.stack([ctx.encrypt(10 + k, sk, rng.integers(0, t, n, dtype=np.uint64)) for k in range(batch)])
    b = np.stack([ctx.encrypt(20 + k, sk, rng.integers(0, t, n, dtype=np.uint64)) for k in range(batch)])
    a[2] = orc.fill_uniform(5, ctx.q, n, 2 * L).reshape(2, L, n)  # one uniform (non-decryptable) pair too
    product = ctx.mul(a, b)
    relinearized = ctx.relinearize(product, rk)
    switched = ctx.mod_switch_down(relinearized)
    np.savez_compressed(os.path.join(HERE, "mul_n64.npz"), n=n, moduli=np.array(moduli, dtype=np.uint64), t=t,
                        secret_key=sk, relin_key=rk, a=a, b=b, product=product, relinearized=relinearized,
                        switched=switched)
    print("wrote", os.listdir(HERE))


if __name__ == "__main__":
    main()
