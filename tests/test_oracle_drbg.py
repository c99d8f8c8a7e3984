"""Pins of oracle/drbg_oracle.py on the reference's NIST CTR_DRBG vectors (NistCtrDrbgTests.swift:22-160)."""
import numpy as np
import pytest

from oracle import drbg_oracle as drbg
from oracle import oracle as orc

VECTORS = [
    ("69a09f6bf5dda15cd4af29e14cf5e0cddd7d07ac39bba587f8bc331104f9c448",
     "f78a4919a6ec899f7b6c69381febbbe083315f3d289e70346db0e4ec4360473ae0b3d916e9b6b964309f753ed66ae59de48da316cc1944bc8dfd0e2575d0ff6d"),
    ("80bfbd340d79888f34f043ed6807a9f28b72b6644d9d9e9d777109482b80788a",
     "80db048d2f130d864b19bfc547c92503e580cb1a8e1f74f3d97fdda6501fb1aa81fcedac0dd18b6ccfdc183ca28a44fc9f3a08834ba8751a2f4495367c54a185"),
    ("a559ac9872791d79197e54da70a8d858fbe39e8514d2c86a7bcffadc68782edf",
     "d14b72e17c2f6f77b46d0717b788420e503bb18de542135f586a90c5c73fceeee50fd1633b5b09ab061b9367ca785ecb400e1f3681583661aaf8352184454ae6"),
    ("300fe148dd39de1edb993ca5260373b3f5f09a5cf7a32b0c41fe6224f981d3b1",
     "deea89b5128fb992696d7b97ebc2c0793614b172f4c75bb83c12a1b389bac3bfecb773cd7717583c2b61b3b243ac9683dba4fbc07182bad8271a7f16d833e4d9"),
    ("0c6ee2a5d46325baa8e9a3f6b598fc790c513d387d47001116d19a614d2038c4",
     "f1ee11be189263fed9932c1192219d00378e36ce81a431318545da9f81f50c2913d1f7be499ce9e1e39f93ee2360668f127340691c17711707cf5f1f8a4d93ee"),
    ("bdbba1ad4803fdc783ef5d6e2aa66dc948e960bc11cca89a60cff5c60e984302",
     "260a32c3973750e0c10f7f7495d46e7c3691c27a58e828cdef48ef660716f771d61c3c76db407d816066f5afbf16993485cdb653d418dd65ffa5d3825732b8cb"),
    ("22587bfdce62f4afc1dd2673f5308364f27db9912ad01b045e74db4518435959",
     "c904d03089b7dd1f17564a7ef70b17bb1b29c0c1793cc8d92b8c158c04ca5366919f8caf544d5d07c28abe6d14baaa0c56602df1c373e9acc419e3c932e577e6"),
    ("8abefbb23dfd58d82b88a4c4fcfcee183ce01db975edeeb404bd216e6177ea0d",
     "8a708e8a99035389a4d66d57d12f488ecba57a3b2ca78015bedae06aaa414d791196e262b28fbd745dff94f8fe600687c9ce2f50cf6d79d39b8c5ea36533755d"),
    ("c45c9fec6bb83fb08008877c70b632d792119a35c4c5988c4026cf3f8612b800",
     "84430e49a9b4d395d055ca0efdf285a7551c5f7119dbea5c10daaa9e8be041e23e9bc893c90a35b77b19dc202ec834172e6c8cea97c9d7c68df1374aeea94537"),
    ("58cbccd7f86e5f0472dcb377f598f2d42ed96afdf0c8e45f12c4ff4a969c5b6b",
     "41ff55d058beaa04308bd0b39d4801f70f23d829037e4cc9b2ea0eacf5aef9b8e33fc59c528b53bce08d2b536d37bf194c797f03290494dd00ef244ac223e350"),
    ("d50558dfb7a8966c63b3a1d0a837970ad0bff5adbd8adacae5d3accfde64cd4d",
     "e91361511d926be4d997fc970b1a5dcdb33a711f215cbdbffabfcdaa6248596891d55a9e64f4e9f5185ed7056f7cbb42f474a23542fe9e9c2495182cefb38a6a"),
    ("f70ce283efd5ba36c284cb267d22e23dc41671b2aaae98e638c6e451bc9c3cbb",
     "fd9b3b53e12b6702e4c6e4acac33aeae5ceb34cebfffa7007cb1ab1c3b4be1a38e5c86dea0775ab0c89ae135e0b36da087921d3ff275ffc8e5dcee6e3d66ee43"),
    ("58eb544f44dfe1048a8113d4b6909050abf9010036233be7f8fcc41f39baff9c",
     "5c6aedc020e764f4d3bb8abc2907c9c604dd98e1cfc2882ea72d554e39fe86463a51886d980ac8cdda0f4e584226d45344e43dd84e8430f58c3880a0ce930863"),
    ("b694ce5f4d9af4ce93626636c9ecb341f3f5152fd580745202cd0c83f4d5b4c5",
     "78b32d396f5a919f5ccb9be2afaf5f6212d75bf084e99357e28ccc98d433696455b10a85ecaf61686a96606ff3e8962321358a56fa53cabbf16c65c1c32debcd"),
    ("42cb183d2a04c89c69efbcec08bee2003b9a1cd56878a774f0162bf70f2c708f",
     "cb4afdec033b42949ebbb27245fd33c1503c1278027e11a1f050e04080abe4850821b71ed5a6bd83da6bde8e56c5faed49da26887028bab807d1ad055e2a8a27"),
]


def test_ctr_drbg_state_vector():
    """NistCtrDrbgTests.vector (:22-55): key / V after instantiate and after each generate."""
    prng = drbg.NistCtrDrbg(bytes.fromhex(VECTORS[0][0]))
    assert prng.key.hex() == "314263a50fa3913de2d034b6e812a597"
    assert prng.nonce.to_bytes(16, "big").hex() == "def5dd62590d06150b94f1a8754b3a30"
    prng.generate(64)
    assert prng.key.hex() == "4b0f2ae7d0b330fa709b0844c7eedb5c"
    assert prng.nonce.to_bytes(16, "big").hex() == "dae190eb55353de50e494cdef2a544d4"
    out = prng.generate(64)
    assert prng.key.hex() == "b4d5d6de074612076e496f241ebcf017"
    assert prng.nonce.to_bytes(16, "big").hex() == "034eeae49adbdfccff79bfdc0d83ed70"
    assert out.hex() == VECTORS[0][1]


@pytest.mark.parametrize("entropy,expected", VECTORS)
def test_ctr_drbg_nist_vectors(entropy, expected):
    prng = drbg.NistCtrDrbg(bytes.fromhex(entropy))
    prng.generate(len(expected) // 2)
    assert prng.generate(len(expected) // 2).hex() == expected


def test_buffered_stream_and_uniform_poly():
    seed = bytes(range(32))
    a, b = drbg.NistAes128Ctr(seed), drbg.NistCtrDrbg(seed)
    stream = b.generate(4096) + b.generate(4096) + b.generate(4096)
    assert a.fill(100) + a.fill(5000) + a.fill(7188) == stream      # any request pattern reads the same stream
    n = 512
    moduli = orc.generate_primes([55, 40, 61], False, n)
    poly = drbg.random_poly(n, moduli, seed)
    words = drbg.NistAes128Ctr(seed).fill(3 * n * 16)
    for r, q in enumerate(moduli):
        for c in (0, 1, n - 1):
            k = r * n + c
            assert int(poly[r, c]) == int.from_bytes(words[16 * k:16 * k + 16], "little") % q
    assert all(int(poly[r].max()) < q for r, q in enumerate(moduli))


def test_device_aes_drbg_functions_emulated_on_host():
    """The __host__ __device__ AES-128 / CTR_DRBG functions of csrc/drbg.cuh, replayed on the CPU, produce the oracle's
    (NIST-pinned) stream."""
    import os
    import shutil
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    src, binary = os.path.join(root, "tests", "emu", "drbg_emulate.cu"), os.path.join(root, "tests", "emu", "drbg_emulate")
    hdr = os.path.join(root, "swift-homomorphic-encryption_b200", "csrc", "drbg.cuh")
    if not os.path.exists(binary) or os.path.getmtime(binary) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call([nvcc, "-O1", "-std=c++17", "-Wno-deprecated-gpu-targets", "-o", binary, src])
    for seed_hex in (VECTORS[0][0], VECTORS[5][0], bytes(range(32)).hex()):
        count = 3 * 4096 + 100
        got = subprocess.run([binary, seed_hex, str(count)], capture_output=True, text=True, check=True).stdout.strip()
        assert got == drbg.NistAes128Ctr(bytes.fromhex(seed_hex)).fill(count).hex()
    # S-box spot values (FIPS-197 figure 7) through the first round key of the all-zero key: E_0(0) is the FIPS-197 KAT
    got = subprocess.run([binary, "00" * 32, "16"], capture_output=True, text=True, check=True).stdout.strip()
    assert got == drbg.NistAes128Ctr(bytes(32)).fill(16).hex()
