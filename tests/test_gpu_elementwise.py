"""GPU parity for the coefficient-wise PolyRq operations against the oracle's poly ops (pinned on
PolyRqTests.swift:46-143) and exact integer arithmetic."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import hecuda
from oracle import oracle as orc


@pytest.mark.parametrize("n,bits", [(2, [30, 30]), (16, [55, 52, 62, 58]), (64, [61, 61, 40]), (4096, [27, 28, 28]),
                                    (8192, [55, 55, 55, 55])])
def test_elementwise_ops_match_oracle(n, bits):
    moduli = orc.generate_primes(bits, False, n)
    g = hecuda.Context(n, moduli, 2)
    L = g.L
    q = moduli[:L]
    a = orc.fill_uniform(1, q, n, 3 * L).reshape(3, L, n)
    b = orc.fill_uniform(2, q, n, 3 * L).reshape(3, L, n)
    a[0, :, 0] = 0
    b[0, :, 1 % n] = 0
    for name, fn in (("add", hecuda.Bfv.polyAdd), ("sub", hecuda.Bfv.polySub), ("mul", hecuda.Bfv.polyMul)):
        got = fn(g, a, b)
        for i in range(3):
            assert np.array_equal(got[i], orc.poly_op(name, n, q, a[i], b[i])), (name, i)
    neg = hecuda.Bfv.polyNeg(g, a)
    scalars = [int(m) - 3 for m in q]
    scaled = hecuda.Bfv.polyMulScalars(g, a, scalars)
    for r, m in enumerate(q):
        assert neg[:, r].tolist() == [[(int(m) - int(v)) % int(m) for v in row] for row in a[:, r]]
        assert scaled[:, r].tolist() == [[int(v) * scalars[r] % int(m) for v in row] for row in a[:, r]]
    # a single row (PolyRq over one modulus) and the extended base used inside multiply
    one = hecuda.Bfv.polyMul(g, a[:, :1], b[:, :1])
    assert np.array_equal(one[1, 0], orc.poly_op("mul", n, q[:1], a[1, :1], b[1, :1])[0])
    with pytest.raises(hecuda.HeError):
        hecuda.Bfv.polyMulScalars(g, a, [int(m) for m in q])   # scalars must be reduced
    with pytest.raises(hecuda.HeError):
        hecuda.Bfv.polyAdd(g, a, b[:2])
    g.close()
