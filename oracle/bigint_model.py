"""Independent pure-Python (arbitrary-precision int) model of the hot path -- TEST INFRASTRUCTURE ONLY.

Written from the mathematical definition of each reference step (no lazy reductions, no Barrett/Shoup):
every stored residue is the canonical representative, so this model and he_oracle.c must agree bit for bit.
Quadratic-time transforms: small N only.  Reference lines cited per function (relative to /root/reference/).
"""
from __future__ import annotations

from math import prod

M_TILDE = 1 << 32  # Sources/ModularArithmetic/Scalar.swift:522-524


def bitrev(i: int, bits: int) -> int:
    return int(format(i, f"0{bits}b")[::-1], 2) if bits else 0


def min_root(degree: int, p: int) -> int:  # PolyRq+Ntt.swift:87-105
    g = 2
    while True:
        r = pow(g, (p - 1) // degree, p)
        if pow(r, degree // 2, p) == p - 1:
            break
        g += 1
    return min(pow(r, k, p) for k in range(1, degree, 2))


def ntt_forward(row, p):  # PolyRq+Ntt.swift:237-319 -- out[i] = poly(psi^(2 bitrev(i)+1))
    n = len(row)
    bits = n.bit_length() - 1
    psi = min_root(2 * n, p)
    out = []
    for i in range(n):
        w = pow(psi, 2 * bitrev(i, bits) + 1, p)
        acc, wk = 0, 1
        for c in row:
            acc = (acc + c * wk) % p
            wk = wk * w % p
        out.append(acc)
    return out


def ntt_inverse(row, p):  # PolyRq+Ntt.swift:379-483
    n = len(row)
    bits = n.bit_length() - 1
    psi = min_root(2 * n, p)
    ninv = pow(n, -1, p)
    winv = [pow(pow(psi, 2 * bitrev(i, bits) + 1, p), -1, p) for i in range(n)]
    out = []
    for k in range(n):
        acc = 0
        for i in range(n):
            acc = (acc + row[i] * pow(winv[i], k, p)) % p
        out.append(acc * ninv % p)
    return out


def fast_base_conv(cols, q, out_moduli):  # RnsBaseConverter.swift:68-143 (no a_x*q correction)
    Q = prod(q)
    res = []
    for x in cols:  # x = residues of one coefficient in base q
        y = [xi * pow(Q // qi, -1, qi) % qi for xi, qi in zip(x, q)]
        res.append([sum(yi * (Q // qi) for yi, qi in zip(y, q)) % m for m in out_moduli])
    return res


def lift(cols, q, bsk):  # RnsTool.swift:313-368
    Q = prod(q)
    scaled = [[xi * M_TILDE % qi for xi, qi in zip(x, q)] for x in cols]
    conv = fast_base_conv(scaled, q, bsk + [M_TILDE])
    out = []
    for x, v in zip(cols, conv):
        r = (-v[-1] * pow(Q, -1, M_TILDE)) % M_TILDE
        rc = r if r < M_TILDE // 2 else r - M_TILDE
        out.append(list(x) + [(vj + Q * rc) * pow(M_TILDE, -1, b) % b for vj, b in zip(v[:-1], bsk)])
    return out


def floor_qbsk_to_q(cols, q, bsk):  # RnsTool.swift:378-456
    Q = prod(q)
    L = len(q)
    B = bsk[:-1]
    msk = bsk[-1]
    Bprod = prod(B)
    conv = fast_base_conv([c[:L] for c in cols], q, bsk)
    out = []
    for c, s in zip(cols, conv):
        f = [(c[L + j] - s[j]) * pow(Q, -1, b) % b for j, b in enumerate(bsk)]  # approximateFloor
        conv_msk = fast_base_conv([f[:-1]], B, [msk])[0][0]
        alpha = (conv_msk - f[-1]) * pow(Bprod, -1, msk) % msk
        alpha_c = alpha - msk if alpha > msk // 2 else alpha  # centered, :419-445
        conv_q = fast_base_conv([f[:-1]], B, q)[0]
        out.append([(v - alpha_c * Bprod) % qi for v, qi in zip(conv_q, q)])
    return out


def cols_of(rows):
    return [list(c) for c in zip(*rows)]


def rows_of(cols):
    return [list(r) for r in zip(*cols)]


def bfv_mul(a, b, q, bsk, t):  # Bfv+Multiply.swift:18-85; a, b = [poly][row][coeff] with 2 polys
    qbsk = q + bsk

    def behz(poly):
        lifted = rows_of(lift(cols_of(poly), q, bsk))
        return [ntt_forward(r, m) for r, m in zip(lifted, qbsk)]

    l0, l1 = behz(a[0]), behz(a[1])
    r0, r1 = behz(b[0]), behz(b[1])
    polys = [[], [], []]
    for ri, m in enumerate(qbsk):
        polys[0].append([x * y % m for x, y in zip(l0[ri], r0[ri])])
        polys[1].append([(x * y + u * v) % m for x, y, u, v in zip(l0[ri], r1[ri], l1[ri], r0[ri])])
        polys[2].append([x * y % m for x, y in zip(l1[ri], r1[ri])])
    out = []
    for poly in polys:
        coeff = [ntt_inverse([v * t % m for v in r], m) for r, m in zip(poly, qbsk)]
        out.append(rows_of(floor_qbsk_to_q(cols_of(coeff), q, bsk)))
    return out


def divide_round_qlast(rows, moduli):  # PolyRq.swift:365-393
    ql = moduli[-1]
    half = ql >> 1
    last = [(v + half) % ql for v in rows[-1]]
    out = []
    for r, qi in zip(rows[:-1], moduli[:-1]):
        inv = pow(ql, -1, qi)
        out.append([((x + half % qi - (lv % qi)) * inv) % qi for x, lv in zip(r, last)])
    return out


def keyswitch_update(target, q_all, l, ksk):  # Bfv+Keys.swift:123-208
    """target: l rows (Coeff); q_all = [q_0..q_{L-1}, q_ks]; ksk[j][comp][row][coeff] (Eval, K = L+1 rows)."""
    L = len(q_all) - 1
    ksm = q_all[:l] + [q_all[L]]
    prods = [[], []]
    for r, m in enumerate(ksm):
        key_index = L if r == l else r
        digits = [ntt_forward([v % m for v in target[j]], m) for j in range(l)]
        for comp in range(2):
            prods[comp].append([sum(digits[j][c] * ksk[j][comp][key_index][c] for j in range(l)) % m
                                for c in range(len(target[0]))])
    out = []
    for comp in range(2):
        coeff = [ntt_inverse(row, m) for row, m in zip(prods[comp], ksm)]
        out.append(divide_round_qlast(coeff, ksm))
    return out
