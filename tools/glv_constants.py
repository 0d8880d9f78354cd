#!/usr/bin/env python3
"""Derive and check the BN254 GLV constants used by k_groth16_products (csrc/prove_host.cuh):
lambda (cube root of unity mod r), beta (mod q) with phi(x, y) = (beta x, y) = lambda (x, y), the short lattice
basis (a1, b1), (a2, b2) of {(a, b): a + b lambda = 0 mod r}, and g_i = round(2^256 |.| / r).  Uses the oracle."""
import math
import os
import random
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle import ref_py as o  # noqa: E402

Q, R = o.Q, o.R
L = 4407920970296243842393367215006156084916469457145843978461
B = 2203960485148121921418603742825762020974279258880205651966
G1 = o.BN.G1
P = G1.affine(G1.mul_scalar(G1.G, 987654321))
assert pow(L, 3, R) == 1 and pow(B, 3, Q) == 1
assert G1.affine(G1.mul_scalar((P[0], P[1], 1), L)) == (B * P[0] % Q, P[1])


def basis(n, lam):
    sq = math.isqrt(n)
    r0, r1, t0, t1 = n, lam, 0, 1
    rs, ts = [], []
    while r1:
        q = r0 // r1
        r0, r1 = r1, r0 - q * r1
        t0, t1 = t1, t0 - q * t1
        rs.append(r0)
        ts.append(t0)
    for i, (r_, t_) in enumerate(zip(rs, ts)):
        if r_ < sq:
            cands = [(rs[i - 1], -ts[i - 1])] + ([(rs[i + 1], -ts[i + 1])] if i + 1 < len(rs) else [])
            a2, b2 = min(cands, key=lambda v: v[0] ** 2 + v[1] ** 2)
            return r_, -t_, a2, b2


a1, b1, a2, b2 = basis(R, L)
assert a1 * b2 - a2 * b1 == R and a1 > 0 and b1 < 0 and a2 > 0 and b2 > 0
g1 = ((b2 << 256) + R // 2) // R
g2 = ((-b1 << 256) + R // 2) // R
mx = 0
for _ in range(20000):
    k = random.randrange(R)
    c1, c2 = (k * g1 + (1 << 255)) >> 256, (k * g2 + (1 << 255)) >> 256
    k1, k2 = k - c1 * a1 - c2 * a2, c1 * (-b1) - c2 * b2
    assert (k1 + k2 * L - k) % R == 0
    mx = max(mx, abs(k1).bit_length(), abs(k2).bit_length())
print(f"a1={a1:#x} b1={b1:#x} a2={a2:#x} b2={b2:#x} g1={g1:#x} g2={g2:#x} max|k_i| bits={mx}")
print("beta (Montgomery, u32 limbs):", ", ".join(f"0x{(((B << 256) % Q) >> (32 * i)) & 0xffffffff:08x}u" for i in range(8)))
