"""Pin the C oracle (oracle/ref_c.c, the timed CPU baseline) against the Python
oracle (itself pinned against the Go binary): X,Y,Z-exact in reference order."""
import ctypes
import random

import numpy as np
import pytest

import build as b200build
from gosnark_b200._lib import ints_to_limbs, limbs_to_ints, ptr
from gosnark_b200.bn128 import _flatten_g1, _flatten_g2, _unflatten_g1, _unflatten_g2
from oracle import ref_py as o

R = o.R


@pytest.fixture(scope="module")
def oc():
    return ctypes.CDLL(b200build.build_oracle())


@pytest.mark.parametrize("g", [1, 2])
def test_mul_scalar_and_loop_exact(oc, g):
    G = o.BN.G1 if g == 1 else o.BN.G2
    flat, unflat, words = (_flatten_g1, _unflatten_g1, 12) if g == 1 else (_flatten_g2, _unflatten_g2, 24)
    rng = random.Random(g)
    n = 6
    pts = [G.mul_scalar(G.G, rng.randrange(1, R)) for _ in range(n)]
    pts[2] = G.zero3()
    sc = [rng.randrange(R) for _ in range(n)]
    sc[1], sc[4] = 0, 1
    mul = oc.oc_g1_mul_scalar if g == 1 else oc.oc_g2_mul_scalar
    for p, s in zip(pts, sc):
        out = np.zeros(words, dtype=np.uint64)
        mul(ptr(flat([p])), ptr(ints_to_limbs([s])), ptr(out))
        assert unflat(out)[0] == G.mul_scalar(p, s)
    loop = oc.oc_g1_msm_loop if g == 1 else oc.oc_g2_msm_loop
    out = np.zeros(words, dtype=np.uint64)
    loop(ptr(flat(pts)), ptr(ints_to_limbs(sc)), ctypes.c_long(n), 1, ptr(out))
    assert unflat(out)[0] == o.msm_reference_order(G, pts, sc)          # reference order: X,Y,Z-exact
    loop(ptr(flat(pts)), ptr(ints_to_limbs(sc)), ctypes.c_long(n), 3, ptr(out))
    assert G.affine(unflat(out)[0]) == G.affine(o.msm_reference_order(G, pts, sc))   # threaded: same point
