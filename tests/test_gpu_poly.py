"""GPU parity tests for the F_r polynomial kernels behind
PolynomialField.Mul / Div / DivisorPolynomial (r1csqap/r1csqap.go:57-84,213-216)."""
import random

import pytest

from oracle import ref_py as o

pytestmark = pytest.mark.gpu
R = o.R


@pytest.fixture(scope="module")
def pf():
    from gosnark_b200 import _lib, r1csqap
    _lib.init()
    return r1csqap.PolynomialField()


def test_literals(pf):                                   # r1csqap/r1csqap_test.go:59-88
    a, b = [1, 0, 5], [3, 0, 1]
    c = pf.Mul(a, b)
    assert c == [3, 0, 16, 0, 5]
    q, r = pf.Div(c, a)
    assert q == b and all(x == 0 for x in r)


@pytest.mark.parametrize("la,lb", [(1, 1), (2, 1), (1, 7), (5, 5), (33, 17), (100, 257), (513, 511)])
def test_mul_vs_oracle(pf, la, lb):
    rng = random.Random(la * 1000 + lb)
    a = [rng.randrange(R) for _ in range(la)]
    b = [rng.randrange(R) for _ in range(lb)]
    assert pf.Mul(a, b) == o.PF.mul(a, b)


@pytest.mark.parametrize("la,lb", [(1, 1), (5, 1), (5, 5), (6, 5), (13, 7), (64, 33), (300, 120), (1023, 513)])
def test_div_vs_oracle(pf, la, lb):
    """Quotient AND remainder equal the reference's long division, including
    non-monic divisors and non-zero remainders."""
    rng = random.Random(la * 7919 + lb)
    a = [rng.randrange(R) for _ in range(la)]
    b = [rng.randrange(R) for _ in range(lb - 1)] + [rng.randrange(1, R)]
    q, rem = pf.Div(a, b)
    oq, orem = o.PF.div(a, b)
    assert q == oq
    assert rem == orem
    assert pf.DivisorPolynomial(a, b) == oq


def test_div_edge_cases(pf):
    from gosnark_b200 import _lib
    a = [5, 6, 7]
    assert pf.Div(a, [1, 2, 3, 4]) == ([], a)                       # len(a) < len(b): loop never runs
    with pytest.raises(_lib.B200Error) as e:                        # reference: ModInverse(0) -> panic
        pf.Div([1, 2, 3], [1, 0])
    assert e.value.code == -5
    neg = [-1, 2, -3]                                               # big.Int negatives reduce mod r
    assert pf.Mul(neg, [1]) == [x % R for x in neg]


def test_config1_divisor_polynomial(pf, golden_dir):
    """groth16_test.go:77-86 on the GPU: px == hx*Z, rem == 0, len(hx) == len(px)-len(Z)+1."""
    import json, os
    g = json.load(open(os.path.join(golden_dir, "gobin_x3x5.json")))
    px, z = g["px"], g["groth16_setup"]["Pk"]["Z"]
    hx, rem = pf.Div(px, z)
    assert len(hx) == len(px) - len(z) + 1 and all(x == 0 for x in rem)
    assert pf.Mul(hx, z) == px
    assert hx == o.PF.divisor_polynomial(px, z)


@pytest.mark.parametrize("n", [1 << 12, 1 << 16])
def test_exact_division_large(pf, n):
    """Full-size shape of config 2: px (2n-1 coeffs) / Z (n+1 coeffs, monic) -> h (n-1).
    px is minted as h*Z + low-degree noise-free, so h must come back exactly."""
    rng = random.Random(n)
    h = [rng.randrange(R) for _ in range(n - 1)]
    z = [rng.randrange(R) for _ in range(n)] + [1]
    px = pf.Mul(h, z)
    assert len(px) == 2 * n - 1
    # spot-check the product against a direct evaluation at a random point
    x = rng.randrange(R)
    ev = lambda p: sum(c * pow(x, i, R) for i, c in enumerate(p[:64])) % R if len(p) <= 64 else None
    hx = pf.DivisorPolynomial(px, z)
    assert hx == h
    # and with a non-zero remainder the quotient is unchanged
    px2 = list(px)
    px2[3] = (px2[3] + 12345) % R
    q2, rem2 = pf.Div(px2, z)
    assert q2 == h and rem2[3] == 12345 and sum(rem2) == 12345
