"""GPU parity tests (through the C ABI) for the curve kernels:
  * b200_g{1,2}_mul_batch  == bn128.G{1,2}.MulScalar, X,Y,Z-exact (oracle restates g1.go:140-155)
  * b200_g{1,2}_msm        == the reference's hot loop acc = Add(acc, MulScalar(P_i, w_i))
                              (groth16.go:243-250), compared on affine coordinates
Edge cases as the reference's data has them: (0,0,0) CRS entries, zero / one /
tiny / full-width scalars, repeated points (where the reference's Add itself is
wrong, H6 — expectation from discrete logs)."""
import random

import pytest

from oracle import ref_py as o

pytestmark = pytest.mark.gpu

R = o.R
OG = {1: o.BN.G1, 2: o.BN.G2}


@pytest.fixture(scope="module")
def bn():
    from gosnark_b200 import _lib, bn128
    _lib.init()
    return bn128


def grp(bn, g):
    return bn.G1() if g == 1 else bn.G2()


def aff(g, p):
    """oracle affine with our output convention (x, y, 1) / zeros for infinity."""
    G = OG[g]
    if G.is_zero(p):
        return (0, 0, 0) if g == 1 else ((0, 0), (0, 0), (0, 0))
    a = G.affine(p)
    return (a[0], a[1], 1) if g == 1 else (a[0], a[1], (1, 0))


@pytest.mark.parametrize("g", [1, 2])
def test_mul_batch_exact_jacobian(bn, g):
    G = OG[g]
    rng = random.Random(100 + g)
    base = G.mul_scalar(G.G, 0xABCDEF123456789)            # Jacobian, Z != 1
    inf = G.zero3()
    scalars = [0, 1, 2, 3, R - 1, R - 2, 1 << 253, (1 << 200) - 1] + [rng.randrange(R) for _ in range(24)]
    pts = [base] * (len(scalars) - 2) + [inf, G.G]
    got = grp(bn, g).MulScalarBatch(pts, scalars)
    for p, s, out in zip(pts, scalars, got):
        assert out == G.mul_scalar(p, s)
    # broadcast form: generator times many scalars (CRS minting, groth16.go:139-219)
    got = grp(bn, g).MulScalarBatch([G.G], scalars)
    for s, out in zip(scalars, got):
        assert out == G.mul_scalar(G.G, s)
    assert grp(bn, g).MulScalar(base, 77) == G.mul_scalar(base, 77)


def test_k1_g1_kat_on_gpu(bn):
    """bn128/g1_test.go:11-31 through the CUDA path: 33G + 44G via MSM == 77G."""
    g1 = bn.G1()
    out = g1.MSM([g1.G, g1.G], [33, 44])
    assert out[0] == 0x2f978c0ab89ebaa576866706b14787f360c4d6c3869efe5a72f7c3651a72ff00
    assert out[1] == 0x12e4ba7f0edca8b4fa668fe153aebd908d322dc26ad964d4cd314795844b62b2
    assert out[2] == 1


@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("g,n,c", [(1, 1, 0), (1, 7, 4), (1, 33, 0), (1, 200, 9), (1, 200, 16), (1, 64, 13), (1, 300, 5),
                                   (2, 1, 0), (2, 9, 5), (2, 40, 0), (2, 40, 16), (2, 120, 4)])
def test_msm_vs_reference_loop(bn, g, n, c, mode):
    """Same inputs as the reference loop (Jacobian CRS entries with Z != 1, some
    (0,0,0), full-width / zero / one scalars) -> same affine point."""
    G = OG[g]
    rng = random.Random(1000 * g + n + c)
    ks = [rng.randrange(1, R) for _ in range(n)]
    pts = grp(bn, g).MulScalarBatch([G.G], ks)               # exact Jacobian (checked above)
    scalars = [rng.randrange(R) for _ in range(n)]
    if n > 4:
        pts[2] = G.zero3()                                   # BACDelta[0..NPublic] style entries
        scalars[3] = 0
        scalars[4] = 1
    exp = o.msm_reference_order(G, pts, scalars)
    bs = bn.BaseSet(g, pts, window_bits=c, acc_mode=mode)
    try:
        if mode == 2:
            assert bs.acc_mode() == 2
        assert bs.msm(scalars) == aff(g, exp)
        # a prefix of the base set with fewer scalars (PowersTauDelta[:len(hx)], groth16.go:269-271)
        if n > 4:
            assert bs.msm(scalars[:n - 3]) == aff(g, o.msm_reference_order(G, pts[:n - 3], scalars[:n - 3]))
        assert bs.msm([0] * n) == aff(g, G.zero3())
    finally:
        bs.free()


@pytest.mark.parametrize("g", [1, 2])
def test_msm_repeated_points_and_cancellation(bn, g):
    """Equal points in one bucket must DOUBLE (the reference's Add returns
    infinity for P+P) and P + (-P) must vanish."""
    G = OG[g]
    P = G.mul_scalar(G.G, 12345)
    negP = G.neg(P)
    pts = [P, P, P, negP, P]
    scalars = [5, 5, 5, 5, 7]
    exp = G.mul_scalar(G.G, (12345 * (5 + 5 + 5 - 5 + 7)) % R)
    assert grp(bn, g).MSM(pts, scalars, window_bits=6) == aff(g, exp)
    assert grp(bn, g).MSM([P, negP], [9, 9]) == aff(g, G.zero3())
    assert grp(bn, g).MSM([P, P], [R - 1, 1]) == aff(g, G.zero3())     # (r-1)P + P = rP = O


@pytest.mark.parametrize("mode", [1, 2])
@pytest.mark.parametrize("g,n", [(1, 1 << 14), (2, 1 << 11)])
def test_msm_known_discrete_logs(bn, g, n, mode):
    """SURVEY §8(c) large-N parity: P_i = k_i*G, so sum s_i P_i = (sum s_i k_i mod r)*G."""
    G = OG[g]
    rng = random.Random(77 + g)
    ks = [rng.randrange(1, R) for _ in range(n)]
    pts = grp(bn, g).MulScalarBatch([G.G], ks)
    for i in (0, n // 2, n - 1):
        assert pts[i] == G.mul_scalar(G.G, ks[i])
    bs = bn.BaseSet(g, pts, acc_mode=mode)
    try:
        assert bs.acc_mode() == mode               # both accumulation kernels are covered explicitly
        for dist in ("full", "small", "ones"):
            if dist == "full":
                ss = [rng.randrange(R) for _ in range(n)]
            elif dist == "small":
                ss = [rng.randrange(1 << 16) for _ in range(n)]       # witness-like small values
            else:
                ss = [1] * n
            exp = G.mul_scalar(G.G, sum(k * s for k, s in zip(ks, ss)) % R)
            assert bs.msm(ss) == aff(g, exp), dist
    finally:
        bs.free()


def test_scalar_range_rejected(bn):
    """ABI contract: scalars must be < r (B200_ERANGE), coordinates < q."""
    from gosnark_b200 import _lib
    import numpy as np
    g1 = bn.G1()
    bs = bn.BaseSet(1, [g1.G, g1.G])
    try:
        bad = _lib.ints_to_limbs([R, 1])
        with pytest.raises(_lib.B200Error) as e:
            bs.msm(limbs=bad)
        assert e.value.code == -4
        assert bs.msm([2, 3]) == aff(1, o.BN.G1.mul_scalar(o.BN.G1.G, 5))    # still usable afterwards
    finally:
        bs.free()
    with pytest.raises(_lib.B200Error):
        bn.BaseSet(1, [(o.Q, 2, 1)])


@pytest.mark.timeout(900)
@pytest.mark.parametrize("mode", [1, 2])
@pytest.mark.parametrize("group,logn", [(1, 20), (2, 22)])
def test_msm_full_size_known_discrete_logs(bn, group, logn, mode):
    """BASELINE configs 3 and 5 (G1 MSM N = 2^20, G2 MSM N = 2^22) at full size, under BOTH accumulation kernels:
    P_i = k_i*G minted on the GPU, so sum s_i P_i = (sum s_i k_i mod r)*G — one CPU scalar multiplication gives the exact
    expected point.  Full-width scalars with 1 % zeros and a block of small (witness-like) values;
    linearity: msm(s) + msm(s') = msm(s + s')."""
    import numpy as np
    from gosnark_b200 import _lib, bn128
    n = 1 << logn
    G = OG[group]
    rng = np.random.default_rng(1000 + group)

    def rand_limbs(count):
        a = rng.integers(0, 1 << 63, size=(count, 4), dtype=np.uint64) * np.uint64(2) + \
            rng.integers(0, 2, size=(count, 4), dtype=np.uint64)
        a[:, 3] &= np.uint64((1 << 61) - 1)                 # < 2^253 < r
        return a

    ks = rand_limbs(n)
    ks[:, 0] |= np.uint64(1)                                # non-zero
    words = 12 if group == 1 else 24
    gen = bn128._flatten_g1([bn128.G1.G]) if group == 1 else bn128._flatten_g2([bn128.G2.G])
    pts = np.zeros((n, words), dtype=np.uint64)
    L = _lib.lib()
    _lib.check((L.b200_g1_mul_batch_bcast if group == 1 else L.b200_g2_mul_batch_bcast)(_lib.ptr(gen), _lib.ptr(ks), n, _lib.ptr(pts)))
    bs = bn128.BaseSet(group, limbs=pts, acc_mode=mode)
    try:
        assert bs.acc_mode() == mode
        s1, s2 = rand_limbs(n), rand_limbs(n)
        s1[rng.integers(0, n, size=n // 100)] = 0           # 1 % zero scalars
        s1[: n // 8, 1:] = 0                                # an eighth of the vector: 64-bit values
        kk = _lib.limbs_to_ints(ks)
        exp = []
        for s in (s1, s2):
            sv = _lib.limbs_to_ints(s)
            e = G.affine(G.mul_scalar(G.G, sum(a * b for a, b in zip(kk, sv)) % o.R))
            exp.append(e)
            got = bs.msm(limbs=s)
            assert (got[0], got[1]) == (e[0], e[1])
        ssum = _lib.ints_to_limbs([(a + b) % o.R for a, b in zip(_lib.limbs_to_ints(s1), _lib.limbs_to_ints(s2))])
        got = bs.msm(limbs=ssum)
        one = 1 if group == 1 else (1, 0)
        esum = G.affine(G.add((exp[0][0], exp[0][1], one), (exp[1][0], exp[1][1], one)))
        assert (got[0], got[1]) == (esum[0], esum[1])
    finally:
        bs.free()


@pytest.mark.parametrize("staging", [1, 2])
@pytest.mark.parametrize("g,n", [(1, 1 << 14), (2, 1 << 11)])
def test_msm_staged_backward_pass(bn, g, n, staging):
    """The opt-in backward pass with staged operands (B200_CFG_TMA_STAGING: TMA bulk copies + mbarrier in the contiguous
    rounds, cp.async gathers in round 1) gives the same sums as the default register-load kernel."""
    from gosnark_b200 import _lib
    G = OG[g]
    rng = random.Random(900 + g)
    ks = [rng.randrange(1, R) for _ in range(n)]
    pts = grp(bn, g).MulScalarBatch([G.G], ks)
    bs = bn.BaseSet(g, pts, acc_mode=1)
    try:
        _lib.check(_lib.lib().b200_config(_lib.CFG_TMA_STAGING, staging))
        for dist in ("full", "small"):
            ss = [rng.randrange(R if dist == "full" else 1 << 16) for _ in range(n)]
            ss[0], ss[1] = 0, 1
            exp = G.mul_scalar(G.G, sum(k * s for k, s in zip(ks, ss)) % R)
            assert bs.msm(ss) == aff(g, exp), dist
    finally:
        _lib.check(_lib.lib().b200_config(_lib.CFG_TMA_STAGING, 0))
        bs.free()
