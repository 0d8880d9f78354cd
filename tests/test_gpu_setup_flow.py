"""GPU: trusted-setup minting and the full flow
R1CS -> QAP -> setup -> prove -> verify, every heavy step on the GPU, verified by the oracle's pairing
check (groth16/groth16_test.go:16-107 is the model; the reference itself stops working at 22 constraints)."""
import random

import pytest

from oracle import ref_py as o

pytestmark = pytest.mark.gpu
R = o.R
G1, G2 = o.BN.G1, o.BN.G2

# groth16_test.go:20-28 circuit (y = x^3 + x + 5), R1CS literals of circuit_test.go:39-69
A = [[0, 0, 1, 0, 0, 0, 0, 0], [0, 0, 0, 1, 0, 0, 0, 0], [0, 0, 1, 0, 1, 0, 0, 0], [5, 0, 0, 0, 0, 1, 0, 0],
     [0, 0, 0, 0, 0, 0, 1, 0], [0, 1, 0, 0, 0, 0, 0, 0], [1, 0, 0, 0, 0, 0, 0, 0]]
B = [[0, 0, 1, 0, 0, 0, 0, 0], [0, 0, 1, 0, 0, 0, 0, 0], [1, 0, 0, 0, 0, 0, 0, 0], [1, 0, 0, 0, 0, 0, 0, 0],
     [1, 0, 0, 0, 0, 0, 0, 0], [1, 0, 0, 0, 0, 0, 0, 0], [1, 0, 0, 0, 0, 0, 0, 0]]
C = [[0, 0, 0, 1, 0, 0, 0, 0], [0, 0, 0, 0, 1, 0, 0, 0], [0, 0, 0, 0, 0, 1, 0, 0], [0, 0, 0, 0, 0, 0, 1, 0],
     [0, 1, 0, 0, 0, 0, 0, 0], [0, 0, 0, 0, 0, 0, 1, 0], [0, 0, 0, 0, 0, 0, 0, 1]]
W = [1, 35, 3, 9, 27, 30, 35, 1]


@pytest.fixture(scope="module")
def mods():
    from gosnark_b200 import _lib, groth16, r1csqap
    _lib.init()
    return groth16, r1csqap.PolynomialField()


def test_setup_exact_vs_oracle_and_flow(mods):
    """TestGroth16MinimalFlow on the GPU path; the minted CRS equals the oracle's X,Y,Z-exactly for the same toxic values."""
    groth16, pf = mods
    circuit = {"NVars": 8, "NPublic": 1}
    alphas, betas, gammas, z = pf.R1CSToQAP(A, B, C)
    assert len(alphas) == 8 and len(alphas[0]) == 7                      # groth16_test.go:57-60
    ax, bx, cx, px = pf.CombinePolynomials(W, alphas, betas, gammas)
    assert len(ax) == 7 and len(px) == 13                                # :61-66
    tox = {"T": 0x1234567890abcdef1234567890abcdef, "Kalpha": 0x1111, "Kbeta": 0x2222222222, "Kgamma": 0x3333, "Kdelta": 0x444444}
    setup = groth16.GenerateTrustedSetup(len(W), circuit, alphas, betas, gammas, toxic=tox)
    opk, ovk = o.groth16_setup(8, 1, alphas, betas, gammas, tox)
    pk, vk = setup["Pk"], setup["Vk"]
    assert pk["Z"] == opk["Z"] == z
    assert pk["PowersTauDelta"] == opk["PowersTauDelta"]                 # Jacobian X,Y,Z identical
    assert pk["G1"]["At"] == opk["G1"]["At"] and pk["G1"]["BACGamma"] == opk["G1"]["BACGamma"]
    assert pk["G2"]["BACGamma"] == opk["G2"]["BACGamma"] and pk["BACDelta"] == opk["BACDelta"]
    assert vk["IC"] == ovk["IC"] and vk["G2"]["Gamma"] == ovk["G2"]["Gamma"]
    hx = pf.DivisorPolynomial(px, pk["Z"])
    assert pf.Mul(hx, pk["Z"]) == px and len(hx) == len(px) - len(pk["Z"]) + 1      # :77-86
    proof = groth16.GenerateProofs(circuit, pk, W, px)
    assert o.groth16_verify(vk, proof, [35])                             # :100
    assert not o.groth16_verify(vk, proof, [34])                         # :106


def test_flow_beyond_reference_limit(mods):
    """n = 64 constraints (the reference's R1CSToQAP breaks at 22): synthetic multiplication chain
    x_{k+1} = x_k * x_{k-1}, m = n + 2 signals [one, pub, x0, x1, ..., out]; QAP, setup, proof on the GPU;
    the oracle's pairing check accepts the proof for the right public input only."""
    groth16, pf = mods
    n = 64
    m = n + 2
    rng = random.Random(5)
    x = [rng.randrange(R), rng.randrange(R)]
    # signals: 0 one, 1 pub, 2.. chain values v_0..v_{n+1}?  keep m = n + 2: one, pub, v0..v_{n-1}
    # constraints k = 0..n-3: v_{k+2} = v_{k+1} * v_k ; k = n-2: pub = v_{n-1} * one ; k = n-1: one * one = one
    v = [x[0], x[1]]
    for k in range(n - 2):
        v.append(v[-1] * v[-2] % R)
    pub = v[n - 1]
    w = [1, pub] + v
    assert len(w) == m
    a = [[0] * m for _ in range(n)]
    b = [[0] * m for _ in range(n)]
    c = [[0] * m for _ in range(n)]
    for k in range(n - 2):
        a[k][2 + k + 1] = 1
        b[k][2 + k] = 1
        c[k][2 + k + 2] = 1
    a[n - 2][2 + n - 1] = 1; b[n - 2][0] = 1; c[n - 2][1] = 1
    a[n - 1][0] = 1; b[n - 1][0] = 1; c[n - 1][0] = 1
    alphas, betas, gammas, z = pf.R1CSToQAP(a, b, c)
    ax, bx, cx, px = pf.CombinePolynomials(w, alphas, betas, gammas)
    hx, rem = pf.Div(px, z)
    assert all(r == 0 for r in rem) and len(z) == m - 1                  # Z covers all n constraint points (H5)
    circuit = {"NVars": m, "NPublic": 1}
    setup = groth16.GenerateTrustedSetup(m, circuit, alphas, betas, gammas)
    proof = groth16.GenerateProofs(circuit, setup["Pk"], w, px)
    assert o.groth16_verify(setup["Vk"], proof, [pub])
    assert not o.groth16_verify(setup["Vk"], proof, [(pub + 1) % R])
