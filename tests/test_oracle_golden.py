"""Pin the CPU oracle (oracle/ref_py.py) against every golden the reference
holds for the prove path and against outputs of the
reference's own Go binary (tests/golden/gobin_*.json, oracle/make_golden.py).
CPU only."""
import json
import os
import re

import pytest

from oracle import ref_py as o

G1, G2, PF = o.BN.G1, o.BN.G2, o.PF


def load(golden_dir, name):
    with open(os.path.join(golden_dir, name)) as f:
        return json.load(f)


def t3(p):
    return tuple(p)


def g2t(p):
    return tuple(tuple(c) for c in p)


# ---- K1: bn128/g1_test.go:11-31 ------------------------------------------------
def test_k1_g1_kat():
    g = G1.G
    g33, g44, g77 = G1.mul_scalar(g, 33), G1.mul_scalar(g, 44), G1.mul_scalar(g, 77)
    s = G1.affine(G1.add(g33, g44))
    assert s == G1.affine(g77)
    assert s[0] == 0x2f978c0ab89ebaa576866706b14787f360c4d6c3869efe5a72f7c3651a72ff00
    assert s[1] == 0x12e4ba7f0edca8b4fa668fe153aebd908d322dc26ad964d4cd314795844b62b2


def test_g2_add_property():           # bn128/g2_test.go:9-24
    g = G2.G
    assert G2.equal(G2.add(G2.mul_scalar(g, 33), G2.mul_scalar(g, 44)), G2.mul_scalar(g, 77))


# ---- fields/fqn_test.go:22-83: F7 literals -------------------------------------
def test_fq_f7_literals():
    f = o.Fq(7)
    assert f.add(4, 4) == 1 and f.double(5) == 3 and f.sub(2, 4) == 5
    assert f.neg(2) == 5 and f.mul(4, 4) == 2 and f.inverse(4) == 2 and f.square(5) == 4
    f2 = o.Fq2(f, 7 - 1)
    a, b = (4, 4), (3, 4)
    assert f2.add(a, b) == (0, 1)
    assert f2.mul(f2.div(a, b), b) == a
    assert f2.mul(f2.inverse(a), a) == (1, 0)


# ---- r1csqap/r1csqap_test.go:59-88 ---------------------------------------------
def test_poly_literals():
    pf = o.PolynomialField(o.Fq(o.R))
    a, b = [1, 0, 5], [3, 0, 1]
    c = pf.mul(a, b)
    assert c == [3, 0, 16, 0, 5]
    q, r = pf.div(c, a)
    assert q == b and all(x == 0 for x in r)
    assert pf.add(a, b) == [4, 0, 6]
    assert pf.sub(a, b) == [o.R - 2, 0, 4]


def _parse_go_matrices(stdout):
    """`fmt.Println` of [][]*big.Int -> list of matrices."""
    mats = []
    for line in stdout.splitlines():
        line = line.strip()
        if line.startswith("[[") and line.endswith("]]"):
            rows = re.findall(r"\[([0-9 ]*)\]", line[1:-1])
            mats.append([[int(x) for x in r.split()] for r in rows])
    return mats


@pytest.mark.parametrize("name", ["x3x5", "mul", "chain21"])
def test_qap_against_go_binary(golden_dir, name):
    """K3/K6/K7: R1CSToQAP + CombinePolynomials reproduce the Go binary's
    alphas/betas/gammas (stdout of `compile`, cli/main.go:143-147) and px.json."""
    g = load(golden_dir, f"gobin_{name}.json")
    r1cs = g["compiledcircuit"]["R1CS"]
    alphas, betas, gammas, z = PF.r1cs_to_qap(r1cs["A"], r1cs["B"], r1cs["C"])
    go_qap = _parse_go_matrices(g["compile_stdout"].split("qap", 1)[1])
    assert go_qap[0] == alphas and go_qap[1] == betas and go_qap[2] == gammas
    ax, bx, cx, px = PF.combine_polynomials(g["witness"], alphas, betas, gammas)
    assert px == g["px"]
    hx, rem = PF.div(px, z)
    assert all(x == 0 for x in rem)                      # groth16_test.go:77-83
    assert len(hx) == len(px) - len(z) + 1               # groth16_test.go:86
    assert PF.mul(hx, z) == px
    assert z == g["groth16_setup"]["Pk"]["Z"] == g["pinocchio_setup"]["Pk"]["Z"]


def test_k3_k4_wasm_literals(golden_dir):
    """K3 (wasm/index.js:8 first/last px coefficient) and K4 (Pk.Z of config 1)."""
    g = load(golden_dir, "gobin_x3x5.json")
    assert g["px"][-1] == 14598495318168266605115151979982065705759253455717291424254733984391562089814
    assert str(g["px"][0]).endswith("808491809")
    R = o.R
    assert g["groth16_setup"]["Pk"]["Z"] == [720, R - 1764, 1624, R - 735, 175, R - 21, 1]


def _pinocchio_pk(setup):
    pk = {k: [t3(p) for p in setup["Pk"][k]] for k in ("A", "C", "Kp", "Ap", "Bp", "Cp")}
    pk["B"] = [g2t(p) for p in setup["Pk"]["B"]]
    pk["Z"] = setup["Pk"]["Z"]
    pk["G1T"] = [t3(p) for p in setup["G1T"]]           # binary's (older) layout, SURVEY E2
    return pk


@pytest.mark.parametrize("name", ["x3x5", "mul", "chain21"])
def test_k5_pinocchio_prove_bit_exact_jacobian(golden_dir, name):
    """snark.GenerateProofs is deterministic: the oracle must reproduce the Go
    binary's proofs.json bit-for-bit, Jacobian X,Y,Z included."""
    g = load(golden_dir, f"gobin_{name}.json")
    cc = g["compiledcircuit"]
    proof, _ = o.pinocchio_prove(cc["NVars"], cc["NPublic"], _pinocchio_pk(g["pinocchio_setup"]),
                                 g["witness"], g["px"])
    ref = g["pinocchio_proofs"]
    for k in ("PiA", "PiAp", "PiBp", "PiC", "PiCp", "PiH", "PiKp"):
        assert proof[k] == t3(ref[k]), k
    assert proof["PiB"] == g2t(ref["PiB"])


def _groth_pk(setup):
    pk = setup["Pk"]
    return {"Z": pk["Z"], "BACDelta": [t3(p) for p in pk["BACDelta"]],
            "PowersTauDelta": [t3(p) for p in pk["PowersTauDelta"]],
            "G1": {"Alpha": t3(pk["G1"]["Alpha"]), "Beta": t3(pk["G1"]["Beta"]), "Delta": t3(pk["G1"]["Delta"]),
                   "At": [t3(p) for p in pk["G1"]["At"]], "BACGamma": [t3(p) for p in pk["G1"]["BACGamma"]]},
            "G2": {"Beta": g2t(pk["G2"]["Beta"]), "Delta": g2t(pk["G2"]["Delta"]),
                   "BACGamma": [g2t(p) for p in pk["G2"]["BACGamma"]]}}


def _groth_vk(setup):
    vk = setup["Vk"]
    return {"IC": [t3(p) for p in vk["IC"]], "G1": {"Alpha": t3(vk["G1"]["Alpha"])},
            "G2": {k: g2t(vk["G2"][k]) for k in ("Beta", "Gamma", "Delta")}}


@pytest.mark.slow
@pytest.mark.parametrize("name", ["x3x5"])
def test_groth16_verify_go_proof_and_own_proof(golden_dir, name):
    """The oracle's VerifyProof accepts the Go binary's (randomised) proof and
    the oracle's own proof under the Go binary's CRS; rejects a wrong public
    input (groth16_test.go:100,106)."""
    g = load(golden_dir, f"gobin_{name}.json")
    cc = g["compiledcircuit"]
    vk = _groth_vk(g["groth16_setup"])
    goproof = {"PiA": t3(g["groth16_proofs"]["PiA"]), "PiB": g2t(g["groth16_proofs"]["PiB"]),
               "PiC": t3(g["groth16_proofs"]["PiC"])}
    assert o.groth16_verify(vk, goproof, g["public"])
    proof, _ = o.groth16_prove(cc["NVars"], cc["NPublic"], _groth_pk(g["groth16_setup"]),
                               g["witness"], g["px"], r=0x1234567, s=0x7654321)
    assert o.groth16_verify(vk, proof, g["public"])
    assert not o.groth16_verify(vk, proof, [g["public"][0] - 1])


@pytest.mark.slow
def test_k8_pairing_golden(golden_dir):
    """e(vk_alfa_1, vk_beta_2) == vk_alfabeta_12 of the snarkjs fixture
    (externalVerif/circom-test/verification_key.json:62-91)."""
    c = load(golden_dir, "circom_groth16.json")
    vk = c["vk"]
    a1 = tuple(int(x) for x in vk["vk_alfa_1"])
    b2 = tuple(tuple(int(x) for x in c2) for c2 in vk["vk_beta_2"])
    e = o.BN.pairing(a1, b2)
    gold = tuple(tuple(tuple(int(x) for x in f2) for f2 in f6) for f6 in vk["vk_alfabeta_12"])
    assert e == gold


@pytest.mark.slow
def test_circom_proof_verifies(golden_dir):
    """externalVerif/circomVerifier_test.go:9-13 — the snarkjs proof verifies."""
    c = load(golden_dir, "circom_groth16.json")
    vk, pr = c["vk"], c["proof"]
    i3 = lambda p: tuple(int(x) for x in p)
    i32 = lambda p: tuple(tuple(int(x) for x in q) for q in p)
    ovk = {"IC": [i3(p) for p in vk["IC"]], "G1": {"Alpha": i3(vk["vk_alfa_1"])},
           "G2": {"Beta": i32(vk["vk_beta_2"]), "Gamma": i32(vk["vk_gamma_2"]), "Delta": i32(vk["vk_delta_2"])}}
    proof = {"PiA": i3(pr["pi_a"]), "PiB": i32(pr["pi_b"]), "PiC": i3(pr["pi_c"])}
    assert o.groth16_verify(ovk, proof, [int(x) for x in c["public"]])


def test_groth16_setup_matches_formulas():
    """groth16_setup (toxic injected) is self-consistent: a proof made under it
    satisfies the exponent identity A*B = alpha*beta + IC*gamma + C*delta
    (checked in F_r with known discrete logs; no pairing needed)."""
    a = [[0, 0, 1, 0, 0, 0], [0, 0, 0, 0, 1, 0], [0, 1, 0, 0, 0, 0], [1, 0, 0, 0, 0, 0]]
    b = [[0, 0, 0, 1, 0, 0], [1, 0, 0, 0, 0, 0], [1, 0, 0, 0, 0, 0], [1, 0, 0, 0, 0, 0]]
    c = [[0, 0, 0, 0, 1, 0], [0, 1, 0, 0, 0, 0], [0, 0, 0, 0, 1, 0], [0, 0, 0, 0, 0, 1]]
    w = [1, 33, 3, 11, 33, 1]
    alphas, betas, gammas, z = PF.r1cs_to_qap(a, b, c)
    tox = {"T": 123456789, "Kalpha": 1111, "Kbeta": 2222, "Kgamma": 3333, "Kdelta": 4444}
    pk, vk = o.groth16_setup(6, 1, alphas, betas, gammas, tox)
    assert pk["Z"] == z
    _, _, _, px = PF.combine_polynomials(w, alphas, betas, gammas)
    r, s = 97, 89
    proof, raw = o.groth16_prove(6, 1, pk, w, px, r, s)
    F = o.FQR
    ev = lambda polys: sum(F.mul(wi, PF.eval(p, tox["T"])) for wi, p in zip(w, polys)) % o.R
    A = (ev(alphas) + tox["Kalpha"] + r * tox["Kdelta"]) % o.R
    B = (ev(betas) + tox["Kbeta"] + s * tox["Kdelta"]) % o.R
    assert G1.affine(proof["PiA"]) == G1.affine(G1.mul_scalar(G1.G, A))
    assert G2.affine(proof["PiB"]) == G2.affine(G2.mul_scalar(G2.G, B))
    # C from the verification equation: A*B = alpha*beta + ic*gamma + C*delta
    ic = sum(F.mul(w[i], F.mul(F.inverse(tox["Kgamma"]),
             (PF.eval(alphas[i], tox["T"]) * tox["Kbeta"] + PF.eval(betas[i], tox["T"]) * tox["Kalpha"]
              + PF.eval(gammas[i], tox["T"])) % o.R)) for i in range(2)) % o.R
    C = F.mul((A * B - tox["Kalpha"] * tox["Kbeta"] - ic * tox["Kgamma"]) % o.R, F.inverse(tox["Kdelta"]))
    assert G1.affine(proof["PiC"]) == G1.affine(G1.mul_scalar(G1.G, C))
