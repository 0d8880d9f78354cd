"""GPU: the format / tooling rows end to end on the device — the same mirrors that
tests/test_snark_host_logic.py and tests/test_utils_parsers.py exercise on the CPU, here over the real library."""
import json
import os
import shutil
import subprocess
import tempfile

import pytest

from oracle import ref_py as o
from test_utils_parsers import K5, K5_PIB

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOBIN = os.path.join(ROOT, "oracle", "_ref", "go-snark-cli")
G1, G2 = o.BN.G1, o.BN.G2


@pytest.fixture(scope="module", autouse=True)
def _init():
    from gosnark_b200 import _lib
    _lib.init()


def test_wasm_string_fixture_gives_the_k5_proof(golden_dir):
    """wasm/index.js:2-8 string-form circuit / setup / px -> snark.GenerateProofs on the GPU == the Go binary's proof."""
    from gosnark_b200 import snark, utils
    wasm = json.load(open(os.path.join(golden_dir, "wasm_index_strings.json")))
    circuit = utils.CircuitFromString(wasm["circuit"])
    setup = utils.SetupFromString(wasm["setup"])
    px = utils.ArrayStringToBigInt(wasm["px"])
    proof = snark.GenerateProofs(circuit, setup["Pk"], [1, 35, 3, 9, 27, 30, 35, 1], px)
    for k, v in K5.items():
        assert G1.affine(proof[k])[:2] == v, k
    assert G2.affine(proof["PiB"])[:2] == K5_PIB
    s = utils.ProofToString(proof)
    assert utils.ProofFromString(s) == proof and utils.ProofFromHex(utils.ProofToHex(proof)) == proof


def test_verify_from_circom_files(golden_dir, tmp_path):
    """externalVerif/circomVerifier_test.go:9-13."""
    from gosnark_b200 import externalVerif
    c = json.load(open(os.path.join(golden_dir, "circom_groth16.json")))
    paths = {}
    for name, obj in (("verification_key.json", c["vk"]), ("proof.json", c["proof"]), ("public.json", c["public"])):
        paths[name] = str(tmp_path / name)
        json.dump(obj, open(paths[name], "w"))
    ok, err = externalVerif.VerifyFromCircom(paths["verification_key.json"], paths["proof.json"], paths["public.json"])
    assert ok and err is None
    json.dump([str(int(c["public"][0]) + 1)], open(paths["public.json"], "w"))
    ok, err = externalVerif.VerifyFromCircom(paths["verification_key.json"], paths["proof.json"], paths["public.json"])
    assert not ok and err is None


@pytest.mark.parametrize("proto", ["groth16", "pinocchio"])
def test_cli_trustedsetup_prove_verify_with_go_in_the_loop(golden_dir, proto, capsys):
    """Our `trustedsetup` (GPU-minted CRS) -> Go `genproofs` + `verify` accept it -> our `genproofs` overwrites
    proofs.json -> Go `verify` and our `verify` accept that too (cli/main.go:231-549)."""
    if not os.path.exists(GOBIN):
        pytest.skip("oracle/_ref/go-snark-cli not staged")
    from gosnark_b200 import cli
    g = json.load(open(os.path.join(golden_dir, "gobin_x3x5.json")))
    d = tempfile.mkdtemp(prefix="clizz_")
    cwd = os.getcwd()
    pre = ["groth16"] if proto == "groth16" else []
    ok_text = (lambda out: "verification passed" in out) if proto == "groth16" else \
        (lambda out: "Proofs verified" in out and "❌" not in out)
    try:
        for fname, key in (("compiledcircuit.json", "compiledcircuit"), ("privateInputs.json", "private"),
                           ("publicInputs.json", "public")):
            json.dump(g[key], open(os.path.join(d, fname), "w"))
        os.chdir(d)
        assert cli.main(pre + ["trustedsetup"]) == 0
        b = os.path.join(d, "gsc")
        shutil.copy(GOBIN, b)
        os.chmod(b, 0o755)
        run = lambda *a: subprocess.run([b, *pre, *a], cwd=d, capture_output=True, text=True, timeout=120)
        p = run("genproofs")
        assert os.path.exists("proofs.json"), p.stdout[-400:] + p.stderr[-400:]
        p = run("verify")
        assert ok_text(p.stdout + p.stderr), p.stdout + p.stderr
        capsys.readouterr()
        assert cli.main(pre + ["verify"]) == 0                       # Go's proof, our verifier
        assert "Proofs verified" in capsys.readouterr().out
        assert cli.main(pre + ["genproofs"]) == 0                    # our proof under our setup
        p = run("verify")
        assert ok_text(p.stdout + p.stderr), p.stdout + p.stderr
        capsys.readouterr()
        assert cli.main(pre + ["verify"]) == 0
        assert "Proofs verified" in capsys.readouterr().out
    finally:
        os.chdir(cwd)
        shutil.rmtree(d)


@pytest.mark.timeout(900)
@pytest.mark.parametrize("group,logn", [(1, 20), (2, 22)])
def test_msm_full_size_known_discrete_logs(group, logn):
    """BASELINE configs 3 and 5 (G1 MSM N = 2^20, G2 MSM N = 2^22) at full size: P_i = k_i*G minted on the GPU, so
    sum s_i P_i = (sum s_i k_i mod r)*G — one CPU scalar multiplication gives the exact expected point.
    Full-width scalars with 1 % zeros and a block of small (witness-like) values; linearity: msm(s) + msm(s') = msm(s+s')."""
    import numpy as np
    from gosnark_b200 import _lib, bn128
    n = 1 << logn
    G = G1 if group == 1 else G2
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
    bs = bn128.BaseSet(group, limbs=pts)
    try:
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
