"""File-compatible stand-in for the prove commands of the reference's CLI (cli/main.go):

    python -m gosnark_b200.cli groth16 genproofs      # cli/main.go:455-518
    python -m gosnark_b200.cli groth16 verify         # cli/main.go:520-549
    python -m gosnark_b200.cli groth16 trustedsetup   # cli/main.go:407-453
    python -m gosnark_b200.cli trustedsetup [wasm]    # cli/main.go:231-301   (Pinocchio)
    python -m gosnark_b200.cli verify                 # cli/main.go:368-396   (Pinocchio)
    python -m gosnark_b200.cli genproofs              # cli/main.go:303-366   (Pinocchio)

Reads the files the Go CLI writes/reads in the current directory (compiledcircuit.json,
trustedsetup.json, privateInputs.json, publicInputs.json — Go encoding/json of the structs, big.Int as
JSON numbers, Jacobian coordinates), runs witness -> R1CSToQAP -> CombinePolynomials -> GenerateProofs on
the GPU and writes proofs.json in the same layout, so the unmodified `go-snark-cli [groth16] verify`
accepts it.  Host glue only; no kernel logic here.
"""
import json
import sys

from . import _lib, groth16, r1csqap, snark


def calculate_witness(circuit, private_inputs, public_inputs):
    """circuitcompiler.Circuit.CalculateWitness (circuitcompiler/circuit.go:158-186) over the compiled
    circuit's flat constraint list (unreduced integers, as in the reference)."""
    if len(private_inputs) != len(circuit["PrivateInputs"]) or len(public_inputs) != len(circuit["PublicInputs"]):
        raise ValueError("given inputs != circuit inputs")
    signals = circuit["Signals"]
    w = [0] * len(signals)
    w[0] = 1
    for i, v in enumerate(public_inputs):
        w[i + 1] = int(v)
    for i, v in enumerate(private_inputs):
        w[i + len(public_inputs) + 1] = int(v)

    def grab(s):                                   # grabVar, circuit.go:141-149
        try:
            return int(s)
        except ValueError:
            return w[signals.index(s)]

    for c in circuit["Constraints"]:
        op = c["Op"]
        if op == "in":
            continue
        a, b = grab(c["V1"]), grab(c["V2"])
        out = signals.index(c["Out"])
        if op == "+":
            w[out] = a + b
        elif op == "-":
            w[out] = a - b
        elif op == "*":
            w[out] = a * b
        elif op == "/":
            if b == 0:                                 # big.Int.Div panics "division by zero" (circuit.go:182)
                raise ZeroDivisionError("calculate_witness: division by zero in constraint %r" % (c["Out"],))
            w[out] = a // b if b > 0 else -(a // -b)   # big.Int.Div is Euclidean: remainder in [0, |b|)
    return w


def _t3(p):
    return tuple(p)


def _g2(p):
    return tuple(tuple(c) for c in p)


def _load(name):
    with open(name) as f:
        return json.load(f)


def _qap_px(circuit, w):
    pf = r1csqap.PolynomialField()
    r1cs = circuit["R1CS"]
    alphas, betas, gammas, _ = pf.R1CSToQAP(r1cs["A"], r1cs["B"], r1cs["C"])
    _, _, _, px = pf.CombinePolynomials(w, alphas, betas, gammas)
    return px


def groth16_genproofs():
    circuit, setup = _load("compiledcircuit.json"), _load("trustedsetup.json")
    w = calculate_witness(circuit, _load("privateInputs.json"), _load("publicInputs.json"))
    px = _qap_px(circuit, w)
    pk = setup["Pk"]
    pkd = {"Z": pk["Z"], "BACDelta": [_t3(p) for p in pk["BACDelta"]], "PowersTauDelta": [_t3(p) for p in pk["PowersTauDelta"]],
           "G1": {"Alpha": _t3(pk["G1"]["Alpha"]), "Beta": _t3(pk["G1"]["Beta"]), "Delta": _t3(pk["G1"]["Delta"]),
                  "At": [_t3(p) for p in pk["G1"]["At"]], "BACGamma": [_t3(p) for p in pk["G1"]["BACGamma"]]},
           "G2": {"Beta": _g2(pk["G2"]["Beta"]), "Delta": _g2(pk["G2"]["Delta"]),
                  "BACGamma": [_g2(p) for p in pk["G2"]["BACGamma"]]}}
    proof = groth16.GenerateProofs(circuit, pkd, w, px)
    out = {"PiA": list(proof["PiA"]), "PiB": [list(c) for c in proof["PiB"]], "PiC": list(proof["PiC"])}
    with open("proofs.json", "w") as f:
        json.dump(out, f)
    print("witness", w)
    print("Proofs data written to  proofs.json")


def pinocchio_genproofs():
    circuit, setup = _load("compiledcircuit.json"), _load("trustedsetup.json")
    w = calculate_witness(circuit, _load("privateInputs.json"), _load("publicInputs.json"))
    px = _qap_px(circuit, w)
    pk = {k: [_t3(p) for p in setup["Pk"][k]] for k in ("A", "C", "Kp", "Ap", "Bp", "Cp")}
    pk["B"] = [_g2(p) for p in setup["Pk"]["B"]]
    pk["Z"] = setup["Pk"]["Z"]
    # the prebuilt CLI is one commit older than snark.go: G1T sits at the top level of the setup
    pk["G1T"] = [_t3(p) for p in (setup["Pk"].get("G1T") or setup["G1T"])]
    proof = snark.GenerateProofs(circuit, pk, w, px)
    out = {k: (list(v) if k != "PiB" else [list(c) for c in v]) for k, v in proof.items()}
    with open("proofs.json", "w") as f:
        json.dump(out, f)
    print("witness", w)
    print("Proofs data written to  proofs.json")


def groth16_verify():
    """cli/main.go:520-549 — the proof of proofs.json against trustedsetup.json's Vk and publicInputs.json, on the GPU."""
    proof, setup, public = _load("proofs.json"), _load("trustedsetup.json"), _load("publicInputs.json")
    vk = setup["Vk"]
    vkd = {"IC": [_t3(p) for p in vk["IC"]], "G1": {"Alpha": _t3(vk["G1"]["Alpha"])},
           "G2": {k: _g2(vk["G2"][k]) for k in ("Beta", "Gamma", "Delta")}}
    pr = {"PiA": _t3(proof["PiA"]), "PiB": _g2(proof["PiB"]), "PiC": _t3(proof["PiC"])}
    verified = groth16.VerifyProof(vkd, pr, [int(x) for x in public], True)
    print("Proofs verified" if verified else "ERROR: proofs not verified")
    return verified


def _qap(circuit):
    pf = r1csqap.PolynomialField()
    r1cs = circuit["R1CS"]
    return pf.R1CSToQAP(r1cs["A"], r1cs["B"], r1cs["C"])


def _jl(p):
    """Jacobian point -> the nested JSON lists Go's encoding/json writes for [3]*big.Int / [3][2]*big.Int."""
    return [list(c) if isinstance(c, (tuple, list)) else c for c in p]


def _write_setup(obj, wasm_obj=None):
    with open("trustedsetup.json", "w") as f:
        json.dump(obj, f)
    print("Trusted Setup data written to  trustedsetup.json")
    if wasm_obj is not None:
        with open("trustedsetupString.json", "w") as f:
            json.dump(wasm_obj, f)


def groth16_trustedsetup():
    """cli/main.go:407-453: compiledcircuit.json + inputs -> trustedsetup.json (Toxic erased, like the reference)."""
    circuit = _load("compiledcircuit.json")
    w = calculate_witness(circuit, _load("privateInputs.json"), _load("publicInputs.json"))
    alphas, betas, gammas, _ = _qap(circuit)
    setup = groth16.GenerateTrustedSetup(len(w), circuit, alphas, betas, gammas)
    pk, vk = setup["Pk"], setup["Vk"]
    out = {"Toxic": {k: None for k in ("T", "Kalpha", "Kbeta", "Kgamma", "Kdelta")},
           "Pk": {"BACDelta": [_jl(p) for p in pk["BACDelta"]], "Z": pk["Z"],
                  "G1": {"Alpha": _jl(pk["G1"]["Alpha"]), "Beta": _jl(pk["G1"]["Beta"]), "Delta": _jl(pk["G1"]["Delta"]),
                         "At": [_jl(p) for p in pk["G1"]["At"]], "BACGamma": [_jl(p) for p in pk["G1"]["BACGamma"]]},
                  "G2": {"Beta": _jl(pk["G2"]["Beta"]), "Gamma": _jl(pk["G2"]["Gamma"]), "Delta": _jl(pk["G2"]["Delta"]),
                         "BACGamma": [_jl(p) for p in pk["G2"]["BACGamma"]]},
                  "PowersTauDelta": [_jl(p) for p in pk["PowersTauDelta"]]},
           "Vk": {"IC": [_jl(p) for p in vk["IC"]], "G1": {"Alpha": _jl(vk["G1"]["Alpha"])},
                  "G2": {k: _jl(vk["G2"][k]) for k in ("Beta", "Gamma", "Delta")}}}
    _write_setup(out)


def pinocchio_trustedsetup(wasm=False):
    """cli/main.go:231-301 (`trustedsetup [wasm]`).  G1T is written both inside Pk (snark.go:16-26) and at the top level
    (the layout of the prebuilt binary, one commit older — SURVEY E2); Go's json.Unmarshal ignores the one it does not know."""
    circuit = _load("compiledcircuit.json")
    w = calculate_witness(circuit, _load("privateInputs.json"), _load("publicInputs.json"))
    alphas, betas, gammas, _ = _qap(circuit)
    setup = snark.GenerateTrustedSetup(len(w), circuit, alphas, betas, gammas)
    pk, vk = setup["Pk"], setup["Vk"]
    jpk = {k: ([_jl(p) for p in v] if k != "Z" else v) for k, v in pk.items()}
    jvk = {k: ([_jl(p) for p in v] if k == "IC" else _jl(v)) for k, v in vk.items()}
    out = {"Toxic": {k: None for k in ("T", "Ka", "Kb", "Kc", "Kbeta", "Kgamma", "RhoA", "RhoB", "RhoC")},
           "G1T": jpk["G1T"], "G2T": None, "Pk": jpk, "Vk": jvk}
    from . import utils
    _write_setup(out, utils.SetupToString(setup) if wasm else None)


def pinocchio_verify():
    """cli/main.go:368-396 — proofs.json against trustedsetup.json's Vk and publicInputs.json, pairings on the GPU."""
    proof, setup, public = _load("proofs.json"), _load("trustedsetup.json"), _load("publicInputs.json")
    vk = {k: ([_t3(p) for p in v] if k == "IC" else (_g2(v) if isinstance(v[0], list) else _t3(v))) for k, v in setup["Vk"].items()}
    pr = {k: (_g2(v) if k == "PiB" else _t3(v)) for k, v in proof.items()}
    verified = snark.VerifyProof(vk, pr, [int(x) for x in public], True)
    print("Proofs verified" if verified else "ERROR: proofs not verified")
    return verified


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    _lib.init()
    if argv[:2] == ["groth16", "genproofs"]:
        groth16_genproofs()
    elif argv[:2] == ["groth16", "verify"]:
        groth16_verify()
    elif argv[:2] == ["groth16", "trustedsetup"]:
        groth16_trustedsetup()
    elif argv[:1] == ["genproofs"]:
        pinocchio_genproofs()
    elif argv[:1] == ["verify"]:
        pinocchio_verify()
    elif argv[:1] == ["trustedsetup"]:
        pinocchio_trustedsetup(wasm=argv[1:2] == ["wasm"])
    else:
        print("usage: python -m gosnark_b200.cli [groth16] trustedsetup|genproofs|verify", file=sys.stderr)
        return 2
    return 0


if __name__ == "__main__":
    sys.exit(main())
