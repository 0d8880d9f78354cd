"""Host mirror of the reference's ``externalVerif`` package: verification of snarkjs / circom Groth16 proofs
(externalVerif/circomVerifier.go:12-96).  The JSON files are parsed through ``utils`` exactly as the
reference does and the proof is verified by ``groth16.VerifyProof`` on the GPU."""
import json

from . import groth16, utils


def VerifyFromCircom(vkPath, proofPath, publicSignalsPath):
    """(verified, err) = VerifyFromCircom(vkPath, proofPath, publicSignalsPath)  (circomVerifier.go:26-96).
    ``err`` is None or the exception a missing / malformed file produced (the Go function returns it likewise)."""
    try:
        with open(vkPath) as f:
            circom_vk = json.load(f)
        str_vk = {"IC": circom_vk["IC"], "G1": {"Alpha": circom_vk["vk_alfa_1"]},                    # :37-42
                  "G2": {"Beta": circom_vk["vk_beta_2"], "Gamma": circom_vk["vk_gamma_2"], "Delta": circom_vk["vk_delta_2"]}}
        vk = utils.GrothVkFromString(str_vk)
        print("vk parsed:", vk)
        with open(proofPath) as f:
            circom_proof = json.load(f)
        proof = utils.GrothProofFromString({"PiA": circom_proof["pi_a"], "PiB": circom_proof["pi_b"],   # :59-63
                                            "PiC": circom_proof["pi_c"]})
        print("proof parsed:", proof)
        with open(publicSignalsPath) as f:
            public_signals = utils.ArrayStringToBigInt(json.load(f))
        print("publicSignals parsed:", public_signals)
    except (OSError, ValueError, KeyError) as e:
        return False, e
    return groth16.VerifyProof(vk, proof, public_signals, True), None
