"""CPU: the pairing header (csrc/pairing.cuh) compiled for the host through the carry-flag emulation must
reproduce bn128.Pairing bit-for-bit: against the oracle (pinned to the reference by K8) and directly against
the snarkjs golden vk_alfabeta_12 (externalVerif/circom-test/verification_key.json:62-91)."""
import ctypes
import json
import os

import numpy as np
import pytest

import build as b200build
from oracle import ref_py as o

pytestmark = pytest.mark.slow


@pytest.fixture(scope="module")
def lib():
    return ctypes.CDLL(b200build.build_host_arith())


def to_u32(vals):
    buf = b"".join(int(v).to_bytes(32, "little") for v in vals)
    return np.frombuffer(buf, dtype=np.uint32).copy()


def pairing(lib, p1_aff, p2_aff, fast=False):
    g1 = to_u32([p1_aff[0], p1_aff[1]])
    g2 = to_u32([p2_aff[0][0], p2_aff[0][1], p2_aff[1][0], p2_aff[1][1]])
    out = np.zeros(96, dtype=np.uint32)
    (lib.t_pairing_fast if fast else lib.t_pairing)(g1.ctypes.data_as(ctypes.c_void_p), g2.ctypes.data_as(ctypes.c_void_p), out.ctypes.data_as(ctypes.c_void_p))
    raw = out.tobytes()
    v = [int.from_bytes(raw[32 * i:32 * (i + 1)], "little") for i in range(12)]
    return tuple(tuple((v[6 * h + 2 * k], v[6 * h + 2 * k + 1]) for k in range(3)) for h in range(2))


def test_k8_golden(lib, golden_dir):
    vk = json.load(open(os.path.join(golden_dir, "circom_groth16.json")))["vk"]
    a1 = tuple(int(x) for x in vk["vk_alfa_1"])
    b2 = tuple(tuple(int(x) for x in c) for c in vk["vk_beta_2"])
    gold = tuple(tuple(tuple(int(x) for x in f2) for f2 in f6) for f6 in vk["vk_alfabeta_12"])
    assert pairing(lib, a1[:2], b2[:2]) == gold


def test_vs_oracle_and_bilinearity(lib):
    G1, G2 = o.BN.G1, o.BN.G2
    p = G1.affine(G1.mul_scalar(G1.G, 25))
    q = G2.affine(G2.mul_scalar(G2.G, 30))
    e1 = pairing(lib, p, q[:2])
    assert e1 == o.BN.pairing((p[0], p[1], 1), (q[0], q[1], (1, 0)))
    assert e1[0][0][0] == 8016119724813186033542830391460394070015218389456422587891475873290878009957   # bn128_test.go:66
    p2 = G1.affine(G1.mul_scalar(G1.G, 30))
    q2 = G2.affine(G2.mul_scalar(G2.G, 25))
    assert pairing(lib, p2, q2[:2]) == e1                                                               # bn128_test.go:45-67


def _f12_flat(x):
    return [c for h in x for f2 in h for c in f2]


def _f12_op(lib, op, x):
    buf = to_u32(_f12_flat(x))
    out = np.zeros(96, dtype=np.uint32)
    lib.t_f12_op(op, buf.ctypes.data_as(ctypes.c_void_p), out.ctypes.data_as(ctypes.c_void_p))
    raw = out.tobytes()
    v = [int.from_bytes(raw[32 * i:32 * (i + 1)], "little") for i in range(12)]
    return tuple(tuple((v[6 * h + 2 * k], v[6 * h + 2 * k + 1]) for k in range(3)) for h in range(2))


def test_fast_final_exponentiation_is_the_same_field_element(lib, golden_dir):
    """final_exp_fast (easy part + Devegili-Scott-Dahab hard part) == the reference's plain f^((q^12-1)/r): on the snarkjs
    golden, the bn128_test.go:66 literal, and G1 infinity; its building blocks (F_q^12 inverse, Frobenius, x^u) against the
    oracle's plain exponentiations."""
    import random
    vk = json.load(open(os.path.join(golden_dir, "circom_groth16.json")))["vk"]
    a1 = tuple(int(x) for x in vk["vk_alfa_1"])
    b2 = tuple(tuple(int(x) for x in c) for c in vk["vk_beta_2"])
    gold = tuple(tuple(tuple(int(x) for x in f2) for f2 in f6) for f6 in vk["vk_alfabeta_12"])
    assert pairing(lib, a1[:2], b2[:2], fast=True) == gold
    G1, G2 = o.BN.G1, o.BN.G2
    p = G1.affine(G1.mul_scalar(G1.G, 25))
    q = G2.affine(G2.mul_scalar(G2.G, 30))
    e = pairing(lib, p, q[:2], fast=True)
    assert e == pairing(lib, p, q[:2]) and e[0][0][0] == 8016119724813186033542830391460394070015218389456422587891475873290878009957
    assert pairing(lib, (0, 0), q[:2], fast=True) == pairing(lib, (0, 0), q[:2])
    rng = random.Random(6)
    x = tuple(tuple((rng.randrange(o.Q), rng.randrange(o.Q)) for _ in range(3)) for _ in range(2))
    F12 = o.BN.Fq12
    assert _f12_op(lib, 0, x) == F12.inverse(x)
    for k in (1, 2, 3):
        assert _f12_op(lib, k, x) == F12.exp(x, o.Q ** k)
    assert _f12_op(lib, 4, x) == F12.exp(x, 4965661367192848881)
