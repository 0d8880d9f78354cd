"""Regression tests for the implicit proving-key cache of the drop-in GenerateProofs(circuit, pk, w, px) form
(groth16.go:225 / snark.go:254 take the key by value on every call).  Round 1 keyed the cache on id(pk): after the
first key dict was freed CPython handed its address to the next one and the proof came out under the STALE key."""
import gc
import json
import os

import pytest

from gosnark_b200.groth16 import KeyCache


class FakeDeviceKey:
    live = 0

    def __init__(self, pk, m, npub, c):
        self.pk_copy, self.shape, self.freed = dict(pk), (m, npub, c), False
        FakeDeviceKey.live += 1

    def free(self):
        assert not self.freed
        self.freed = True
        FakeDeviceKey.live -= 1


def test_cache_matches_by_identity_not_by_address():
    FakeDeviceKey.live = 0
    cache = KeyCache(FakeDeviceKey, capacity=2)
    circ = {"NVars": 8, "NPublic": 1}
    seen_ids = set()
    for k in range(50):                      # freed dicts get their addresses recycled: ids repeat, keys must not
        pk = {"tag": k}
        seen_ids.add(id(pk))
        dpk = cache.get(circ, pk)
        assert dpk.pk_copy["tag"] == k
        assert cache.get(circ, pk) is dpk    # same object, same shape: hit
        del pk
        gc.collect()
    assert FakeDeviceKey.live <= 2           # evicted keys were freed
    cache.forget()
    assert FakeDeviceKey.live == 0


def test_cache_respects_shape_and_window_bits():
    cache = KeyCache(FakeDeviceKey, capacity=4)
    pk = {"tag": 1}
    a = cache.get({"NVars": 8, "NPublic": 1}, pk)
    assert cache.get({"NVars": 8, "NPublic": 1}, pk, window_bits=12) is not a
    assert cache.get({"NVars": 8, "NPublic": 2}, pk) is not a
    assert cache.get({"NVars": 8, "NPublic": 1}, pk) is a
    cache.forget(pk)
    assert not cache.entries and a.freed


@pytest.mark.gpu
@pytest.mark.parametrize("proto", ["pinocchio", "groth16"])
def test_two_same_shape_keys_back_to_back(golden_dir, proto):
    """Two different keys of the same shape, the first one freed (gc.collect) before the second is built: each proof
    must equal the oracle's under ITS key, and the two must differ."""
    from oracle import ref_py as o
    from gosnark_b200 import _lib, groth16, snark
    _lib.init()
    with open(os.path.join(golden_dir, "gobin_x3x5.json")) as f:
        g = json.load(f)
    cc, w, px = g["compiledcircuit"], g["witness"], g["px"]
    circ = {"NVars": cc["NVars"], "NPublic": cc["NPublic"]}
    alphas, betas, gammas, _ = o.PF.r1cs_to_qap(cc["R1CS"]["A"], cc["R1CS"]["B"], cc["R1CS"]["C"])
    proofs = []
    for seed in (11, 12, 13):
        if proto == "pinocchio":
            tox = {k: 1000 * seed + j for j, k in enumerate(("T", "Ka", "Kb", "Kc", "Kbeta", "Kgamma", "RhoA", "RhoB"), 7)}
            pk, _ = o.pinocchio_setup(cc["NVars"], cc["NPublic"], alphas, betas, gammas, tox)
            ours = snark.GenerateProofs(circ, pk, w, px)
            ref, _ = o.pinocchio_prove(cc["NVars"], cc["NPublic"], pk, w, px)
            for k in ref:
                G = o.BN.G2 if k == "PiB" else o.BN.G1
                assert G.affine(ours[k]) == G.affine(ref[k]), (seed, k)
            proofs.append(o.BN.G1.affine(ours["PiA"]))
        else:
            tox = {k: 1000 * seed + j for j, k in enumerate(("T", "Kalpha", "Kbeta", "Kgamma", "Kdelta"), 7)}
            pk, _ = o.groth16_setup(cc["NVars"], cc["NPublic"], alphas, betas, gammas, tox)
            ours = groth16.GenerateProofs(circ, pk, w, px, r=5, s=9)
            ref, _ = o.groth16_prove(cc["NVars"], cc["NPublic"], pk, w, px, 5, 9)
            for k, G in (("PiA", o.BN.G1), ("PiB", o.BN.G2), ("PiC", o.BN.G1)):
                assert G.affine(ours[k]) == G.affine(ref[k]), (seed, k)
            proofs.append(o.BN.G1.affine(ours["PiA"]))
        del pk
        gc.collect()
    assert len(set(proofs)) == 3
