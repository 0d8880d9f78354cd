"""Host mirror of the reference's ``groth16`` package prove path
(groth16/groth16.go:225-278) over libb200snark.

    pk  = {"G1": {"Alpha","Beta","Delta","At","BACGamma"}, "G2": {"Beta","Delta","BACGamma"},
           "BACDelta", "PowersTauDelta", "Z"}            (field names of groth16.Pk, groth16.go:15-32)
    proof = GenerateProofs(circuit, pk, w, px)           -> {"PiA","PiB","PiC"} Jacobian tuples

``circuit`` only needs ``NVars`` and ``NPublic`` (all the reference reads from it).
The proving key is uploaded once (``LoadProvingKey``) and cached per pk object.
"""
import secrets

import numpy as np

from . import _lib
from ._lib import check, ints_to_limbs, lib, ptr
from .bn128 import R, _flatten_g1, _flatten_g2, _unflatten_g1, _unflatten_g2, reduce_scalar


def _attr(obj, name):
    return obj[name] if isinstance(obj, dict) else getattr(obj, name)


def rand_fr():
    """Fq.Rand over r (fields/fq.go:116-132): bitlen/8-1 = 30 random bytes, mod r."""
    return int.from_bytes(secrets.token_bytes(30), "big") % R


class DeviceProvingKey:
    def __init__(self, pk, n_vars, n_public, window_bits=0):
        m = n_vars
        g1, g2 = pk["G1"], pk["G2"]
        assert len(g1["At"]) >= m and len(g1["BACGamma"]) >= m and len(g2["BACGamma"]) >= m and len(pk["BACDelta"]) >= m
        at = _flatten_g1(g1["At"][:m])
        b1 = _flatten_g1(g1["BACGamma"][:m])
        b2 = _flatten_g2(g2["BACGamma"][:m])
        cd = _flatten_g1(pk["BACDelta"][:m])
        ptd = _flatten_g1(pk["PowersTauDelta"])
        z = ints_to_limbs([int(x) % R for x in pk["Z"]])
        a1, be1, d1 = (_flatten_g1([g1[k]]) for k in ("Alpha", "Beta", "Delta"))
        be2, d2 = (_flatten_g2([g2[k]]) for k in ("Beta", "Delta"))
        h = _lib._h(0)
        check(lib().b200_groth16_pk_load(ptr(at), ptr(b1), ptr(b2), ptr(cd), m, ptr(ptd), len(pk["PowersTauDelta"]),
                                         ptr(z), len(pk["Z"]), ptr(a1), ptr(be1), ptr(d1), ptr(be2), ptr(d2),
                                         n_public, window_bits, h))
        self.handle = h.value
        self.m = m

    def prove_limbs(self, w_limbs, px_limbs, r, s):
        pa = np.zeros(12, dtype=np.uint64)
        pb = np.zeros(24, dtype=np.uint64)
        pc = np.zeros(12, dtype=np.uint64)
        rr, ss = ints_to_limbs([r]), ints_to_limbs([s])
        check(lib().b200_groth16_prove(self.handle, ptr(w_limbs), w_limbs.shape[0], ptr(px_limbs), px_limbs.shape[0],
                                       ptr(rr), ptr(ss), ptr(pa), ptr(pb), ptr(pc)))
        return {"PiA": _unflatten_g1(pa)[0], "PiB": _unflatten_g2(pb)[0], "PiC": _unflatten_g1(pc)[0]}

    def free(self):
        if self.handle:
            check(lib().b200_pk_free(self.handle))
            self.handle = 0


_pk_cache = {}


def LoadProvingKey(circuit, pk, window_bits=0):
    key = id(pk)
    if key not in _pk_cache:
        _pk_cache[key] = DeviceProvingKey(pk, _attr(circuit, "NVars"), _attr(circuit, "NPublic"), window_bits)
    return _pk_cache[key]


def GenerateProofs(circuit, pk, w, px, r=None, s=None):
    """groth16.GenerateProofs(circuit, pk, w, px) (groth16.go:225).  r, s default
    to fresh randomness like the reference; pass them for reproducible proofs."""
    dpk = pk if isinstance(pk, DeviceProvingKey) else LoadProvingKey(circuit, pk)
    r = rand_fr() if r is None else int(r) % R
    s = rand_fr() if s is None else int(s) % R
    wl = ints_to_limbs([reduce_scalar(x) for x in w])
    pl = ints_to_limbs([int(x) % R for x in px])
    return dpk.prove_limbs(wl, pl, r, s)
