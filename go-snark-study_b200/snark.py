"""Host mirror of the reference's root ``snark`` package (Pinocchio) prove path
(snark.go:254-289) over libb200snark.

    pk = {"A","Ap","B","Bp","C","Cp","Kp","G1T","Z"}   (snark.Pk, snark.go:16-26; B is in G2)
    proof = GenerateProofs(circuit, pk, w, px) -> PiA PiAp PiB PiBp PiC PiCp PiH PiKp
"""
import numpy as np

from . import _lib
from ._lib import check, ints_to_limbs, lib, ptr
from .bn128 import R, _flatten_g1, _flatten_g2, _unflatten_g1, _unflatten_g2, reduce_scalar
from .groth16 import _attr

_ORDER = ("PiA", "PiAp", "PiBp", "PiC", "PiCp", "PiH", "PiKp")


class DeviceProvingKey:
    def __init__(self, pk, n_vars, n_public, window_bits=0):
        m = n_vars
        arr = {k: _flatten_g1(pk[k][:m]) for k in ("A", "Ap", "Bp", "C", "Cp", "Kp")}
        b2 = _flatten_g2(pk["B"][:m])
        g1t = _flatten_g1(pk["G1T"])
        z = ints_to_limbs([int(x) % R for x in pk["Z"]])
        h = _lib._h(0)
        check(lib().b200_pinocchio_pk_load(ptr(arr["A"]), ptr(arr["Ap"]), ptr(b2), ptr(arr["Bp"]), ptr(arr["C"]),
                                           ptr(arr["Cp"]), ptr(arr["Kp"]), m, ptr(g1t), len(pk["G1T"]), ptr(z),
                                           len(pk["Z"]), n_public, window_bits, h))
        self.handle = h.value

    def prove_limbs(self, w_limbs, px_limbs):
        g1 = np.zeros(84, dtype=np.uint64)
        pb = np.zeros(24, dtype=np.uint64)
        check(lib().b200_pinocchio_prove(self.handle, ptr(w_limbs), w_limbs.shape[0], ptr(px_limbs),
                                         px_limbs.shape[0], ptr(g1), ptr(pb)))
        proof = dict(zip(_ORDER, _unflatten_g1(g1)))
        proof["PiB"] = _unflatten_g2(pb)[0]
        return proof

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


def GenerateProofs(circuit, pk, w, px):
    """snark.GenerateProofs(circuit, pk, w, px) (snark.go:254). Deterministic."""
    dpk = pk if isinstance(pk, DeviceProvingKey) else LoadProvingKey(circuit, pk)
    wl = ints_to_limbs([reduce_scalar(x) for x in w])
    pl = ints_to_limbs([int(x) % R for x in px])
    return dpk.prove_limbs(wl, pl)
