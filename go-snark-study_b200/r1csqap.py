"""Host mirror of the reference's ``r1csqap`` package for the prove path
(r1csqap/r1csqap.go): ``PolynomialField.Mul / Div / DivisorPolynomial`` run on
the GPU (NTT-based exact arithmetic over F_r) through libb200snark.
Polynomials are lists of Python ints, index = power of x, like ``[]*big.Int``.
"""
import numpy as np

from ._lib import check, ints_to_limbs, lib, limbs_to_ints, ptr
from .bn128 import R


def _coeffs(v):
    """Coefficients as the cgo shim sends them: canonical residues mod r (the
    reference reduces on first use: Fq.Mul / Fq.Add apply Mod, fields/fq.go:32-59)."""
    return ints_to_limbs([int(x) % R for x in v])


class PolynomialField:
    """r1csqap.PolynomialField over F_r (r1csqap.go:45-55)."""

    def Mul(self, a, b):                     # r1csqap.go:57-67
        if len(a) == 0 or len(b) == 0:
            return []
        A, B = _coeffs(a), _coeffs(b)
        out = np.zeros((len(a) + len(b) - 1, 4), dtype=np.uint64)
        check(lib().b200_poly_mul(ptr(A), len(a), ptr(B), len(b), ptr(out)))
        return limbs_to_ints(out)

    def Div(self, a, b):                     # r1csqap.go:70-84 -> (quotient, remainder)
        A, B = _coeffs(a), _coeffs(b)
        if len(a) < len(b):
            return [], [int(x) % R for x in a]
        nq = len(a) - len(b) + 1
        q = np.zeros((nq, 4), dtype=np.uint64)
        rem = np.zeros((max(len(b) - 1, 1), 4), dtype=np.uint64)
        check(lib().b200_poly_div(ptr(A), len(a), ptr(B), len(b), ptr(q), ptr(rem)))
        return limbs_to_ints(q), (limbs_to_ints(rem) if len(b) > 1 else [])

    def DivisorPolynomial(self, px, z):      # r1csqap.go:213-216
        A, B = _coeffs(px), _coeffs(z)
        if len(px) < len(z):
            return []
        q = np.zeros((len(px) - len(z) + 1, 4), dtype=np.uint64)
        check(lib().b200_poly_div(ptr(A), len(px), ptr(B), len(z), ptr(q), None))
        return limbs_to_ints(q)
