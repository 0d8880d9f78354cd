"""Host mirror of the reference's ``r1csqap`` package for the prove path
(r1csqap/r1csqap.go): ``PolynomialField.Mul / Div / DivisorPolynomial`` run on
the GPU (NTT-based exact arithmetic over F_r) through libb200snark.
Polynomials are lists of Python ints, index = power of x, like ``[]*big.Int``.
"""
import numpy as np

from ._lib import check, ints_to_limbs, lib, limbs_to_ints, ptr  # noqa: F401
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

    def Add(self, a, b):                     # r1csqap.go:94-103
        return self._addsub(a, b, lib().b200_poly_add)

    def Sub(self, a, b):                     # r1csqap.go:106-115
        return self._addsub(a, b, lib().b200_poly_sub)

    def _addsub(self, a, b, fn):
        n = max(len(a), len(b))
        if n == 0:
            return []
        A = _coeffs(a) if len(a) else np.zeros((1, 4), dtype=np.uint64)
        B = _coeffs(b) if len(b) else np.zeros((1, 4), dtype=np.uint64)
        out = np.zeros((n, 4), dtype=np.uint64)
        check(fn(ptr(A), len(a), ptr(B), len(b), ptr(out)))
        return limbs_to_ints(out)

    def Eval(self, v, x):                    # r1csqap.go:118-126
        V = _coeffs(v) if len(v) else np.zeros((1, 4), dtype=np.uint64)
        X = ints_to_limbs([int(x) % R])
        out = np.zeros(4, dtype=np.uint64)
        check(lib().b200_poly_eval(ptr(V), len(v), ptr(X), ptr(out)))
        return limbs_to_ints(out)[0]

    def R1CSToQAP(self, a, b, c):            # r1csqap.go:161-188 -> (alphas, betas, gammas, z)
        n, m = len(a), len(a[0])
        mats = [ints_to_limbs([int(x) % R for row in M for x in row]) for M in (a, b, c)]
        outs = [np.zeros((m * n, 4), dtype=np.uint64) for _ in range(3)]
        z = np.zeros((m - 1, 4), dtype=np.uint64)
        check(lib().b200_r1cs_to_qap(ptr(mats[0]), ptr(mats[1]), ptr(mats[2]), n, m, ptr(outs[0]), ptr(outs[1]),
                                     ptr(outs[2]), ptr(z)))
        split = lambda o: [v[i * n:(i + 1) * n] for v in [limbs_to_ints(o)] for i in range(m)]
        return split(outs[0]), split(outs[1]), split(outs[2]), limbs_to_ints(z)

    def LagrangeInterpolation(self, v):      # r1csqap.go:150-158 (one column)
        col = [[x, 0] for x in v]            # n x 2 matrix: column 0 = v, column 1 = 0
        alphas, _, _, _ = self.R1CSToQAP(col, col, col)
        return alphas[0]

    def CombinePolynomials(self, r, ap, bp, cp):   # r1csqap.go:191-210 -> (ax, bx, cx, px)
        m, n = len(r), len(ap[0])
        R_ = ints_to_limbs([int(x) % R for x in r])
        mats = [ints_to_limbs([int(x) % R for row in M[:m] for x in row]) for M in (ap, bp, cp)]
        ax, bx, cx = (np.zeros((n, 4), dtype=np.uint64) for _ in range(3))
        px = np.zeros((2 * n - 1, 4), dtype=np.uint64)
        check(lib().b200_combine_polynomials(ptr(R_), m, ptr(mats[0]), ptr(mats[1]), ptr(mats[2]), n, ptr(ax), ptr(bx),
                                             ptr(cx), ptr(px)))
        return limbs_to_ints(ax), limbs_to_ints(bx), limbs_to_ints(cx), limbs_to_ints(px)


def Transpose(matrix):                        # r1csqap.go:11-21
    return [[matrix[j][i] for j in range(len(matrix))] for i in range(len(matrix[0]))]
