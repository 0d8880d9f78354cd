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

    def LagrangeInterpolation(self, v):      # r1csqap.go:150-158: values at x = 1..n -> n coefficients, any n
        if len(v) == 0:
            return []
        V = _coeffs(v)
        out = np.zeros((len(v), 4), dtype=np.uint64)
        check(lib().b200_interpolate(ptr(V), len(v), ptr(out)))
        return limbs_to_ints(out)

    def NewPolZeroAt(self, pointPos, totalPoints, height):   # r1csqap.go:129-147 (exact, no native-int overflow)
        """height * prod_{i != pointPos} (x - i) / (pointPos - i) over i = 1..totalPoints."""
        v = [0] * totalPoints
        v[pointPos - 1] = int(height) % R
        return self.LagrangeInterpolation(v)

    def CombinePolynomials(self, r, ap, bp, cp):   # r1csqap.go:191-210 -> (ax, bx, cx, px)
        m, n = len(r), len(ap[0])
        R_ = ints_to_limbs([int(x) % R for x in r])
        mats = [ints_to_limbs([int(x) % R for row in M[:m] for x in row]) for M in (ap, bp, cp)]
        ax, bx, cx = (np.zeros((n, 4), dtype=np.uint64) for _ in range(3))
        px = np.zeros((2 * n - 1, 4), dtype=np.uint64)
        check(lib().b200_combine_polynomials(ptr(R_), m, ptr(mats[0]), ptr(mats[1]), ptr(mats[2]), n, ptr(ax), ptr(bx),
                                             ptr(cx), ptr(px)))
        return limbs_to_ints(ax), limbs_to_ints(bx), limbs_to_ints(cx), limbs_to_ints(px)


class SparseR1CS:
    """A sparse R1CS resident on the device (b200_r1cs_load): the large-n form of the dense ``a, b, c [][]*big.Int``
    arguments of R1CSToQAP (r1csqap.go:161), which cannot exist at 2^16+ constraints.

    ``mats`` = three (rowptr, col, val) triples in CSR form — rowptr/col numpy uint32, val a list of ints or an
    (nnz, 4) uint64 limb array — or three dense row-major matrices (lists of lists), which are converted."""

    def __init__(self, n, m, mats):
        self.n, self.m = n, m
        self._keep = []
        args = []
        for M in mats:
            if isinstance(M, (list, tuple)) and len(M) == 3 and hasattr(M[0], "dtype"):
                rowptr, col, val = M
            else:
                rowptr, col, val = self.dense_to_csr(M)
            rowptr = np.ascontiguousarray(rowptr, dtype=np.uint32)
            col = np.ascontiguousarray(col, dtype=np.uint32)
            if not hasattr(val, "dtype"):
                val = ints_to_limbs([int(x) % R for x in val]) if len(val) else np.zeros((1, 4), dtype=np.uint64)
            val = np.ascontiguousarray(val, dtype=np.uint64)
            assert rowptr.shape[0] == n + 1
            self._keep += [rowptr, col, val]
            args += [ptr(rowptr), ptr(col) if col.size else None, ptr(val)]
        from . import _lib
        h = _lib._h(0)
        check(lib().b200_r1cs_load(n, m, *args, h))
        self.handle = h.value

    @staticmethod
    def dense_to_csr(M):
        rowptr, col, val = [0], [], []
        for row in M:
            for i, x in enumerate(row):
                if int(x) % R:
                    col.append(i)
                    val.append(int(x) % R)
            rowptr.append(len(col))
        return np.array(rowptr, dtype=np.uint32), np.array(col, dtype=np.uint32), val

    def CombinePolynomials(self, w, want_abc=True):
        """(ax, bx, cx, px) == PolynomialField.CombinePolynomials(w, *R1CSToQAP(a, b, c)[:3]) (r1csqap.go:161-210)."""
        return [limbs_to_ints(x) if x is not None else None for x in self.combine_limbs(_coeffs(w), want_abc)]

    def combine_limbs(self, w_limbs, want_abc=True):
        n = self.n
        abc = [np.zeros((n, 4), dtype=np.uint64) if want_abc else None for _ in range(3)]
        px = np.zeros((2 * n - 1, 4), dtype=np.uint64)
        check(lib().b200_qap_px(self.handle, ptr(w_limbs), w_limbs.shape[0], *[ptr(x) if x is not None else None for x in abc],
                                ptr(px)))
        return abc[0], abc[1], abc[2], px

    def EvalAt(self, tau, nz=None):
        """([Eval(alphas[i], tau)], [.. betas ..], [.. gammas ..], Eval(Z, tau)) for all m signals — the evaluation loops of
        GenerateTrustedSetup (groth16.go:164-205) without forming the dense polynomials."""
        m = self.m
        nz = m - 2 if nz is None else nz
        outs = [np.zeros((m, 4), dtype=np.uint64) for _ in range(3)]
        zt = np.zeros(4, dtype=np.uint64)
        T = ints_to_limbs([int(tau) % R])
        check(lib().b200_qap_eval_at(self.handle, ptr(T), nz, ptr(outs[0]), ptr(outs[1]), ptr(outs[2]), ptr(zt)))
        return outs[0], outs[1], outs[2], limbs_to_ints(zt)[0]

    def free(self):
        if self.handle:
            check(lib().b200_r1cs_free(self.handle))
            self.handle = 0


def Transpose(matrix):                        # r1csqap.go:11-21
    return [[matrix[j][i] for j in range(len(matrix))] for i in range(len(matrix[0]))]
