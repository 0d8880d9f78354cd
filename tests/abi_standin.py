"""TEST INFRASTRUCTURE ONLY: an oracle-backed stand-in for the part of the libb200snark C ABI that the host mirrors of
the setup / verify rows call, so that their host logic (operation order, marshalling, check order, error behaviour)
is exercised by the CPU suite.  It reads and writes the same raw limb buffers through the pointers the mirrors pass.
The `-m gpu` tests run the same mirrors against the real library; nothing in the package imports this file."""
import ctypes

import numpy as np

from oracle import ref_py as o

G1, G2, F, PF = o.BN.G1, o.BN.G2, o.FQR, o.PF


def _rd(p, nwords):
    addr = p.value if isinstance(p, ctypes.c_void_p) else ctypes.cast(p, ctypes.c_void_p).value
    raw = ctypes.string_at(addr, 8 * nwords)
    return [int.from_bytes(raw[i:i + 32], "little") for i in range(0, len(raw), 32)]


def _wr(p, vals):
    buf = b"".join(int(v).to_bytes(32, "little") for v in vals)
    ctypes.memmove(p.value, buf, len(buf))


def _g1s(v):
    return [tuple(v[i:i + 3]) for i in range(0, len(v), 3)]


def _g2s(v):
    return [((v[i], v[i + 1]), (v[i + 2], v[i + 3]), (v[i + 4], v[i + 5])) for i in range(0, len(v), 6)]


def _flat(pts):
    out = []
    for p in pts:
        for c in p:
            out.extend(c if isinstance(c, tuple) else (c,))
    return out


def _f12s(v):
    return [tuple(tuple((v[12 * i + 6 * h + 2 * k], v[12 * i + 6 * h + 2 * k + 1]) for k in range(3)) for h in range(2))
            for i in range(len(v) // 12)]


class StandIn:
    def __init__(self):
        self.err = b""
        self.calls = []

    def b200_last_error(self):
        return self.err

    def _mul(self, grp, rd, words, pts, sc, n, out, bcast):
        self.calls.append(("mul", n))
        P = rd(_rd(pts, words * (1 if bcast else n)))
        S = _rd(sc, 4 * n)
        _wr(out, _flat([grp.mul_scalar(P[0] if bcast else P[i], S[i]) for i in range(n)]))
        return 0

    def b200_g1_mul_batch(self, pts, sc, n, out):
        return self._mul(G1, _g1s, 12, pts, sc, n, out, False)

    def b200_g1_mul_batch_bcast(self, pts, sc, n, out):
        return self._mul(G1, _g1s, 12, pts, sc, n, out, True)

    def b200_g2_mul_batch(self, pts, sc, n, out):
        return self._mul(G2, _g2s, 24, pts, sc, n, out, False)

    def b200_g2_mul_batch_bcast(self, pts, sc, n, out):
        return self._mul(G2, _g2s, 24, pts, sc, n, out, True)

    def b200_g1_add_batch(self, a, b, n, out):
        A, B = _g1s(_rd(a, 12 * n)), _g1s(_rd(b, 12 * n))
        _wr(out, _flat([G1.add(x, y) for x, y in zip(A, B)]))
        return 0

    def b200_g1_affine_batch(self, a, n, out):
        res = []
        for p in _g1s(_rd(a, 12 * n)):
            res.extend((0, 0) if G1.is_zero(p) else G1.affine(p)[:2])
        _wr(out, res)
        return 0

    def b200_zero_poly(self, n, out):
        z = [1]
        for i in range(1, n + 1):
            z = PF.mul(z, [F.neg(i), 1])
        _wr(out, z)
        return 0

    def b200_poly_eval_batch(self, P, m, n, x, out):
        coeffs = _rd(P, 4 * m * n)
        xv = _rd(x, 4)[0]
        _wr(out, [PF.eval(coeffs[i * n:(i + 1) * n], xv) for i in range(m)])
        return 0

    def b200_r1cs_to_qap(self, a, b, c, n, m, alphas, betas, gammas, z):
        for src, dst in ((a, alphas), (b, betas), (c, gammas)):
            v = _rd(src, 4 * n * m)
            rows = [v[j * m:(j + 1) * m] for j in range(n)]
            _wr(dst, [x for col in o.transpose(rows) for x in PF.lagrange_interpolation(col)])
        zp = [1]
        for i in range(1, m - 1):
            zp = PF.mul(zp, [F.neg(i), 1])
        _wr(z, zp)
        return 0

    def b200_combine_polynomials(self, r, m, ap, bp, cp, n, ax, bx, cx, px):
        rv = _rd(r, 4 * m)
        mats = [[v[i * n:(i + 1) * n] for i in range(m)] for v in (_rd(p, 4 * m * n) for p in (ap, bp, cp))]
        res = PF.combine_polynomials(rv, *mats)
        for dst, val, ln in zip((ax, bx, cx, px), res, (n, n, n, 2 * n - 1)):
            _wr(dst, list(val) + [0] * (ln - len(val)))
        return 0

    def b200_init(self, dev):
        return 0

    def b200_pairing_batch(self, g1, g2, n, out):
        self.calls.append(("pairing", n))
        A, B = _g1s(_rd(g1, 12 * n)), _g2s(_rd(g2, 24 * n))
        res = []
        for p, q in zip(A, B):
            if G2.is_zero(q):
                self.err = b"pairing_batch: q1[2] != Fq2.One()"
                return -3
            res.extend(c for h in o.BN.pairing(p, q) for f2 in h for c in f2)
        _wr(out, res)
        return 0

    def b200_fq12_mul_batch(self, a, b, n, out):
        X, Y = _f12s(_rd(a, 48 * n)), _f12s(_rd(b, 48 * n))
        _wr(out, [c for x, y in zip(X, Y) for h in o.BN.Fq12.mul(x, y) for f2 in h for c in f2])
        return 0

    def b200_groth16_verify(self, ic, n_ic, a1, b2, g2, d2, pa, pb, pc, pub, npub, ok):
        if n_ic < npub + 1:
            self.err = b"groth16_verify: len(IC) < len(publicSignals) + 1"
            return -3
        vk = {"IC": _g1s(_rd(ic, 12 * n_ic)), "G1": {"Alpha": _g1s(_rd(a1, 12))[0]},
              "G2": {"Beta": _g2s(_rd(b2, 24))[0], "Gamma": _g2s(_rd(g2, 24))[0], "Delta": _g2s(_rd(d2, 24))[0]}}
        proof = {"PiA": _g1s(_rd(pa, 12))[0], "PiB": _g2s(_rd(pb, 24))[0], "PiC": _g1s(_rd(pc, 12))[0]}
        sig = _rd(pub, 4 * npub) if npub else []
        ok._obj.value = 1 if o.groth16_verify(vk, proof, sig) else 0
        return 0


def install(monkeypatch):
    """Route the mirrors' lib() to a StandIn (every module binds `lib` by name at import)."""
    import gosnark_b200  # noqa: F401  (import shim)
    from gosnark_b200 import _lib, bn128, groth16, r1csqap, snark
    s = StandIn()
    for mod in (_lib, bn128, groth16, r1csqap, snark):
        monkeypatch.setattr(mod, "lib", lambda s=s: s)
    return s
