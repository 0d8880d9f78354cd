"""CPU oracle (TEST INFRASTRUCTURE ONLY) — a Python-int restatement of the
reference's prove-path arithmetic, in the reference's exact operation order.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` leg may import this module.  The product path
(``go-snark-study_b200/``) never does; it fails loudly without its CUDA library.

Every function cites the reference file:line (relative to the upstream repo
arnaucube/go-snark-study @ 4780061) it restates.  Because the operation order is
the reference's, results match the Go code bit-for-bit *including* the Jacobian
(X, Y, Z) representation, which is how this oracle is pinned against outputs
of the reference's own prebuilt Go binary (tests/golden/, oracle/make_golden.py).

Pinned by: K1 (bn128/g1_test.go:27-28), K3/K4 (wasm/index.js:7-8), K5
(Pinocchio proof from the Go binary), K6, K8 (circom-test/verification_key.json)
— see tests/test_oracle_golden.py.
"""

# ---------------------------------------------------------------- constants
# bn128/bn128.go:40-93
Q = 21888242871839275222246405745257275088696311157297823662689037894645226208583
R = 21888242871839275222246405745257275088548364400416034343698204186575808495617
G1_GEN = (1, 2, 1)
G2_GEN = (
    (10857046999023057135944570762232829481370756359578518086990519993285655852781,
     11559732032986387107991004021392285783925812861821192530917403151452391805634),
    (8495653923123431417604973247489272438418190587263600148770280649306958101930,
     4082367875863433681332203403145435568316851327593401208105741076214120093531),
    (1, 0),
)
NONRESIDUE_FQ2 = Q - 1          # bn128.go:86
NONRESIDUE_FQ6 = (9, 1)         # bn128.go:90-93


# ------------------------------------------------------------------- fields
class Fq:
    """fields/fq.go:10-171 — canonical-range modular ops over modulus ``q``."""

    def __init__(self, q):
        self.Q = q

    def zero(self):
        return 0

    def one(self):
        return 1

    def add(self, a, b):            # fq.go:32-35
        return (a + b) % self.Q

    def double(self, a):            # fq.go:38-41
        return (a + a) % self.Q

    def sub(self, a, b):            # fq.go:44-47
        return (a - b) % self.Q

    def neg(self, a):               # fq.go:50-53
        return (-a) % self.Q

    def mul(self, a, b):            # fq.go:56-59
        return (a * b) % self.Q

    def mul_scalar(self, a, e):     # fq.go:61-63
        return self.mul(a, e)

    def inverse(self, a):           # fq.go:66-68 (big.Int.ModInverse)
        return pow(a, -1, self.Q)

    def div(self, a, b):            # fq.go:71-74
        return self.mul(a, self.inverse(b)) % self.Q

    def square(self, a):            # fq.go:77-80
        return (a * a) % self.Q

    def exp(self, base, e):         # fq.go:101-114 (LSB-first square & multiply)
        res, rem, ex = 1, abs(e), base
        while rem:
            if rem & 1:
                res = self.mul(res, ex)
            ex = self.square(ex)
            rem >>= 1
        return res

    def is_zero(self, a):           # fq.go:134-136
        return a == 0

    def copy(self, a):              # fq.go:138-140: via Bytes() => |a| (sign lost, H7)
        return abs(a)

    def affine(self, a):            # fq.go:142-160
        return a % self.Q

    def equal(self, a, b):          # fq.go:162-166
        return a % self.Q == b % self.Q


class Fq2:
    """fields/fq2.go:8-154 — F[u]/(u^2 - nonresidue), elements are 2-tuples."""

    def __init__(self, f, nonresidue):
        self.F = f
        self.nonresidue = nonresidue

    def zero(self):
        return (0, 0)

    def one(self):
        return (1, 0)

    def _mul_nr(self, a):           # fq2.go:33-35
        return self.F.mul(self.nonresidue, a)

    def add(self, a, b):            # fq2.go:37-42
        return (self.F.add(a[0], b[0]), self.F.add(a[1], b[1]))

    def double(self, a):
        return self.add(a, a)

    def sub(self, a, b):            # fq2.go:49-54
        return (self.F.sub(a[0], b[0]), self.F.sub(a[1], b[1]))

    def neg(self, a):               # fq2.go:57-59
        return self.sub(self.zero(), a)

    def mul(self, a, b):            # fq2.go:63-76 (Karatsuba)
        F = self.F
        v0 = F.mul(a[0], b[0])
        v1 = F.mul(a[1], b[1])
        return (F.add(v0, self._mul_nr(v1)),
                F.sub(F.mul(F.add(a[0], a[1]), F.add(b[0], b[1])), F.add(v0, v1)))

    def mul_scalar(self, p, e):     # fq2.go:78-96 (double-and-add == componentwise e*p)
        e = abs(e)
        return (self.F.mul(p[0], e % self.F.Q), self.F.mul(p[1], e % self.F.Q))

    def inverse(self, a):           # fq2.go:99-108
        F = self.F
        t0 = F.square(a[0])
        t1 = F.square(a[1])
        t2 = F.sub(t0, self._mul_nr(t1))
        t3 = F.inverse(t2)
        return (F.mul(a[0], t3), F.neg(F.mul(a[1], t3)))

    def div(self, a, b):
        return self.mul(a, self.inverse(b))

    def square(self, a):            # fq2.go:118-133
        F = self.F
        ab = F.mul(a[0], a[1])
        return (F.sub(F.mul(F.add(a[0], a[1]), F.add(a[0], self._mul_nr(a[1]))),
                      F.add(ab, self._mul_nr(ab))),
                F.add(ab, ab))

    def is_zero(self, a):
        return a[0] == 0 and a[1] == 0

    def affine(self, a):
        return (self.F.affine(a[0]), self.F.affine(a[1]))

    def equal(self, a, b):
        return self.F.equal(a[0], b[0]) and self.F.equal(a[1], b[1])

    def copy(self, a):
        return (a[0], a[1])


class Fq6:
    """fields/fq6.go:9-192 — Fq2[v]/(v^3 - nonresidue), 3-tuples of Fq2."""

    def __init__(self, f2, nonresidue):
        self.F = f2
        self.nonresidue = nonresidue

    def zero(self):
        return (self.F.zero(),) * 3

    def one(self):
        return (self.F.one(), self.F.zero(), self.F.zero())

    def _mul_nr(self, a):
        return self.F.mul(self.nonresidue, a)

    def add(self, a, b):
        return tuple(self.F.add(x, y) for x, y in zip(a, b))

    def sub(self, a, b):
        return tuple(self.F.sub(x, y) for x, y in zip(a, b))

    def neg(self, a):
        return self.sub(self.zero(), a)

    def mul(self, a, b):            # fq6.go:65-95
        F = self.F
        v0, v1, v2 = F.mul(a[0], b[0]), F.mul(a[1], b[1]), F.mul(a[2], b[2])
        return (
            F.add(v0, self._mul_nr(F.sub(F.mul(F.add(a[1], a[2]), F.add(b[1], b[2])), F.add(v1, v2)))),
            F.add(F.sub(F.mul(F.add(a[0], a[1]), F.add(b[0], b[1])), F.add(v0, v1)), self._mul_nr(v2)),
            F.add(F.sub(F.mul(F.add(a[0], a[2]), F.add(b[0], b[2])), F.add(v0, v2)), v1),
        )

    def inverse(self, a):           # fq6.go:115-140
        F = self.F
        t0, t1, t2 = F.square(a[0]), F.square(a[1]), F.square(a[2])
        t3, t4, t5 = F.mul(a[0], a[1]), F.mul(a[0], a[2]), F.mul(a[1], a[2])
        c0 = F.sub(t0, self._mul_nr(t5))
        c1 = F.sub(self._mul_nr(t2), t3)
        c2 = F.sub(t1, t4)
        t6 = F.inverse(F.add(F.mul(a[0], c0),
                             self._mul_nr(F.add(F.mul(a[2], c1), F.mul(a[1], c2)))))
        return (F.mul(t6, c0), F.mul(t6, c1), F.mul(t6, c2))

    def square(self, a):            # fq6.go:146-172
        F = self.F
        s0 = F.square(a[0])
        ab = F.mul(a[0], a[1])
        s1 = F.add(ab, ab)
        s2 = F.square(F.add(F.sub(a[0], a[1]), a[2]))
        bc = F.mul(a[1], a[2])
        s3 = F.add(bc, bc)
        s4 = F.square(a[2])
        return (F.add(s0, self._mul_nr(s3)),
                F.add(s1, self._mul_nr(s4)),
                F.sub(F.add(F.add(s1, s2), s3), F.add(s0, s4)))

    def equal(self, a, b):
        return all(self.F.equal(x, y) for x, y in zip(a, b))


class Fq12:
    """fields/fq12.go:11-165 — Fq6[w]/(w^2 - v), 2-tuples of Fq6."""

    def __init__(self, f6, f2, nonresidue):
        self.F = f6
        self.Fq2 = f2
        self.nonresidue = nonresidue

    def zero(self):
        return (self.F.zero(), self.F.zero())

    def one(self):
        return (self.F.one(), self.F.zero())

    def _mul_nr(self, a):           # fq12.go:35-41
        return (self.Fq2.mul(self.nonresidue, a[2]), a[0], a[1])

    def add(self, a, b):
        return (self.F.add(a[0], b[0]), self.F.add(a[1], b[1]))

    def sub(self, a, b):
        return (self.F.sub(a[0], b[0]), self.F.sub(a[1], b[1]))

    def mul(self, a, b):            # fq12.go:72-84
        F = self.F
        v0 = F.mul(a[0], b[0])
        v1 = F.mul(a[1], b[1])
        return (F.add(v0, self._mul_nr(v1)),
                F.sub(F.mul(F.add(a[0], a[1]), F.add(b[0], b[1])), F.add(v0, v1)))

    def inverse(self, a):           # fq12.go:105-114
        F = self.F
        t0 = F.square(a[0])
        t1 = F.square(a[1])
        t2 = F.sub(t0, self._mul_nr(t1))
        t3 = F.inverse(t2)
        return (F.mul(a[0], t3), F.neg(F.mul(a[1], t3)))

    def div(self, a, b):
        return self.mul(a, self.inverse(b))

    def square(self, a):            # fq12.go:122-137
        F = self.F
        ab = F.mul(a[0], a[1])
        return (F.sub(F.mul(F.add(a[0], a[1]), F.add(a[0], self._mul_nr(a[1]))),
                      F.add(ab, self._mul_nr(ab))),
                F.add(ab, ab))

    def exp(self, base, e):         # fq12.go:139-156
        res, rem, ex = self.one(), e, base
        while rem:
            if rem & 1:
                res = self.mul(res, ex)
            ex = self.square(ex)
            rem >>= 1
        return res

    def equal(self, a, b):
        return self.F.equal(a[0], b[0]) and self.F.equal(a[1], b[1])


# ------------------------------------------------------------------- groups
class Jacobian:
    """bn128/g1.go:28-193 and bn128/g2.go:25-223 — the two files hold the same
    formulas over Fq and Fq2; ``F`` supplies the field.  Points are 3-tuples."""

    def __init__(self, f, gen, zero_affine):
        self.F = f
        self.G = gen
        self._zero_affine = zero_affine

    def is_zero(self, p):           # g1.go:28-30 / g2.go:28-30
        return self.F.is_zero(p[2])

    def zero3(self):
        z = self.F.zero()
        return (z, z, z)

    def add(self, p1, p2):          # g1.go:32-89 / g2.go:32-89 (add-2007-bl, NO doubling branch: H6)
        F = self.F
        if self.is_zero(p1):
            return p2
        if self.is_zero(p2):
            return p1
        x1, y1, z1 = p1
        x2, y2, z2 = p2
        z1z1 = F.square(z1)
        z2z2 = F.square(z2)
        u1 = F.mul(x1, z2z2)
        u2 = F.mul(x2, z1z1)
        t0 = F.mul(z2, z2z2)
        s1 = F.mul(y1, t0)
        t1 = F.mul(z1, z1z1)
        s2 = F.mul(y2, t1)
        h = F.sub(u2, u1)
        t2 = F.add(h, h)
        i = F.square(t2)
        j = F.mul(h, i)
        t3 = F.sub(s2, s1)
        r = F.add(t3, t3)
        v = F.mul(u1, i)
        t4 = F.square(r)
        t5 = F.add(v, v)
        t6 = F.sub(t4, j)
        x3 = F.sub(t6, t5)
        t7 = F.sub(v, x3)
        t8 = F.mul(s1, j)
        t9 = F.add(t8, t8)
        t10 = F.mul(r, t7)
        y3 = F.sub(t10, t9)
        t11 = F.add(z1, z2)
        t12 = F.square(t11)
        t13 = F.sub(t12, z1z1)
        t14 = F.sub(t13, z2z2)
        z3 = F.mul(t14, h)
        return (x3, y3, z3)

    def neg(self, p):               # g1.go:91-97
        return (p[0], self.F.neg(p[1]), p[2])

    def sub(self, a, b):
        return self.add(a, self.neg(b))

    def double(self, p):            # g1.go:101-138 / g2.go:103-140 (dbl-2009-l)
        F = self.F
        if self.is_zero(p):
            return p
        a = F.square(p[0])
        b = F.square(p[1])
        c = F.square(b)
        t0 = F.add(p[0], b)
        t1 = F.square(t0)
        t2 = F.sub(t1, a)
        t3 = F.sub(t2, c)
        d = F.double(t3)
        e = F.add(F.add(a, a), a)
        f = F.square(e)
        t4 = F.double(d)
        x3 = F.sub(f, t4)
        t5 = F.sub(d, x3)
        two_c = F.add(c, c)
        four_c = F.add(two_c, two_c)
        t6 = F.add(four_c, four_c)
        t7 = F.mul(e, t5)
        y3 = F.sub(t7, t6)
        t8 = F.mul(p[1], p[2])
        z3 = F.double(t8)
        return (x3, y3, z3)

    def mul_scalar(self, p, e):     # g1.go:140-155 / g2.go:142-181 (MSB-first; |e| used, H7)
        q = self.zero3()
        d = abs(e)
        for i in range(d.bit_length() - 1, -1, -1):
            q = self.double(q)
            if (d >> i) & 1:
                q = self.add(q, p)
        return q

    def affine(self, p):            # g1.go:157-170 / g2.go:183-200
        F = self.F
        if self.is_zero(p):
            return self._zero_affine
        zinv = F.inverse(p[2])
        zinv2 = F.square(zinv)
        x = F.mul(p[0], zinv2)
        zinv3 = F.mul(zinv2, zinv)
        y = F.mul(p[1], zinv3)
        if len(self._zero_affine) == 3:          # G2.Affine returns a 3-tuple with Z = one
            return (F.affine(x), F.affine(y), F.one())
        return (x, y)

    def equal(self, p1, p2):        # g1.go:172-193
        F = self.F
        if self.is_zero(p1):
            return self.is_zero(p2)
        if self.is_zero(p2):
            return self.is_zero(p1)
        z1z1 = F.square(p1[2])
        z2z2 = F.square(p2[2])
        u1 = F.mul(p1[0], z2z2)
        u2 = F.mul(p2[0], z1z1)
        s1 = F.mul(p1[1], F.mul(p2[2], z2z2))
        s2 = F.mul(p2[1], F.mul(p1[2], z1z1))
        return F.equal(u1, u2) and F.equal(s1, s2)


class Bn128:
    """bn128/bn128.go:38-421."""

    def __init__(self):
        self.Q, self.R = Q, R
        self.Fq1 = Fq(Q)
        self.Fq2 = Fq2(self.Fq1, NONRESIDUE_FQ2)
        self.Fq6 = Fq6(self.Fq2, NONRESIDUE_FQ6)
        self.Fq12 = Fq12(self.Fq6, self.Fq2, NONRESIDUE_FQ6)
        self.G1 = Jacobian(self.Fq1, G1_GEN, (0, 0))                       # g1.go:25-27
        self.G2 = Jacobian(self.Fq2, G2_GEN, ((0, 0), (1, 0), (0, 0)))     # g2.go:25-27
        # preparePairing, bn128.go:120-177
        self.loop_count = 29793968203157093288
        self.loop_count_neg = False
        self.two_inv = self.Fq1.inverse(2)
        self.twist = (9, 1)
        self.twist_coef_b = self.Fq2.mul_scalar(self.Fq2.inverse(self.twist), 3)
        self.frob_c11 = Q - 1
        self.twist_mul_by_q_x = (
            21575463638280843010398324269430826099269044274347216827212613867836435027261,
            10307601595873709700152284273816112264069230130616436755625194854815875713954)
        self.twist_mul_by_q_y = (
            2821565182194536844548159561693502659359617185244120367078079554186484126554,
            3505843767911556378687030309984248845540243509899259641013678093033130930403)
        self.final_exp = (Q ** 12 - 1) // R                                  # bn128.go:168 literal == (q^12-1)/r

    # -- pairing (verify side; "next" row f2) -------------------------------
    def _doubling_step(self, cur):      # bn128.go:262-294
        F = self.Fq2
        x, y, z = cur
        a = F.mul_scalar(F.mul(x, y), self.two_inv)
        b = F.square(y)
        c = F.square(z)
        d = F.add(c, F.add(c, c))
        e = F.mul(self.twist_coef_b, d)
        f = F.add(e, F.add(e, e))
        g = F.mul_scalar(F.add(b, f), self.two_inv)
        h = F.sub(F.square(F.add(y, z)), F.add(b, c))
        i = F.sub(e, b)
        j = F.square(x)
        e_sqr = F.square(e)
        nx = F.mul(a, F.sub(b, f))
        ny = F.sub(F.sub(F.square(g), e_sqr), F.add(e_sqr, e_sqr))
        nz = F.mul(b, h)
        coef = (F.mul(i, self.twist), F.neg(h), F.add(j, F.add(j, j)))
        return coef, (nx, ny, nz)

    def _mixed_addition_step(self, base, cur):   # bn128.go:296-330
        F = self.Fq2
        x1, y1, z1 = cur
        x2, y2 = base[0], base[1]
        d = F.sub(x1, F.mul(x2, z1))
        e = F.sub(y1, F.mul(y2, z1))
        f = F.square(d)
        g = F.square(e)
        h = F.mul(d, f)
        i = F.mul(x1, f)
        j = F.sub(F.add(h, F.mul(z1, g)), F.add(i, i))
        nx = F.mul(d, j)
        ny = F.sub(F.mul(e, F.sub(i, j)), F.mul(h, y1))
        nz = F.mul(z1, h)
        coef = (F.mul(self.twist, F.sub(F.mul(e, x2), F.mul(d, y2))), d, F.neg(e))
        return coef, (nx, ny, nz)

    def _g2_mul_by_q(self, p):          # bn128.go:331-346
        F1, F2 = self.Fq1, self.Fq2
        fmx = (p[0][0], F1.mul(p[0][1], self.frob_c11))
        fmy = (p[1][0], F1.mul(p[1][1], self.frob_c11))
        fmz = (p[2][0], F1.mul(p[2][1], self.frob_c11))
        return (F2.mul(self.twist_mul_by_q_x, fmx), F2.mul(self.twist_mul_by_q_y, fmy), fmz)

    def _precompute_g2(self, p):        # bn128.go:213-260
        F = self.Fq2
        q = self.G2.affine(p)
        coeffs = []
        r = (q[0], q[1], F.one())
        for i in range(self.loop_count.bit_length() - 2, -1, -1):
            c, r = self._doubling_step(r)
            coeffs.append(c)
            if (self.loop_count >> i) & 1:
                c, r = self._mixed_addition_step(q, r)
                coeffs.append(c)
        q1 = self.G2.affine(self._g2_mul_by_q(q))
        assert F.equal(q1[2], F.one())
        q2 = self.G2.affine(self._g2_mul_by_q(q1))
        assert F.equal(q2[2], F.one())
        q2 = (q2[0], F.neg(q2[1]), q2[2])
        c, r = self._mixed_addition_step(q1, r)
        coeffs.append(c)
        c, r = self._mixed_addition_step(q2, r)
        coeffs.append(c)
        return coeffs

    def _mul_by_024(self, a, ell0, ell_vw, ell_vv):   # bn128.go:402-416
        z = self.Fq2.zero()
        return self.Fq12.mul(a, ((ell0, z, ell_vv), (z, ell_vw, z)))

    def miller_loop(self, p1_affine, coeffs):         # bn128.go:348-400
        F2 = self.Fq2
        px, py = p1_affine
        f = self.Fq12.one()
        idx = 0

        def line(f, c):
            return self._mul_by_024(f, c[0], F2.mul_scalar(c[1], py), F2.mul_scalar(c[2], px))

        for i in range(self.loop_count.bit_length() - 2, -1, -1):
            f = self.Fq12.square(f)
            f = line(f, coeffs[idx]); idx += 1
            if (self.loop_count >> i) & 1:
                f = line(f, coeffs[idx]); idx += 1
        f = line(f, coeffs[idx]); idx += 1
        f = line(f, coeffs[idx]); idx += 1
        return f

    def pairing(self, p1, p2):          # bn128.go:179-186
        pre1 = self.G1.affine(p1)
        pre2 = self._precompute_g2(p2)
        return self.Fq12.exp(self.miller_loop(pre1, pre2), self.final_exp)


BN = Bn128()
FQR = Fq(R)


# ------------------------------------------------------------ r1csqap (L3)
def transpose(m):                       # r1csqap.go:11-21
    return [[m[j][i] for j in range(len(m))] for i in range(len(m[0]))]


class PolynomialField:
    """r1csqap/r1csqap.go:45-216.  Coefficient lists, index = power of x.
    ``new_pol_zero_at`` uses exact big integers for ``fac`` (the reference uses a
    native int that overflows for n > 21); identical for n <= 21."""

    def __init__(self, f):
        self.F = f

    def mul(self, a, b):                # r1csqap.go:57-67
        F = self.F
        r = [0] * (len(a) + len(b) - 1)
        for i, ai in enumerate(a):
            for j, bj in enumerate(b):
                r[i + j] = F.add(r[i + j], F.mul(ai, bj))
        return r

    def div(self, a, b):                # r1csqap.go:70-84
        F = self.F
        r = [0] * max(len(a) - len(b) + 1, 0)
        rem = list(a)
        while len(rem) >= len(b):
            l = F.div(rem[-1], b[-1])
            pos = len(rem) - len(b)
            r[pos] = l
            aux2 = self.sub(rem, self.mul(b, [0] * pos + [l]))
            rem = aux2[:-1]
        return r, rem

    def add(self, a, b):                # r1csqap.go:94-103
        F = self.F
        r = [0] * max(len(a), len(b))
        for i, x in enumerate(a):
            r[i] = F.add(r[i], x)
        for i, x in enumerate(b):
            r[i] = F.add(r[i], x)
        return r

    def sub(self, a, b):                # r1csqap.go:106-115
        F = self.F
        r = [0] * max(len(a), len(b))
        for i, x in enumerate(a):
            r[i] = F.add(r[i], x)
        for i, x in enumerate(b):
            r[i] = F.sub(r[i], x)
        return r

    def eval(self, v, x):               # r1csqap.go:118-126
        F = self.F
        r = 0
        for i, c in enumerate(v):
            r = F.add(r, F.mul(c, F.exp(x, i)))
        return r

    def new_pol_zero_at(self, point_pos, total_points, height):   # r1csqap.go:129-147
        F = self.F
        fac = 1
        for i in range(1, total_points + 1):
            if i != point_pos:
                fac *= point_pos - i
        # big.NewInt(fac) may be negative; Fq.Div -> ModInverse of a negative => python pow handles sign mod R
        hf = F.div(height, fac % F.Q)
        r = [hf]
        for i in range(1, total_points + 1):
            if i != point_pos:
                r = self.mul(r, [-i, 1])
        return r

    def lagrange_interpolation(self, v):   # r1csqap.go:150-158
        r = []
        for i, vi in enumerate(v):
            r = self.add(r, self.new_pol_zero_at(i + 1, len(v), vi))
        return r

    def r1cs_to_qap(self, a, b, c):     # r1csqap.go:161-188
        alphas = [self.lagrange_interpolation(col) for col in transpose(a)]
        betas = [self.lagrange_interpolation(col) for col in transpose(b)]
        gammas = [self.lagrange_interpolation(col) for col in transpose(c)]
        z = [1]
        for i in range(1, len(alphas) - 1):
            z = self.mul(z, [self.F.neg(i), 1])
        return alphas, betas, gammas, z

    def combine_polynomials(self, r, ap, bp, cp):   # r1csqap.go:191-210
        ax, bx, cx = [], [], []
        for i, ri in enumerate(r):
            ax = self.add(ax, self.mul([ri], ap[i]))
        for i, ri in enumerate(r):
            bx = self.add(bx, self.mul([ri], bp[i]))
        for i, ri in enumerate(r):
            cx = self.add(cx, self.mul([ri], cp[i]))
        px = self.sub(self.mul(ax, bx), cx)
        return ax, bx, cx, px

    def divisor_polynomial(self, px, z):   # r1csqap.go:213-216
        return self.div(px, z)[0]


PF = PolynomialField(FQR)


# --------------------------------------------------------- Groth16 (L4)
def groth16_setup(n_vars, n_public, alphas, betas, gammas, toxic):
    """groth16/groth16.go:94-222 with the five toxic values injected
    (``toxic`` = dict T,Kalpha,Kbeta,Kgamma,Kdelta) instead of crypto/rand."""
    G1, G2, F = BN.G1, BN.G2, FQR
    t, ka, kb, kg, kd = (toxic[k] for k in ("T", "Kalpha", "Kbeta", "Kgamma", "Kdelta"))
    zpol = [1]
    for i in range(1, len(alphas) - 1):
        zpol = PF.mul(zpol, [F.neg(i), 1])
    zt = PF.eval(zpol, t)
    inv_delta = F.inverse(kd)
    zt_inv_delta = F.mul(inv_delta, zt)
    ptd = [G1.mul_scalar(G1.G, zt_inv_delta)]
    t_encr = t
    for i in range(1, len(zpol)):
        ptd.append(G1.mul_scalar(G1.G, F.mul(t_encr, zt_inv_delta)))
        t_encr = F.mul(t_encr, t)
    pk = {"Z": zpol, "PowersTauDelta": ptd,
          "G1": {"Alpha": G1.mul_scalar(G1.G, ka), "Beta": G1.mul_scalar(G1.G, kb),
                 "Delta": G1.mul_scalar(G1.G, kd), "At": [], "BACGamma": []},
          "G2": {"Beta": G2.mul_scalar(G2.G, kb), "Gamma": None,
                 "Delta": G2.mul_scalar(G2.G, kd), "BACGamma": []},
          "BACDelta": []}
    vk = {"IC": [], "G1": {"Alpha": G1.mul_scalar(G1.G, ka)},
          "G2": {"Beta": G2.mul_scalar(G2.G, kb), "Gamma": G2.mul_scalar(G2.G, kg),
                 "Delta": G2.mul_scalar(G2.G, kd)}}
    for i in range(n_vars):             # len(circuit.Signals)
        at = PF.eval(alphas[i], t)
        pk["G1"]["At"].append(G1.mul_scalar(G1.G, at))
        bt = PF.eval(betas[i], t)
        pk["G1"]["BACGamma"].append(G1.mul_scalar(G1.G, bt))
        pk["G2"]["BACGamma"].append(G2.mul_scalar(G2.G, bt))
    for i in range(n_public + 1):
        pk["BACDelta"].append((0, 0, 0))
    for i in range(n_public + 1, n_vars):
        at, bt, ct = PF.eval(alphas[i], t), PF.eval(betas[i], t), PF.eval(gammas[i], t)
        c = F.mul(inv_delta, F.add(F.add(F.mul(at, kb), F.mul(bt, ka)), ct))
        pk["BACDelta"].append(G1.mul_scalar(G1.G, c))
    for i in range(n_public + 1):
        at, bt, ct = PF.eval(alphas[i], t), PF.eval(betas[i], t), PF.eval(gammas[i], t)
        ic = F.mul(F.inverse(kg), F.add(F.add(F.mul(at, kb), F.mul(bt, ka)), ct))
        vk["IC"].append(G1.mul_scalar(G1.G, ic))
    return pk, vk


def groth16_prove(n_vars, n_public, pk, w, px, r, s):
    """groth16/groth16.go:225-278 with r, s injected (H2).  Returns
    (PiA, PiB, PiC) Jacobian, plus the raw sums for kernel-level parity."""
    G1, G2, F = BN.G1, BN.G2, FQR
    pi_a, pi_b, pi_c, pi_b_g1 = G1.zero3(), G2.zero3(), G1.zero3(), G1.zero3()
    for i in range(n_vars):             # :243-247
        pi_a = G1.add(pi_a, G1.mul_scalar(pk["G1"]["At"][i], w[i]))
        pi_b_g1 = G1.add(pi_b_g1, G1.mul_scalar(pk["G1"]["BACGamma"][i], w[i]))
        pi_b = G2.add(pi_b, G2.mul_scalar(pk["G2"]["BACGamma"][i], w[i]))
    for i in range(n_public + 1, n_vars):   # :248-250
        pi_c = G1.add(pi_c, G1.mul_scalar(pk["BACDelta"][i], w[i]))
    raw = {"A": pi_a, "B1": pi_b_g1, "B2": pi_b, "C": pi_c}
    pi_a = G1.add(pi_a, pk["G1"]["Alpha"])                       # :253-255
    pi_a = G1.add(pi_a, G1.mul_scalar(pk["G1"]["Delta"], r))
    pi_b_g1 = G1.add(pi_b_g1, pk["G1"]["Beta"])                  # :259-264
    pi_b = G2.add(pi_b, pk["G2"]["Beta"])
    pi_b_g1 = G1.add(pi_b_g1, G1.mul_scalar(pk["G1"]["Delta"], s))
    pi_b = G2.add(pi_b, G2.mul_scalar(pk["G2"]["Delta"], s))
    hx = PF.divisor_polynomial(px, pk["Z"])                      # :266
    h_sum = G1.zero3()
    for i in range(len(hx)):                                     # :269-271
        term = G1.mul_scalar(pk["PowersTauDelta"][i], hx[i])
        pi_c = G1.add(pi_c, term)
        h_sum = G1.add(h_sum, term)
    raw["H"] = h_sum
    raw["hx"] = hx
    pi_c = G1.add(pi_c, G1.mul_scalar(pi_a, s))                  # :272-275
    pi_c = G1.add(pi_c, G1.mul_scalar(pi_b_g1, r))
    neg_rs = F.neg(F.mul(r, s))
    pi_c = G1.add(pi_c, G1.mul_scalar(pk["G1"]["Delta"], neg_rs))
    return {"PiA": pi_a, "PiB": pi_b, "PiC": pi_c}, raw


def groth16_verify(vk, proof, public_signals):
    """groth16/groth16.go:281-305."""
    G1, F12 = BN.G1, BN.Fq12
    ic = vk["IC"][0]
    for i, sig in enumerate(public_signals):
        ic = G1.add(ic, G1.mul_scalar(vk["IC"][i + 1], sig))
    lhs = BN.pairing(proof["PiA"], proof["PiB"])
    rhs = F12.mul(BN.pairing(vk["G1"]["Alpha"], vk["G2"]["Beta"]),
                  F12.mul(BN.pairing(ic, vk["G2"]["Gamma"]),
                          BN.pairing(proof["PiC"], vk["G2"]["Delta"])))
    return F12.equal(lhs, rhs)


# -------------------------------------------------------- Pinocchio (L4)
def pinocchio_prove(n_vars, n_public, pk, w, px):
    """snark.go:254-289.  Deterministic (E4).  pk = dict with A, B(G2), C, Kp,
    Ap, Bp, Cp, G1T, Z."""
    G1, G2 = BN.G1, BN.G2
    p = {k: G1.zero3() for k in ("PiA", "PiAp", "PiBp", "PiC", "PiCp", "PiH", "PiKp")}
    p["PiB"] = G2.zero3()
    for i in range(n_public + 1, n_vars):       # :265-268
        p["PiA"] = G1.add(p["PiA"], G1.mul_scalar(pk["A"][i], w[i]))
        p["PiAp"] = G1.add(p["PiAp"], G1.mul_scalar(pk["Ap"][i], w[i]))
    for i in range(n_vars):                     # :270-278
        p["PiB"] = G2.add(p["PiB"], G2.mul_scalar(pk["B"][i], w[i]))
        p["PiBp"] = G1.add(p["PiBp"], G1.mul_scalar(pk["Bp"][i], w[i]))
        p["PiC"] = G1.add(p["PiC"], G1.mul_scalar(pk["C"][i], w[i]))
        p["PiCp"] = G1.add(p["PiCp"], G1.mul_scalar(pk["Cp"][i], w[i]))
        p["PiKp"] = G1.add(p["PiKp"], G1.mul_scalar(pk["Kp"][i], w[i]))
    hx = PF.divisor_polynomial(px, pk["Z"])     # :280
    for i in range(len(hx)):                    # :284-286
        p["PiH"] = G1.add(p["PiH"], G1.mul_scalar(pk["G1T"][i], hx[i]))
    return p, hx


def pinocchio_setup(n_vars, n_public, alphas, betas, gammas, toxic):
    """snark.go:98-251 with the toxic values injected (``toxic`` = dict T, Ka, Kb, Kc, Kbeta, Kgamma, RhoA, RhoB;
    RhoC = RhoA*RhoB as in :150).  Returns (pk, vk) in the current struct layout (G1T inside Pk, snark.go:16-37)."""
    G1, G2, F = BN.G1, BN.G2, FQR
    t, ka, kb, kc = toxic["T"], toxic["Ka"], toxic["Kb"], toxic["Kc"]
    kbeta, kgamma, rho_a, rho_b = toxic["Kbeta"], toxic["Kgamma"], toxic["RhoA"], toxic["RhoB"]
    rho_c = F.mul(rho_a, rho_b)
    kbg = F.mul(kbeta, kgamma)
    vk = {"Vka": G2.mul_scalar(G2.G, ka), "Vkb": G1.mul_scalar(G1.G, kb), "Vkc": G2.mul_scalar(G2.G, kc), "IC": [],
          "G1Kbg": G1.mul_scalar(G1.G, kbg), "G2Kbg": G2.mul_scalar(G2.G, kbg), "G2Kg": G2.mul_scalar(G2.G, kgamma)}
    pk = {k: [] for k in ("A", "B", "C", "Kp", "Ap", "Bp", "Cp")}
    for i in range(n_vars):                                   # :171-207
        rho_a_at = F.mul(rho_a, PF.eval(alphas[i], t))
        a = G1.mul_scalar(G1.G, rho_a_at)
        pk["A"].append(a)
        if i <= n_public:
            vk["IC"].append(a)
        rho_b_bt = F.mul(rho_b, PF.eval(betas[i], t))
        bg1 = G1.mul_scalar(G1.G, rho_b_bt)
        pk["B"].append(G2.mul_scalar(G2.G, rho_b_bt))
        rho_c_ct = F.mul(rho_c, PF.eval(gammas[i], t))
        c = G1.mul_scalar(G1.G, rho_c_ct)
        pk["C"].append(c)
        kt = F.add(F.add(rho_a_at, rho_b_bt), rho_c_ct)
        k_ = G1.mul_scalar(G1.G, kt)
        assert G1.affine(k_) == G1.affine(G1.add(G1.add(a, bg1), c))     # the reference's os.Exit(1) self-check, :194-199
        pk["Ap"].append(G1.mul_scalar(a, ka))
        pk["Bp"].append(G1.mul_scalar(bg1, kb))
        pk["Cp"].append(G1.mul_scalar(c, kc))
        pk["Kp"].append(G1.mul_scalar(k_, kbeta))
    zpol = [1]
    for i in range(1, len(alphas) - 1):                       # :210-221
        zpol = PF.mul(zpol, [F.neg(i), 1])
    pk["Z"] = zpol
    vk["Vkz"] = G2.mul_scalar(G2.G, F.mul(rho_c, PF.eval(zpol, t)))      # :224-227
    gt1 = [G1.G]
    t_encr = t
    for i in range(1, len(zpol)):                             # :230-237
        gt1.append(G1.mul_scalar(G1.G, t_encr))
        t_encr = F.mul(t_encr, t)
    pk["G1T"] = gt1
    return pk, vk


def pinocchio_verify(vk, proof, public_signals):
    """snark.go:292-372: the five pairing checks, in the reference's order; returns (ok, index of the first failed check or 0)."""
    G1, G2, F12 = BN.G1, BN.G2, BN.Fq12
    if not F12.equal(BN.pairing(proof["PiA"], vk["Vka"]), BN.pairing(proof["PiAp"], G2.G)):
        return False, 1
    if not F12.equal(BN.pairing(vk["Vkb"], proof["PiB"]), BN.pairing(proof["PiBp"], G2.G)):
        return False, 2
    if not F12.equal(BN.pairing(proof["PiC"], vk["Vkc"]), BN.pairing(proof["PiCp"], G2.G)):
        return False, 3
    vkxpia = vk["IC"][0]
    for i, sig in enumerate(public_signals):
        vkxpia = G1.add(vkxpia, G1.mul_scalar(vk["IC"][i + 1], sig))
    if not F12.equal(BN.pairing(G1.add(vkxpia, proof["PiA"]), proof["PiB"]),
                     F12.mul(BN.pairing(proof["PiH"], vk["Vkz"]), BN.pairing(proof["PiC"], G2.G))):
        return False, 4
    pia_pic = G1.add(G1.add(vkxpia, proof["PiA"]), proof["PiC"])
    lhs = F12.mul(BN.pairing(pia_pic, vk["G2Kbg"]), BN.pairing(vk["G1Kbg"], proof["PiB"]))
    if not F12.equal(lhs, BN.pairing(proof["PiKp"], vk["G2Kg"])):
        return False, 5
    return True, 0


# ---------------------------------------------------- generic helper (MSM)
def msm_reference_order(group, points, scalars):
    """The reference has no MSM routine; this is its hot loop shape
    (groth16.go:243-250): acc = Add(acc, MulScalar(P_i, s_i)) in index order."""
    acc = group.zero3()
    for p, s in zip(points, scalars):
        acc = group.add(acc, group.mul_scalar(p, s))
    return acc
