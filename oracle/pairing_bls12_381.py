"""Reduced Tate/ate pairing on BLS12-381 in plain big-int Python.  TEST INFRASTRUCTURE ONLY.

Purpose: the reference's acceptance test for this path is a pairing check (`Verify`,
backend/groth16/bls12-381/verify.go, kzg.Verify in backend/plonk/bls12-381/verify.go), and the one external
fixture with G2 points (the Ethereum KZG ceremony SRS, tests/golden/eth_kzg_srs_v1.bin) relates its G1 and G2
parts only through e(tau^k G1, G2) = e(G1, tau^k G2).  With a pairing the oracle's G2 arithmetic is tied to
external data as well (tests/test_golden_kzg.py::test_g2_msm_pinned_by_pairing).

Construction (textbook, written from the algebra; nothing here is performance-minded):
  Fp12 = Fp[w] / (w^12 - 2 w^6 + 2)       (w^6 = 1 + u = xi, u^2 = -1, so (w^6 - 1)^2 = -1)
  untwist  E'(Fp2): y^2 = x^3 + 4 xi  ->  E(Fp12): y^2 = x^3 + 4,   (x', y') -> (x'/w^2, y'/w^3)
  f = f_{|x|, Q}(P) by an affine Miller loop over Fp12 (vertical lines dropped: they live in a proper subfield and
  die in the final exponentiation), |x| = 0xd201000000010000; the sign of x only conjugates the result, which
  cancels in every comparison e(A, B) == e(C, D) made here;   e = f^((p^12 - 1) / r).
Self-checks (bilinearity, non-degeneracy) in tests/test_golden_kzg.py.
"""
from .params import BLS12_381 as C

P = C.p
R = C.r
X_ABS = 0xd201000000010000
FINAL_EXP = (P ** 12 - 1) // R
assert (P ** 12 - 1) % R == 0


class Fp12:
    __slots__ = ("c",)

    def __init__(self, c):
        self.c = c          # 12 coefficients, little-endian in w

    @staticmethod
    def one():
        return Fp12([1] + [0] * 11)

    @staticmethod
    def from_fp(a):
        return Fp12([a % P] + [0] * 11)

    @staticmethod
    def from_fp2(a):
        """a0 + a1 u  with  u = w^6 - 1"""
        a0, a1 = a
        c = [0] * 12
        c[0] = (a0 - a1) % P
        c[6] = a1 % P
        return Fp12(c)

    def __eq__(self, o):
        return self.c == o.c

    def __add__(self, o):
        return Fp12([(a + b) % P for a, b in zip(self.c, o.c)])

    def __sub__(self, o):
        return Fp12([(a - b) % P for a, b in zip(self.c, o.c)])

    def __mul__(self, o):
        t = [0] * 23
        a, b = self.c, o.c
        for i in range(12):
            ai = a[i]
            if ai:
                for j in range(12):
                    t[i + j] += ai * b[j]
        # w^12 = 2 w^6 - 2
        for k in range(22, 11, -1):
            v = t[k]
            if v:
                t[k - 6] += 2 * v
                t[k - 12] -= 2 * v
        return Fp12([x % P for x in t[:12]])

    def is_zero(self):
        return not any(self.c)

    def inv(self):
        """extended Euclid on polynomials over Fp"""
        mod = [2, 0, 0, 0, 0, 0, P - 2, 0, 0, 0, 0, 0, 1]

        def deg(p):
            d = len(p) - 1
            while d >= 0 and p[d] == 0:
                d -= 1
            return d

        def divmod_poly(a, b):
            a = a[:]
            db = deg(b)
            ib = pow(b[db], P - 2, P)
            q = [0] * max(1, len(a))
            while deg(a) >= db:
                da = deg(a)
                f = a[da] * ib % P
                q[da - db] = f
                for i in range(db + 1):
                    a[da - db + i] = (a[da - db + i] - f * b[i]) % P
            return q, a

        def mul_poly(a, b):
            r = [0] * (len(a) + len(b))
            for i, x in enumerate(a):
                if x:
                    for j, y in enumerate(b):
                        r[i + j] = (r[i + j] + x * y) % P
            return r

        def sub_poly(a, b):
            n = max(len(a), len(b))
            a = a + [0] * (n - len(a))
            b = b + [0] * (n - len(b))
            return [(x - y) % P for x, y in zip(a, b)]

        r0, r1 = mod, self.c[:]
        s0, s1 = [0], [1]
        while deg(r1) > 0:
            q, rem = divmod_poly(r0, r1)
            r0, r1 = r1, rem
            s0, s1 = s1, sub_poly(s0, mul_poly(q, s1))
        assert deg(r1) == 0, "not invertible"
        k = pow(r1[0], P - 2, P)
        out = [(x * k) % P for x in s1]
        out = out + [0] * 12
        # reduce (degree can reach 11 only; the Bezout coefficient of self has degree < 12)
        assert not any(out[12:])
        return Fp12(out[:12])

    def __truediv__(self, o):
        return self * o.inv()

    def __pow__(self, e):
        r = Fp12.one()
        b = self
        while e:
            if e & 1:
                r = r * b
            b = b * b
            e >>= 1
        return r


W = Fp12([0, 1] + [0] * 10)
W2_INV = (W * W).inv()
W3_INV = (W * W * W).inv()
FOUR = Fp12.from_fp(4)


def embed_g1(pt):
    return (Fp12.from_fp(pt[0]), Fp12.from_fp(pt[1]))


def untwist(q):
    """G2 point on the twist (Fp2 coordinates, canonical ints) -> E(Fp12)"""
    x, y = q
    X = Fp12.from_fp2(x) * W2_INV
    Y = Fp12.from_fp2(y) * W3_INV
    assert Y * Y == X * X * X + FOUR, "untwisted point is not on y^2 = x^3 + 4"
    return (X, Y)


def _line(T, Q2, Pt):
    """value at Pt of the line through T and Q2 (tangent when equal); returns (value, T + Q2)"""
    x1, y1 = T
    x2, y2 = Q2
    if x1 == x2 and y1 == y2:
        lam = (x1 * x1 * Fp12.from_fp(3)) / (y1 + y1)
    else:
        lam = (y2 - y1) / (x2 - x1)
    x3 = lam * lam - x1 - x2
    y3 = lam * (x1 - x3) - y1
    xp, yp = Pt
    return (yp - y1) - lam * (xp - x1), (x3, y3)


def miller(Pt, Q):
    """f_{|x|, Q}(P) for P in G1 (affine ints), Q in G2 (affine Fp2)"""
    if Pt is None or Q is None:
        return Fp12.one()
    Pe, Qe = embed_g1(Pt), untwist(Q)
    f = Fp12.one()
    T = Qe
    for bit in bin(X_ABS)[3:]:
        l, T = _line(T, T, Pe)
        f = f * f * l
        if bit == "1":
            l, T = _line(T, Qe, Pe)
            f = f * l
    return f


def pairing(Pt, Q):
    return miller(Pt, Q) ** FINAL_EXP


def pairing_product_is_one(pairs):
    """prod e(P_i, Q_i) == 1 with ONE final exponentiation (how Verify checks its equation)"""
    f = Fp12.one()
    for Pt, Q in pairs:
        f = f * miller(Pt, Q)
    return f ** FINAL_EXP == Fp12.one()
