"""Big-int field helpers + gnark-crypto memory layout.  TEST INFRASTRUCTURE ONLY
(see oracle/params.py header for what pins parity and what does not).

Layout restated from SURVEY.md Appendix A (evidence:
backend/accelerated/icicle/groth16/bn254/icicle.go:119-130,266-315):
  fr.Element / fp.Element = [Limbs]uint64, little-endian limbs, value x stored
  as x*R mod q with R = 2^(64*Limbs)  (Montgomery form).
  G1Affine = {X, Y};  G2Affine = {X, Y} with E2 = {A0, A1};  G1Jac = {X, Y, Z};
  affine infinity = (0, 0);  Jacobian infinity has Z = 0.
"""

import numpy as np


def mont_r(limbs: int) -> int:
    return 1 << (64 * limbs)


def to_mont(x: int, q: int, limbs: int) -> int:
    return (x << (64 * limbs)) % q


def from_mont(xm: int, q: int, limbs: int) -> int:
    return (xm * pow(mont_r(limbs), -1, q)) % q


def int_to_limbs(x: int, limbs: int) -> np.ndarray:
    return np.frombuffer(int(x).to_bytes(8 * limbs, "little"), dtype=np.uint64).copy()


def limbs_to_int(a) -> int:
    return int.from_bytes(np.ascontiguousarray(a, dtype=np.uint64).tobytes(), "little")


def pack_elements(vals, q: int, limbs: int, mont: bool = True) -> np.ndarray:
    """list of ints -> (n, limbs) uint64 array in gnark memory layout."""
    n = len(vals)
    buf = bytearray(8 * limbs * n)
    R = mont_r(limbs)
    nb = 8 * limbs
    for i, v in enumerate(vals):
        v %= q
        if mont:
            v = (v * R) % q
        buf[i * nb:(i + 1) * nb] = v.to_bytes(nb, "little")
    return np.frombuffer(bytes(buf), dtype=np.uint64).reshape(n, limbs).copy()


def unpack_elements(arr, q: int, limbs: int, mont: bool = True):
    """(n, limbs) uint64 array (or flat) -> list of ints (canonical values)."""
    raw = np.ascontiguousarray(arr, dtype=np.uint64).tobytes()
    nb = 8 * limbs
    n = len(raw) // nb
    out = []
    if mont:
        rinv = pow(mont_r(limbs), -1, q)
        for i in range(n):
            out.append((int.from_bytes(raw[i * nb:(i + 1) * nb], "little") * rinv) % q)
    else:
        for i in range(n):
            out.append(int.from_bytes(raw[i * nb:(i + 1) * nb], "little"))
    return out


# ---------------------------------------------------------------------------
# Field objects (plain Python ints for Fp; tuples (a0, a1) for Fp2)
# ---------------------------------------------------------------------------

class Fp:
    degree = 1

    def __init__(self, p):
        self.p = p
        self.zero = 0
        self.one = 1

    def add(self, a, b): return (a + b) % self.p
    def sub(self, a, b): return (a - b) % self.p
    def neg(self, a): return (-a) % self.p
    def mul(self, a, b): return (a * b) % self.p
    def sqr(self, a): return (a * a) % self.p
    def inv(self, a): return pow(a, -1, self.p)
    def is_zero(self, a): return a % self.p == 0
    def eq(self, a, b): return (a - b) % self.p == 0
    def from_int(self, k): return k % self.p
    def coords(self, a): return [a % self.p]
    def from_coords(self, c): return c[0] % self.p


class Fp2:
    """Fp[u]/(u^2 - beta).  Multiplication rule restated from
    std/algebra/emulated/fields_bn254/e2.go:203-213 (beta=-1: b0=x0y0-x1y1,
    b1=x0y1+x1y0) and std/algebra/native/fields_bls12377/e2.go:134 (beta=-5)."""
    degree = 2

    def __init__(self, p, beta):
        self.p = p
        self.beta = beta % p
        self.zero = (0, 0)
        self.one = (1, 0)

    def add(self, a, b): return ((a[0] + b[0]) % self.p, (a[1] + b[1]) % self.p)
    def sub(self, a, b): return ((a[0] - b[0]) % self.p, (a[1] - b[1]) % self.p)
    def neg(self, a): return ((-a[0]) % self.p, (-a[1]) % self.p)

    def mul(self, a, b):
        p = self.p
        return ((a[0] * b[0] + self.beta * a[1] * b[1]) % p, (a[0] * b[1] + a[1] * b[0]) % p)

    def sqr(self, a): return self.mul(a, a)

    def inv(self, a):
        p = self.p
        n = (a[0] * a[0] - self.beta * a[1] * a[1]) % p
        ni = pow(n, -1, p)
        return ((a[0] * ni) % p, (-a[1] * ni) % p)

    def is_zero(self, a): return a[0] % self.p == 0 and a[1] % self.p == 0
    def eq(self, a, b): return self.is_zero(self.sub(a, b))
    def from_int(self, k): return (k % self.p, 0)
    def coords(self, a): return [a[0] % self.p, a[1] % self.p]
    def from_coords(self, c): return (c[0] % self.p, c[1] % self.p)


def base_field(curve, group: int):
    """Field the coordinates of G<group> live in."""
    if group == 1 or curve.fp2_nonresidue is None:
        return Fp(curve.p)
    return Fp2(curve.p, curve.fp2_nonresidue)
