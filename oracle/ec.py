"""Short-Weierstrass (a = 0) group law over Fp / Fp2, naive MSM, and point
(de)serialisation in gnark memory layout.  TEST INFRASTRUCTURE ONLY
(see oracle/params.py header for what pins parity and what does not).

What the reference relies on:
  MultiExp(points, scalars) = sum_i scalars[i] * points[i] as a group element
  (call sites backend/groth16/bn254/prove.go:194,207,227,237,283); results are
  only consumed through AddMixed/AddAssign/FromJacobian (:199-214,241-269), so
  the affine (x, y) of the sum is the bit-exact comparison point.
  Jacobian convention x = X/Z^2, y = Y/Z^3
  (backend/accelerated/icicle/groth16/bn254/icicle.go:266-315).
"""

import numpy as np

from . import ff

INF = None  # affine infinity inside the oracle (serialised as (0, 0))


# --- affine group law (a = 0; the curve coefficient b never enters) ---------

def affine_add(F, P, Q):
    if P is INF:
        return Q
    if Q is INF:
        return P
    x1, y1 = P
    x2, y2 = Q
    if F.eq(x1, x2):
        if F.eq(y1, y2):
            if F.is_zero(y1):
                return INF
            lam = F.mul(F.mul(F.from_int(3), F.sqr(x1)), F.inv(F.add(y1, y1)))
        else:
            return INF
    else:
        lam = F.mul(F.sub(y2, y1), F.inv(F.sub(x2, x1)))
    x3 = F.sub(F.sub(F.sqr(lam), x1), x2)
    y3 = F.sub(F.mul(lam, F.sub(x1, x3)), y1)
    return (x3, y3)


def affine_neg(F, P):
    if P is INF:
        return INF
    return (P[0], F.neg(P[1]))


# --- Jacobian (used for speed in scalar multiplication) ---------------------

def jac_double(F, P):
    X, Y, Z = P
    if F.is_zero(Z) or F.is_zero(Y):
        return (F.one, F.one, F.zero)
    A = F.sqr(X)
    B = F.sqr(Y)
    C = F.sqr(B)
    t = F.sub(F.sub(F.sqr(F.add(X, B)), A), C)
    D = F.add(t, t)
    E = F.add(F.add(A, A), A)
    Fq = F.sqr(E)
    X3 = F.sub(Fq, F.add(D, D))
    C8 = F.add(C, C); C8 = F.add(C8, C8); C8 = F.add(C8, C8)
    Y3 = F.sub(F.mul(E, F.sub(D, X3)), C8)
    Z3 = F.mul(F.add(Y, Y), Z)
    return (X3, Y3, Z3)


def jac_add(F, P, Q):
    X1, Y1, Z1 = P
    X2, Y2, Z2 = Q
    if F.is_zero(Z1):
        return Q
    if F.is_zero(Z2):
        return P
    Z1Z1 = F.sqr(Z1)
    Z2Z2 = F.sqr(Z2)
    U1 = F.mul(X1, Z2Z2)
    U2 = F.mul(X2, Z1Z1)
    S1 = F.mul(F.mul(Y1, Z2), Z2Z2)
    S2 = F.mul(F.mul(Y2, Z1), Z1Z1)
    if F.eq(U1, U2):
        if F.eq(S1, S2):
            return jac_double(F, P)
        return (F.one, F.one, F.zero)
    H = F.sub(U2, U1)
    R = F.sub(S2, S1)
    HH = F.sqr(H)
    HHH = F.mul(H, HH)
    V = F.mul(U1, HH)
    X3 = F.sub(F.sub(F.sqr(R), HHH), F.add(V, V))
    Y3 = F.sub(F.mul(R, F.sub(V, X3)), F.mul(S1, HHH))
    Z3 = F.mul(F.mul(Z1, Z2), H)
    return (X3, Y3, Z3)


def to_jac(F, P):
    if P is INF:
        return (F.one, F.one, F.zero)
    return (P[0], P[1], F.one)


def from_jac(F, P):
    X, Y, Z = P
    if F.is_zero(Z):
        return INF
    zi = F.inv(Z)
    zi2 = F.sqr(zi)
    return (F.mul(X, zi2), F.mul(F.mul(Y, zi2), zi))


def scalar_mul(F, k: int, P):
    """k * P for an arbitrary non-negative integer k (no reduction mod r: valid
    for points outside the r-torsion too, which the known-dlog checks use)."""
    if P is INF or k == 0:
        return INF
    if k < 0:
        return scalar_mul(F, -k, affine_neg(F, P))
    acc = (F.one, F.one, F.zero)
    base = to_jac(F, P)
    for bit in bin(k)[2:]:
        acc = jac_double(F, acc)
        if bit == "1":
            acc = jac_add(F, acc, base)
    return from_jac(F, acc)


def msm_naive(F, points, scalars):
    """sum_i scalars[i] * points[i] by double-and-add (small cases only)."""
    acc = (F.one, F.one, F.zero)
    for P, s in zip(points, scalars):
        if P is INF or s == 0:
            continue
        acc = jac_add(F, acc, to_jac(F, scalar_mul(F, s, P)))
    return from_jac(F, acc)


def is_on_curve(F, P, b):
    if P is INF:
        return True
    x, y = P
    return F.eq(F.sqr(y), F.add(F.mul(F.sqr(x), x), b))


def curve_b_of(F, P):
    """The b for which P lies on y^2 = x^3 + b.  a = 0 formulas never use b, so
    any (x, y) with y != 0 generates a valid test group on *its own* curve."""
    x, y = P
    return F.sub(F.sqr(y), F.mul(F.sqr(x), x))


# --- serialisation (gnark memory layout) ------------------------------------

def pack_points(curve, group: int, points) -> np.ndarray:
    """affine points -> (n, 2*deg*fp_limbs) uint64, Montgomery, infinity=(0,0)."""
    F = ff.base_field(curve, group)
    deg = F.degree
    L = curve.fp_limbs
    flat = []
    for P in points:
        if P is INF:
            flat.extend([0] * (2 * deg))
        else:
            flat.extend(F.coords(P[0]))
            flat.extend(F.coords(P[1]))
    arr = ff.pack_elements(flat, curve.p, L, mont=True)
    return arr.reshape(len(points), 2 * deg * L)


def unpack_points(curve, group: int, arr, ncoords: int = 2):
    """inverse of pack_points; ncoords=3 reads Jacobian {X,Y,Z} triples."""
    F = ff.base_field(curve, group)
    deg = F.degree
    L = curve.fp_limbs
    vals = ff.unpack_elements(arr, curve.p, L, mont=True)
    per = ncoords * deg
    out = []
    for i in range(len(vals) // per):
        c = vals[i * per:(i + 1) * per]
        cs = [F.from_coords(c[j * deg:(j + 1) * deg]) for j in range(ncoords)]
        if ncoords == 2:
            out.append(INF if (F.is_zero(cs[0]) and F.is_zero(cs[1])) else (cs[0], cs[1]))
        else:
            out.append(tuple(cs))
    return out
