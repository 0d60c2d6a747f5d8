"""Order-r test points for the groups whose gnark-crypto generator is not recalled
(BLS12-377 G2, BW6-761 G1/G2).  TEST INFRASTRUCTURE ONLY (see oracle/params.py).

The a = 0 group law never uses the curve coefficient b, so ANY point of prime order r
on ANY curve y^2 = x^3 + b' over the coordinate field exercises exactly the same
arithmetic as the real generator; the Groth16 dlog checks (which reduce exponents mod
r) need nothing more.  Method: j = 0 curves over F_q have one of six group orders
q + 1 - t, with t read off a decomposition q = N(pi), pi in Z[sqrt(-3)]:
    p = x^2 + 3 y^2  (Cornacchia)  ->  t in {+-2x, +-(x+3y), +-(x-3y)}           (q = p)
    pi^2 = (x^2-3y^2) + 2xy sqrt(-3) -> t in {+-2u, +-(u+3v), +-(u-3v)}, u=x^2-3y^2, v=2xy  (q = p^2)
A random (X, Y) fixes b' = Y^2 - X^3; the candidate order that annihilates it is the
order of its curve; when r divides it, (order / r) * (X, Y) has order r.

Run as a script to (re)generate oracle/derived_points.json (committed).
"""

import json
import math
import os
import random

from . import ec, ff
from .params import CURVES

_HERE = os.path.dirname(os.path.abspath(__file__))
CACHE = os.path.join(_HERE, "derived_points.json")


def _sqrt_mod(a, p):
    """Tonelli-Shanks."""
    a %= p
    if a == 0:
        return 0
    assert pow(a, (p - 1) // 2, p) == 1, "not a square"
    if p % 4 == 3:
        return pow(a, (p + 1) // 4, p)
    q, s = p - 1, 0
    while q % 2 == 0:
        q //= 2
        s += 1
    z = 2
    while pow(z, (p - 1) // 2, p) != p - 1:
        z += 1
    m, c, t, r = s, pow(z, q, p), pow(a, q, p), pow(a, (q + 1) // 2, p)
    while t != 1:
        i, t2 = 0, t
        while t2 != 1:
            t2 = t2 * t2 % p
            i += 1
        b = pow(c, 1 << (m - i - 1), p)
        m, c, t, r = i, b * b % p, t * b * b % p, r * b % p
    return r


def cornacchia_x2_3y2(p):
    """p = x^2 + 3 y^2 for a prime p = 1 mod 3."""
    r0 = _sqrt_mod(-3, p)
    if r0 * 2 > p:
        r0 = p - r0
    a, b = p, r0
    lim = math.isqrt(p)
    while b > lim:
        a, b = b, a % b
    x = b
    rest = p - x * x
    assert rest % 3 == 0
    y = math.isqrt(rest // 3)
    assert x * x + 3 * y * y == p
    return x, y


def candidate_orders(p, degree):
    x, y = cornacchia_x2_3y2(p)
    if degree == 1:
        q, u, v = p, x, y
    else:
        q, u, v = p * p, x * x - 3 * y * y, 2 * x * y
    ts = [2 * u, -2 * u, u + 3 * v, -(u + 3 * v), u - 3 * v, -(u - 3 * v)]
    return [q + 1 - t for t in ts]


def derive_point(curve, group, seed=1):
    F = ff.base_field(curve, group)
    cands = candidate_orders(curve.p, F.degree)
    rng = random.Random(seed)
    while True:
        X = F.from_coords([rng.randrange(curve.p) for _ in range(F.degree)])
        Y = F.from_coords([rng.randrange(1, curve.p) for _ in range(F.degree)])
        Q = (X, Y)
        for N in cands:
            if N % curve.r == 0 and ec.scalar_mul(F, N, Q) is ec.INF:
                P = ec.scalar_mul(F, N // curve.r, Q)
                if P is not ec.INF:
                    assert ec.scalar_mul(F, curve.r, P) is ec.INF
                    return P


def _key(curve, group):
    return f"{curve.name}:g{group}"


def load_cache():
    if os.path.exists(CACHE):
        return json.load(open(CACHE))
    return {}


def subgroup_point(curve, group):
    """A point of order r for (curve, group): the public generator when known, else derived."""
    F = ff.base_field(curve, group)
    if group == 1 and curve.g1 is not None:
        return curve.g1
    if group == 2 and curve.g2 is not None:
        return curve.g2
    c = load_cache()
    k = _key(curve, group)
    if k in c:
        co = [int(v, 16) for v in c[k]]
        d = F.degree
        return (F.from_coords(co[:d]), F.from_coords(co[d:]))
    return derive_point(curve, group)


def main():
    out = {}
    for curve in CURVES.values():
        for group in (1, 2):
            if (group == 1 and curve.g1 is not None) or (group == 2 and curve.g2 is not None):
                continue
            F = ff.base_field(curve, group)
            P = derive_point(curve, group, seed=0x6E61726B + curve.curve_id * 2 + group)
            out[_key(curve, group)] = [hex(v) for v in F.coords(P[0]) + F.coords(P[1])]
            print(_key(curve, group), "ok")
    json.dump(out, open(CACHE, "w"), indent=1)


if __name__ == "__main__":
    main()
