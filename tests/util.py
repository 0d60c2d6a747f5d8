"""Shared helpers for the parity tests (oracle side)."""
import random

import numpy as np

from oracle import corelib, ec, ff
from oracle.params import CURVES


def rng_for(*key):
    return random.Random(hash(key) & 0xFFFFFFFF)


def pick_base(curve, group, rng):
    """A base point for known-dlog tests.  a = 0 formulas never use b, so any
    (x, y) with y != 0 generates a valid group on its own curve y^2 = x^3 + b'."""
    F = ff.base_field(curve, group)
    if group == 1 and curve.g1 is not None:
        return F, curve.g1
    if group == 2 and curve.g2 is not None:
        return F, curve.g2
    x = F.from_coords([rng.randrange(curve.p) for _ in range(F.degree)])
    y = F.from_coords([rng.randrange(1, curve.p) for _ in range(F.degree)])
    return F, (x, y)


def known_dlog_instance(curve, group, n, seed, skew=False):
    """bases P_i = k_i * B (C++ oracle fixed-base), scalars s_i; expected = (sum s_i k_i) * B."""
    rng = random.Random(seed)
    F, base = pick_base(curve, group, rng)
    ks = [rng.randrange(1, curve.r) for _ in range(n)]
    if skew:
        sc = []
        for _ in range(n):
            u = rng.random()
            sc.append(0 if u < 0.5 else (rng.choice((1, 2)) if u < 0.75 else rng.randrange(curve.r)))
    else:
        sc = [rng.randrange(curve.r) for _ in range(n)]
    KS = ff.pack_elements(ks, curve.r, curve.fr_limbs)
    SC = ff.pack_elements(sc, curve.r, curve.fr_limbs)
    PTS = corelib.fixed_base(curve, group, ec.pack_points(curve, group, [base]), KS)
    total = sum(s * k for s, k in zip(sc, ks))
    expected = ec.scalar_mul(F, total, base)
    return F, base, PTS, SC, expected


def jac_to_affine(curve, group, jac_arr):
    F = ff.base_field(curve, group)
    return ec.from_jac(F, ec.unpack_points(curve, group, jac_arr, ncoords=3)[0])
