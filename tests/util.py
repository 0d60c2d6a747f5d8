"""Shared helpers for the parity tests (oracle side)."""
import random

import numpy as np

from oracle import corelib, ec, ff
from oracle.params import CURVES


def rng_for(*key):
    return random.Random(hash(key) & 0xFFFFFFFF)


def pick_base(curve, group, rng=None):
    """A base point of prime order r: gnark-crypto's public generator where recalled, else a
    derived order-r point on an a = 0 curve over the same field (oracle/derive.py) - the group
    law never uses b, so the arithmetic exercised is identical."""
    from oracle import derive
    return ff.base_field(curve, group), derive.subgroup_point(curve, group)


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


def build_groth16_pk(curve, r1cs, toxic, seed=1):
    """Trapdoor Setup (oracle/groth16.py, restating backend/groth16/bn254/setup.go:75-331):
    every pk point = dlog * base, materialised with the C++ oracle's fixed-base batch.
    Returns (gnark_b200.groth16.ProvingKey, ProvingKeyDlog, (F1, g1), (F2, g2))."""
    from gnark_b200 import groth16 as b200_g16
    from oracle import groth16 as g16
    rng = random.Random(seed)
    F1, g1 = pick_base(curve, 1, rng)
    F2, g2 = pick_base(curve, 2, rng)
    pkd = g16.setup_dlog(curve, r1cs, toxic)

    def pts(group, base, dlogs):
        if len(dlogs) == 0:
            deg = 1 if group == 1 else curve.g2_degree
            return np.zeros((0, 2 * deg * curve.fp_limbs), dtype=np.uint64)
        ks = ff.pack_elements(dlogs, curve.r, curve.fr_limbs)
        return corelib.fixed_base(curve, group, ec.pack_points(curve, group, [base]), ks)

    g1s = pts(1, g1, [pkd.alpha, pkd.beta, pkd.delta])
    g2s = pts(2, g2, [pkd.beta, pkd.delta])
    pk = b200_g16.ProvingKey.from_arrays(
        curve.curve_id, pkd.domain.n, g1s[0], g1s[1], g1s[2],
        pts(1, g1, pkd.A), pts(1, g1, pkd.B), pts(1, g1, pkd.Z), pts(1, g1, pkd.K),
        g2s[0], g2s[1], pts(2, g2, pkd.B),
        np.array(pkd.infinity_a, dtype=np.uint8), np.array(pkd.infinity_b, dtype=np.uint8), r1cs.nb_public)
    return pk, pkd, (F1, g1), (F2, g2)


def pack_solution(curve, r1cs, W):
    from gnark_b200 import groth16 as b200_g16
    from oracle import groth16 as g16
    A, B, C = g16.solve_abc(r1cs, W, curve.r)
    p = lambda v: ff.pack_elements(v, curve.r, curve.fr_limbs)
    return b200_g16.R1CSSolution(W=p(W), A=p(A), B=p(B), C=p(C))
