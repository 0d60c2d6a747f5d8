"""EXTERNAL known answers, part 2: elliptic-curve constants the reference spells out in its own sources (computed by
gnark-crypto, pasted into gnark; extracted to tests/golden/gnark_intree_points_v1.json by
tests/golden/make_golden_intree_points.py).  One scalar multiplication each, on every curve of this path:

  G1, all four curves:  [lambda] P = (omega * x_P, y_P)   for every P of order r (the GLV endomorphism;
                        std/algebra/emulated/sw_emulated/params.go:70-71,88-89,157-158,
                        std/algebra/native/sw_bls12377/inner.go:58-63)
  G2, BN254 / BLS12-381 / BW6-761:  [2^k] G2 = g2GenNbits, k = 65 / 65 / 96
                        (std/algebra/emulated/sw_bn254/g2.go:75-96, sw_bls12381/g2.go:89-110, sw_bw6761/g2.go:90-99)

Each is checked as stated on the big-int oracle, and folded into an N-point MSM for the C++ oracle and the device
templates (bases k_i * P, scalars s_i with sum s_i k_i = lambda resp. 2^k mod r): the ANSWER is external, the bases are
multiples computed by the big-int oracle.  CUDA: tests/test_gpu_zz_late.py::test_cuda_reproduces_intree_known_answers."""
import ctypes
import json
import os
import random

import numpy as np
import pytest

from oracle import corelib, derive, ec, ff
from oracle.params import CURVES

KAT = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "gnark_intree_points_v1.json")))["curves"]
P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
CASES = [(name, group) for name, e in KAT.items() for group in (1, 2) if group == 1 or "g2" in e]


def known_answer(c, group):
    """-> (F, base point P, scalar k, expected k * P) with `expected` taken from the reference's constants"""
    e = KAT[c.name]
    F = ff.base_field(c, group)
    if group == 1:
        lam, om = int(e["glv"]["lambda"]), int(e["glv"]["omega"])
        base = derive.subgroup_point(c, 1)
        return F, base, lam, (om * base[0] % c.p, base[1])
    g = e["g2"]
    n = [int(x) for x in g["g2Gen"]]
    m = [int(x) for x in g["g2GenNbits"]]
    pt = (lambda v: ((v[0], v[1]), (v[2], v[3]))) if len(n) == 4 else (lambda v: (v[0], v[1]))
    return F, pt(n), 1 << g["log2_multiple"], pt(m)


def folded_msm(c, group, n, seed):
    """bases k_i * P and scalars s_i with sum s_i k_i = k (mod r): an n-point MSM whose answer is the external point"""
    F, base, k, expected = known_answer(c, group)
    rng = random.Random(seed)
    ks = [rng.randrange(1, 1 << 16) for _ in range(n)]
    pts = [ec.scalar_mul(F, ki, base) for ki in ks]
    sc = [rng.randrange(c.r) for _ in range(n - 1)]
    partial = sum(s * ki for s, ki in zip(sc, ks)) % c.r
    sc.append((k - partial) * pow(ks[-1], -1, c.r) % c.r)
    return F, pts, sc, expected


@pytest.mark.parametrize("cname,group", CASES)
def test_oracle_reproduces_intree_constants(cname, group):
    c = CURVES[cname]
    F, base, k, expected = known_answer(c, group)
    assert ec.scalar_mul(F, c.r, base) is None                       # order r
    assert ec.scalar_mul(F, k, base) == expected
    if group == 2:
        assert base == c.g2                                          # the generator oracle/params.py carries
    else:
        lam, om = k, int(KAT[cname]["glv"]["omega"])
        assert (lam * lam + lam + 1) % c.r == 0 and pow(om, 3, c.p) == 1 and om != 1
    # C++ oracle: one-point MSM and the folded 24-point MSM
    out = corelib.msm(c, group, ec.pack_points(c, group, [base]), ff.pack_elements([k], c.r, c.fr_limbs))
    assert ec.from_jac(F, ec.unpack_points(c, group, out, ncoords=3)[0]) == expected
    F, pts, sc, expected = folded_msm(c, group, 24, 7)
    assert ec.msm_naive(F, pts, sc) == expected
    out = corelib.msm(c, group, ec.pack_points(c, group, pts), ff.pack_elements(sc, c.r, c.fr_limbs), c=5)
    assert ec.from_jac(F, ec.unpack_points(c, group, out, ncoords=3)[0]) == expected


@pytest.mark.parametrize("cname,group", CASES)
def test_device_templates_reproduce_intree_constants(hostemu, cname, group):
    """the CUDA kernels' per-thread code compiled for the host (default and A/B arithmetic builds)"""
    c = CURVES[cname]
    F, pts, sc, expected = folded_msm(c, group, 12 if c.fp_limbs > 6 else 24, 11)
    PA, SA = ec.pack_points(c, group, pts), ff.pack_elements(sc, c.r, c.fr_limbs)
    for lib, (cw, pre, tl, ch) in ((hostemu, (6, 0, 3, 8)), (hostemu, (5, 1, 2, 4)), (hostemu, (6, 0, 4, 8))):
        out = np.zeros(3 * F.degree * c.fp_limbs, dtype=np.uint64)
        assert lib.emu_msm(c.curve_id, group, P(PA), P(SA), len(pts), cw, pre, tl, ch, P(out)) == 0
        assert ec.from_jac(F, ec.unpack_points(c, group, out, ncoords=3)[0]) == expected, (cname, group, cw, pre)
