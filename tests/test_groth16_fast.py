"""oracle/groth16_fast.py (the full-size Groth16 fixture: vectorised satisfied circuit, trapdoor key as discrete-log
arrays, expected MSM results as dot products) against the big-int oracle oracle/groth16.py on the SAME circuit at
small sizes - the fixture that checks the 2^20 proof of BASELINE configs[2] is itself checked here."""
import random

import numpy as np
import pytest

from oracle import corelib, ec, ff
from oracle import groth16 as g16
from oracle import groth16_fast as gf
from oracle.params import CURVES


@pytest.mark.parametrize("cname,logn,m,lanes", [("bn254", 6, 63, 8), ("bn254", 7, 100, 16), ("bls12-381", 5, 32, 4),
                                                ("bw6-761", 4, 13, 2), ("bls12-377", 4, 15, 1024)])
def test_fast_fixture_equals_bigint_oracle(cname, logn, m, lanes):
    c = CURVES[cname]
    r, L = c.r, c.fr_limbs
    inst = gf.satisfied_instance(c, logn, seed=77, m=m, lanes=lanes)
    assert gf.check_satisfied(inst)
    cs = inst.as_r1cs()
    W = inst.witness_ints()
    assert cs.nb_wires == inst.nb_wires == len(W)
    A, B, C = g16.solve_abc(cs, W, r)
    assert all(a * b % r == cc for a, b, cc in zip(A, B, C))
    for got, want in zip(inst.solution_abc(), (A, B, C)):
        assert ff.unpack_elements(got, r, L) == want
    pkd = g16.setup_dlog(c, cs, inst.toxic)
    up = lambda a: ff.unpack_elements(a, r, L)
    assert up(inst.a_dl) == pkd.A and up(inst.b_dl) == pkd.B and up(inst.k_dl) == pkd.K and up(inst.z_dl) == pkd.Z
    assert inst.vk_k == pkd.vk_K
    assert list(inst.inf_a) == [int(x) for x in pkd.infinity_a] and list(inst.inf_b) == [int(x) for x in pkd.infinity_b]
    rng = random.Random(5)
    rr, ss = rng.randrange(r), rng.randrange(r)
    want = g16.prove_dlog(c, cs, pkd, W, rr, ss)
    got = gf.expected(inst, rr, ss)
    assert (got.msm_a, got.msm_b, got.msm_z, got.msm_k) == (want.msm_a, want.msm_b, want.msm_z, want.msm_k)
    assert (got.ar, got.bs, got.krs) == (want.ar, want.bs, want.krs)
    assert g16.verify_dlog(c, cs, pkd, want, W) and gf.verify_in_exponent(inst, got)
    assert ff.unpack_elements(gf.compute_h(inst), r, L) == want.h
    # a wrong proof element breaks the relation
    bad = gf.Expected(got.msm_a, got.msm_b, got.msm_z, got.msm_k, got.ar, got.bs, (got.krs + 1) % r)
    assert not gf.verify_in_exponent(inst, bad)


def test_fast_fixture_points_and_pairing():
    """key points from the C++ oracle's fixed-base batch, the proof assembled from CPU MSMs (corelib.msm), checked by
    verify_points: dlog * G and the verifier's pairing equation - the path the GPU test takes, without a GPU"""
    c = CURVES["bn254"]
    r, L = c.r, c.fr_limbs
    inst = gf.satisfied_instance(c, 5, seed=3, m=31, lanes=4)
    F1, F2 = ff.Fp(c.p), ff.base_field(c, 2)
    fb = lambda group, dl: corelib.fixed_base(c, group, ec.pack_points(c, group, [c.g1 if group == 1 else c.g2]), dl)
    kp = gf.key_points(inst, fb)
    m, S, n = inst.m, inst.lanes, inst.n
    h = gf.compute_h(inst)
    aff = lambda group, jac: ec.from_jac(ff.base_field(c, group), ec.unpack_points(c, group, np.ascontiguousarray(jac), ncoords=3)[0])
    va, vb, vk = (np.ascontiguousarray(inst.v[:k]) for k in (m, m - 1, m + S - 1))
    P_a = aff(1, corelib.msm(c, 1, kp["A"], va))
    P_b1 = aff(1, corelib.msm(c, 1, kp["B"], vb))
    P_b2 = aff(2, corelib.msm(c, 2, kp["B2"], vb))
    P_k = aff(1, corelib.msm(c, 1, kp["K"], vk))
    P_z = aff(1, corelib.msm(c, 1, kp["Z"], np.ascontiguousarray(h[:n - 1])))
    rr, ss = 12345, 67890
    e = gf.expected(inst, rr, ss)
    assert P_a == ec.scalar_mul(F1, e.msm_a, c.g1) and P_b2 == ec.scalar_mul(F2, e.msm_b, c.g2)
    assert P_k == ec.scalar_mul(F1, e.msm_k, c.g1) and P_z == ec.scalar_mul(F1, e.msm_z, c.g1)
    tox = inst.toxic
    add1 = lambda P, Q: ec.affine_add(F1, P, Q)
    sm1 = lambda k, P: ec.scalar_mul(F1, k % r, P)
    delta1, delta2 = sm1(tox.delta, c.g1), ec.scalar_mul(F2, tox.delta, c.g2)
    ar = add1(add1(P_a, sm1(tox.alpha, c.g1)), sm1(rr, delta1))
    bs1 = add1(add1(P_b1, sm1(tox.beta, c.g1)), sm1(ss, delta1))
    bs = ec.affine_add(F2, ec.affine_add(F2, P_b2, ec.scalar_mul(F2, tox.beta, c.g2)), ec.scalar_mul(F2, ss, delta2))
    krs = add1(add1(add1(P_k, P_z), sm1(-rr * ss, delta1)), add1(sm1(ss, ar), sm1(rr, bs1)))
    assert gf.verify_points(inst, ar, bs, krs, e, with_pairing=True)
    assert not gf.verify_points(inst, ar, bs, add1(krs, c.g1), e, with_pairing=False)
