"""Committed known-answer vectors (tests/golden/kat_v1.json, made by tests/golden/make_golden.py from the
big-int oracle): the oracle and the C++ oracle must reproduce them on the CPU, the CUDA path on the GPU."""
import json
import os

import numpy as np
import pytest

from oracle import corelib, ec, ff, groth16 as g16, ntt
from oracle.params import CURVES

KAT = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "kat_v1.json")))
ALL = list(CURVES.values())
H = lambda s: int(s, 16)


def pt(F, v):
    if v is None:
        return None
    conv = (lambda x: (H(x[0]), H(x[1]))) if F.degree == 2 else H
    return (conv(v[0]), conv(v[1]))


@pytest.mark.parametrize("c", ALL, ids=lambda c: c.name)
def test_oracles_reproduce_golden(c):
    e = KAT["curves"][c.name]
    for a, b, ab in e["fp_mul"]:
        assert H(a) * H(b) % c.p == H(ab)
    for group in (1, 2):
        F = ff.base_field(c, group)
        m = e[f"msm_g{group}"]
        pts = [pt(F, p) for p in m["points"]]
        sc = [H(s) for s in m["scalars"]]
        want = pt(F, m["result"])
        assert ec.msm_naive(F, pts, sc) == want
        got = corelib.msm(c, group, ec.pack_points(c, group, pts), ff.pack_elements(sc, c.r, c.fr_limbs), c=4)
        assert ec.from_jac(F, ec.unpack_points(c, group, got, ncoords=3)[0]) == want
    n8 = e["ntt8"]
    dom = ntt.Domain(c, 8)
    assert dom.generator == H(n8["generator"]) and dom.coset_gen == H(n8["coset_gen"])
    a = [H(x) for x in n8["input"]]
    for key, out in n8["out"].items():
        inv, dec, cos = int(key[3]), int(key[8]), int(key[-1])
        assert (dom.fft_inverse if inv else dom.fft)(a, dec, on_coset=bool(cos)) == [H(x) for x in out]
        A = ff.pack_elements(a, c.r, c.fr_limbs)
        corelib.ntt(c, A, 3, inv, dec, cos)
        assert ff.unpack_elements(A, c.r, c.fr_limbs) == [H(x) for x in out]
    gc = e["groth16_cubic"]
    cs, W = g16.cubic_r1cs(), g16.cubic_witness(c.r)
    pk = g16.setup_dlog(c, cs, g16.Toxic(*[H(v) for v in gc["toxic"]]))
    pr = g16.prove_dlog(c, cs, pk, W, H(gc["r"]), H(gc["s"]))
    assert [pr.ar, pr.bs, pr.krs] == [H(gc["ar"]), H(gc["bs"]), H(gc["krs"])]
    assert pr.h == [H(v) for v in gc["h_bitreversed"]]


PLONK_KAT = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "kat_plonk_v1.json")))


def _plonk_case(cname):
    from oracle import plonk_prover as pp
    c = CURVES[cname]
    e = PLONK_KAT["curves"][cname]
    HL = lambda v: [H(x) for x in v]
    k = e["circuit"]
    circ = pp.Circuit(n=e["n"], ql=HL(k["ql"]), qr=HL(k["qr"]), qm=HL(k["qm"]), qo=HL(k["qo"]), qk=HL(k["qk"]),
                      perm=list(k["perm"]), qcp=[HL(v) for v in k["qcp"]])
    w = e["witness"]
    x = e["challenges"]
    ch = pp.Challenges(gamma=H(x["gamma"]), beta=H(x["beta"]), alpha=H(x["alpha"]), zeta=H(x["zeta"]), v=H(x["v"]),
                       bl=HL(x["bl"]), br=HL(x["br"]), bo=HL(x["bo"]), bz=HL(x["bz"]))
    return c, circ, HL(w["l"]), HL(w["r"]), HL(w["o"]), [HL(v) for v in w["pi2"]], ch, H(e["tau"]), e["proof"]


@pytest.mark.parametrize("cname", ["bn254", "bls12-381"])
def test_oracle_prover_reproduces_plonk_golden(cname):
    """the oracle PLONK prover (with one BSB22 commitment gate) on the committed instance: every digest and opened value"""
    from oracle import plonk_prover as pp
    c, circ, l, rr, o, pi2, ch, tau, want = _plonk_case(cname)
    pr = pp.prove(c, circ, l, rr, o, ch, tau, pi2=pi2)
    assert pp.verify(c, circ, pr, ch, tau)
    HL = lambda v: [H(x) for x in v]
    assert [pr.L, pr.R, pr.O, pr.Z] == [H(want[k]) for k in ("L", "R", "O", "Z")]
    assert pr.H == HL(want["H"]) and pr.lin == H(want["lin"])
    assert pr.batch_opening == H(want["batch_opening"]) and pr.z_opening == H(want["z_opening"])
    assert pr.bsb22 == HL(want["bsb22"]) and pr.claimed == HL(want["claimed"]) and pr.zu == H(want["zu"])
