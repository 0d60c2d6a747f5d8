#!/usr/bin/env python3
"""Generates tests/golden/kat_v1.json from the big-int oracle (oracle/*.py).

The reference pins no vectors for this path and cannot be run here (Go, absent
gnark-crypto), so these are SELF-GENERATED known-answer vectors: they freeze the oracle's
conventions (Montgomery layout, DIF/DIT ordering, coset handling, computeH, Groth16 cubic with fixed
toxic waste) so that neither the oracle nor the CUDA path can drift silently.  Values are hex strings of
canonical integers; points are affine [x, y] (Fp2 coordinates as [a0, a1]).
    python tests/golden/make_golden.py
"""
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import derive, ec, ff, groth16 as g16, ntt  # noqa: E402
from oracle.params import CURVES  # noqa: E402


def hx(v):
    if isinstance(v, tuple):
        return [hx(x) for x in v]
    return hex(v)


def main():
    out = {"version": 1, "curves": {}}
    for c in CURVES.values():
        rng = random.Random(0xC0FFEE + c.curve_id)
        e = {}
        # field products (canonical a, b, a*b mod q) for Fp and Fr
        e["fp_mul"] = [[hex(a), hex(b), hex(a * b % c.p)] for a, b in ((rng.randrange(c.p), rng.randrange(c.p)) for _ in range(4))]
        e["fr_mul"] = [[hex(a), hex(b), hex(a * b % c.r)] for a, b in ((rng.randrange(c.r), rng.randrange(c.r)) for _ in range(4))]
        # small MSMs on both groups (bases = k_i * B, B of order r)
        for group in (1, 2):
            F = ff.base_field(c, group)
            B = derive.subgroup_point(c, group)
            ks = [rng.randrange(1, c.r) for _ in range(6)]
            pts = [ec.scalar_mul(F, k, B) for k in ks]
            sc = [rng.randrange(c.r) for _ in range(6)]
            sc[0], sc[1] = 0, c.r - 1
            res = ec.msm_naive(F, pts, sc)
            e[f"msm_g{group}"] = {"points": [hx(p) for p in pts], "scalars": [hex(s) for s in sc], "result": hx(res)}
        # NTT, n = 8, all modes
        dom = ntt.Domain(c, 8)
        a = [rng.randrange(c.r) for _ in range(8)]
        e["ntt8"] = {"input": [hex(x) for x in a], "generator": hex(dom.generator), "coset_gen": hex(dom.coset_gen), "out": {}}
        for inv in (0, 1):
            for dec in (0, 1):
                for cos in (0, 1):
                    r_ = (dom.fft_inverse if inv else dom.fft)(a, dec, on_coset=bool(cos))
                    e["ntt8"]["out"][f"inv{inv}_dec{dec}_coset{cos}"] = [hex(x) for x in r_]
        # Groth16 cubic (x = 3, y = 35) with fixed toxic waste and r, s: h and the proof's discrete logs
        cs = g16.cubic_r1cs()
        W = g16.cubic_witness(c.r)
        tox = g16.random_toxic(c, 42)
        pk = g16.setup_dlog(c, cs, tox)
        pr = g16.prove_dlog(c, cs, pk, W, 0x1111, 0x2222)
        assert g16.verify_dlog(c, cs, pk, pr, W)
        e["groth16_cubic"] = {"toxic": [hex(v) for v in (tox.t, tox.alpha, tox.beta, tox.gamma, tox.delta)],
                              "r": hex(0x1111), "s": hex(0x2222), "h_bitreversed": [hex(v) for v in pr.h],
                              "ar": hex(pr.ar), "bs": hex(pr.bs), "krs": hex(pr.krs),
                              "msm": [hex(v) for v in (pr.msm_a, pr.msm_b, pr.msm_z, pr.msm_k)]}
        out["curves"][c.name] = e
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "kat_v1.json")
    json.dump(out, open(dst, "w"), indent=1)
    print("wrote", dst)


if __name__ == "__main__":
    main()
