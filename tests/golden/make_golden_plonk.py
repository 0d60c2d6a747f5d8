#!/usr/bin/env python3
"""Generates tests/golden/kat_plonk_v1.json from the big-int oracle prover (oracle/plonk_prover.py): one PLONK
proof per curve (BN254, BLS12-381) on an 8-row circuit with one BSB22 commitment gate, fixed challenges, blinding
and SRS trapdoor.  SELF-GENERATED known-answer vectors (the reference pins none for this path):
they freeze the prover's conventions - blinding placement, coset order, quotient split, linearised polynomial,
opening folds - so that neither the oracle nor the device orchestrations can drift silently.  Digests are stored
as discrete logs w.r.t. G1 (digest = dlog * G), values as canonical integers, all hex.
    python tests/golden/make_golden_plonk.py
"""
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import plonk_prover as pp  # noqa: E402
from oracle.params import CURVES  # noqa: E402


def build(c):
    rng = random.Random(0xB10C + c.curve_id)
    n = 8
    circ, l, rr, o, pi2 = pp.random_satisfied_instance(c, n, seed=1234 + c.curve_id, n_commit=1)
    rnd = lambda: rng.randrange(c.r)
    ch = pp.Challenges(gamma=rnd(), beta=rnd(), alpha=rnd(), zeta=rnd(), v=rnd(), bl=[rnd(), rnd()], br=[rnd(), rnd()],
                       bo=[rnd(), rnd()], bz=[rnd(), rnd(), rnd()])
    tau = rnd()
    proof = pp.prove(c, circ, l, rr, o, ch, tau, pi2=pi2)
    assert pp.verify(c, circ, proof, ch, tau)
    H = lambda v: [hex(x) for x in v]
    return {
        "n": n, "tau": hex(tau),
        "circuit": {"ql": H(circ.ql), "qr": H(circ.qr), "qm": H(circ.qm), "qo": H(circ.qo), "qk": H(circ.qk),
                    "perm": circ.perm, "qcp": [H(v) for v in circ.qcp]},
        "witness": {"l": H(l), "r": H(rr), "o": H(o), "pi2": [H(v) for v in pi2]},
        "challenges": {"gamma": hex(ch.gamma), "beta": hex(ch.beta), "alpha": hex(ch.alpha), "zeta": hex(ch.zeta),
                       "v": hex(ch.v), "bl": H(ch.bl), "br": H(ch.br), "bo": H(ch.bo), "bz": H(ch.bz)},
        "proof": {"L": hex(proof.L), "R": hex(proof.R), "O": hex(proof.O), "Z": hex(proof.Z), "H": H(proof.H),
                  "lin": hex(proof.lin), "batch_opening": hex(proof.batch_opening), "z_opening": hex(proof.z_opening),
                  "bsb22": H(proof.bsb22), "claimed": H(proof.claimed), "zu": hex(proof.zu)},
    }


def main():
    out = {"version": 1, "curves": {name: build(CURVES[name]) for name in ("bn254", "bls12-381")}}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "kat_plonk_v1.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=0, sort_keys=True)
        f.write("\n")
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
