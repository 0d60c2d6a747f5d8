#!/usr/bin/env python3
"""One verified Groth16 proof at 2^logn constraints (oracle/groth16_fast.py instance) and the library's step profile
(GB200_STEP_PROFILE, the twin of the reference's ICICLE_STEP_PROFILE, icicle.go:72-75): per-stage ms on stderr, with the
stages serialised.  Usage: python tools/run_groth16.py [curve] [logn] [proofs]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gnark_b200 import groth16 as g16, lib  # noqa: E402
from oracle import ec, groth16_fast as gf  # noqa: E402
from oracle.params import CURVES  # noqa: E402


def main():
    cname = sys.argv[1] if len(sys.argv) > 1 else "bn254"
    logn = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    c = CURVES[cname]
    lib.load(); lib.init([0])
    inst = gf.satisfied_instance(c, logn, seed=20)
    fb = lambda group, dl: lib.fixed_base_batch(c.curve_id, group, ec.pack_points(c, group, [c.g1 if group == 1 else c.g2]),
                                                np.ascontiguousarray(dl))
    kp = gf.key_points(inst, fb)
    pk = g16.ProvingKey.from_arrays(c.curve_id, inst.n, kp["alpha"], kp["beta"], kp["delta"], kp["A"], kp["B"], kp["Z"], kp["K"],
                                    kp["beta2"], kp["delta2"], kp["B2"], inst.inf_a, inst.inf_b, inst.nb_public)
    a, b, cc = inst.solution_abc()
    sol = g16.R1CSSolution(W=inst.wires(), A=a, B=b, C=cc)
    rs = [12345, 67890]
    it = iter(rs)
    proof = g16.ProveSolution(pk, sol, g16.WithDeviceID(0), g16.WithRandomness(lambda q: next(it)))
    e = gf.expected(inst, *rs)
    ok = gf.verify_points(inst, ec.unpack_points(c, 1, proof.Ar)[0], ec.unpack_points(c, 2, proof.Bs)[0],
                          ec.unpack_points(c, 1, proof.Krs)[0], e)
    print("verified:", ok)
    for mode in ("", "1"):
        if mode:
            os.environ["GB200_STEP_PROFILE"] = "1"
        for _ in range(reps):
            t0 = time.perf_counter()
            g16.ProveSolution(pk, sol, g16.WithDeviceID(0))
            print("prove ms%s: %.3f" % (" (step profile, serialised)" if mode else "", 1e3 * (time.perf_counter() - t0)))
    pk.free_gpu_resources()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
