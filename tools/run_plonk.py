#!/usr/bin/env python3
"""Small driver for profiling: PLONK proofs of a satisfied instance on cuda:0 (used under ncu for the launch list).
   python tools/run_plonk.py <curve> <log2 n> [reps]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gnark_b200 import lib  # noqa: E402
from oracle import plonk_fast  # noqa: E402  (fixture + check only)
from oracle.params import CURVES  # noqa: E402


def main():
    c = CURVES[sys.argv[1]]
    logn = int(sys.argv[2])
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    lib.load(); lib.init([0])
    inst = plonk_fast.satisfied_instance(c, logn, seed=logn)
    srs = plonk_fast.trapdoor_srs_gpu(lib, c, inst)
    key = lib.PlonkKey(c.curve_id, logn, inst.ql, inst.qr, inst.qm, inst.qo, inst.qk, inst.perm, srs)
    ch = inst.challenges_packed()
    for i in range(reps):
        t0 = time.perf_counter()
        pts, vals = key.prove(inst.l, inst.r, inst.o, *ch)
        print(f"prove {i}: {1e3 * (time.perf_counter() - t0):.1f} ms, stages {key.last_stage_ms()}", flush=True)
    print("verified:", plonk_fast.verify(c, inst, pts, vals, with_pairing=False))
    key.free()


if __name__ == "__main__":
    main()
