#!/usr/bin/env python3
"""Small driver for profiling: one MSM of a given curve/group/size on cuda:0 (used under ncu).
   python tools/run_msm.py <curve> <group> <log2 n> [reps]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gnark_b200 import lib  # noqa: E402
from oracle import corelib, ec, ff  # noqa: E402  (input generation + result check only)
from oracle.params import CURVES  # noqa: E402
from oracle import derive  # noqa: E402


def main():
    c = CURVES[sys.argv[1]]
    group = int(sys.argv[2])
    logn = int(sys.argv[3])
    reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
    n = 1 << logn
    rs = np.random.RandomState(1)
    L = c.fr_limbs

    def rand_fr(count):
        a = rs.randint(0, 1 << 62, size=(count, L), dtype=np.int64).astype(np.uint64)
        a[:, L - 1] &= np.uint64((1 << (c.r.bit_length() - 64 * (L - 1) - 2)) - 1)
        return a
    base = derive.subgroup_point(c, group)
    small = min(n, 1 << 14)
    ks = rand_fr(small)
    pts = corelib.fixed_base(c, group, ec.pack_points(c, group, [base]), ks)
    pts = np.tile(pts, (n // small, 1))
    sc = rand_fr(n)
    lib.load(); lib.init([0])
    import torch
    t = lib.Table(c.curve_id, group, pts, precomp=True)
    d_sc = torch.from_numpy(sc.view(np.int64)).cuda()
    d_out = torch.zeros(3 * t.coord_limbs, dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()      # the library runs on its own stream: torch's uploads must have landed
    for _ in range(2):
        t.msm_async(d_sc, d_out, n=n)
    lib.sync(0)
    t0 = time.perf_counter()
    for _ in range(reps):
        t.msm_async(d_sc, d_out, n=n)
    lib.sync(0)
    dt = (time.perf_counter() - t0) / reps
    prof = t.msm_profile(d_sc, d_out, n=n)
    print(f"{c.name} G{group} n=2^{logn}: {1e3 * dt:.3f} ms per MSM, {n / dt / 1e6:.1f} M scalar-muls/s; stages {prof}; table {t.info()}")


if __name__ == "__main__":
    main()
