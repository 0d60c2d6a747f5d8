#!/usr/bin/env python3
"""NTT timing on cuda:0 for the tile sizes of GB200_NTT_TILE_LOG (8 = default: 128-thread blocks, eight per SM;
11: one 1024-thread block per SM).  One JSON line per (curve, log2n, tile); every tile size must reproduce the first
one's output bit for bit (the default is pinned against the oracle by tests/test_gpu_ntt.py).

   python tools/sweep_ntt.py [--curve bn254] [--logs 20,22,24] [--tiles 8,9,10,11] [--reps 20]"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from gnark_b200 import lib
    ap = argparse.ArgumentParser()
    ap.add_argument("--curve", default="bn254")
    ap.add_argument("--logs", default="20,22,24")
    ap.add_argument("--tiles", default="8,9,10,11")
    ap.add_argument("--reps", type=int, default=20)
    args = ap.parse_args()
    curve = {"bn254": lib.BN254, "bls12-381": lib.BLS12_381, "bls12-377": lib.BLS12_377, "bw6-761": lib.BW6_761}[args.curve]
    L = lib.CURVE_SHAPES[curve][0]
    lib.load(); lib.init([0])
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    lib.set_stream(0, stream.cuda_stream)
    rs = np.random.RandomState(3)
    for logn in [int(x) for x in args.logs.split(",")]:
        n = 1 << logn
        X = rs.randint(0, 1 << 62, size=(n, L), dtype=np.int64).astype(np.uint64)
        X[:, L - 1] &= np.uint64((1 << 56) - 1)
        x0 = torch.from_numpy(X.view(np.int64).reshape(-1)).cuda()
        ref = None
        for tile in [int(t) for t in args.tiles.split(",")]:
            os.environ["GB200_NTT_TILE_LOG"] = str(tile)
            d = lib.Domain(curve, logn)
            outs = []
            for inv, dec, cos in ((False, lib.DIF, False), (True, lib.DIT, True)):
                y = x0.clone()
                d.ntt_async(y, inverse=inv, decimation=dec, on_coset=cos)
                outs.append(y)
            lib.sync(0)
            if ref is None:
                ref = outs
            same = all(torch.equal(a, b) for a, b in zip(outs, ref))
            y = x0.clone()
            d.ntt_async(y); lib.sync(0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.reps):
                d.ntt_async(y)
            e1.record()
            lib.sync(0)
            ms = e0.elapsed_time(e1) / args.reps
            print(json.dumps({"curve": args.curve, "log2n": logn, "tile_log": tile, "ms": ms, "matches_first": bool(same),
                              "GBps_model": 2 * n * L * 8 * (2 if logn <= 2 * tile else 3) / ms / 1e6}), flush=True)
            d.free()
        os.environ.pop("GB200_NTT_TILE_LOG", None)


if __name__ == "__main__":
    main()
