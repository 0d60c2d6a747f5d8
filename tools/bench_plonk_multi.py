#!/usr/bin/env python3
"""PLONK prove on N GPUs (BASELINE configs[3]: BLS12-381, 2^22 gates), one process per GPU:

   torchrun --nproc-per-node N tools/bench_plonk_multi.py [--curve bls12-381] [--log2n 22] [--steps 3]

gnark_b200/plonk.py with shard=(rank, world): every KZG commitment is a point-range-sharded MSM (all_gather of
the partial digests), coset i of the quotient is evaluated on rank i mod N (one all_reduce of the disjoint quarters),
the O(n) stages are replicated.  Synthetic (unsatisfied) instance: same work as a real proof, timing only - the
parity of the sharded prover is pinned on the CPU over gloo (tests/test_dist.py::test_sharded_plonk_gloo) and of the
single-GPU prover on hardware at small sizes.  One JSON line on rank 0 (wall clock around Prove, max over ranks)."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from gnark_b200 import lib, plonk
    from oracle import corelib, ec           # fixture generation only
    from oracle.params import CURVES
    ap = argparse.ArgumentParser()
    ap.add_argument("--curve", default="bls12-381")
    ap.add_argument("--log2n", type=int, default=22)
    ap.add_argument("--steps", type=int, default=3)
    args = ap.parse_args()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib.load(); lib.init([local])
    c = CURVES[args.curve]
    L = c.fr_limbs
    n = 1 << args.log2n
    rs = np.random.RandomState(21)                       # the same instance on every rank

    def rand_fr(count):
        a = rs.randint(0, 1 << 62, size=(count, L), dtype=np.int64).astype(np.uint64)
        a[:, L - 1] &= np.uint64((1 << (c.r.bit_length() - 64 * (L - 1) - 1)) - 1)
        return a
    cols = {k: rand_fr(n) for k in ("ql", "qr", "qm", "qo", "qk", "l", "r", "o")}
    perm = rs.permutation(3 * n).astype(np.int64)
    small = 1 << 14
    srs = np.tile(corelib.fixed_base(c, 1, ec.pack_points(c, 1, [c.g1]), rand_fr(small)), ((n + 3) // small + 1, 1))[:n + 3].copy()
    t0 = time.perf_counter()
    pk = plonk.ProvingKey.from_trace(c.curve_id, args.log2n, cols["ql"], cols["qr"], cols["qm"], cols["qo"], cols["qk"], perm,
                                     srs, dev=local, shard=(rank, world, None))
    load_s = time.perf_counter() - t0
    ri = lambda: int(rs.randint(1, 1 << 62))
    ch = plonk.Challenges(gamma=ri(), beta=ri(), alpha=ri(), zeta=ri(), v=ri(), bl=[ri(), ri()], br=[ri(), ri()],
                          bo=[ri(), ri()], bz=[ri(), ri(), ri()])
    times, stages = [], None
    for _ in range(args.steps + 1):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        proof = plonk.Prove(pk, cols["l"], cols["r"], cols["o"], ch)
        dt = 1e3 * (time.perf_counter() - t0)
        if world > 1:
            tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt = float(tt[0])
        times.append(dt)
        stages = proof.timings_ms
    if rank == 0:
        os.write(real_stdout, (json.dumps({
            "config": f"PLONK prove 2^{args.log2n} {args.curve}, {world} GPU(s): sharded KZG commitments + coset-parallel quotient",
            "n_gpus": world, "metric": "plonk_prove_ms", "value": float(np.median(times[1:])), "first_call_ms": times[0],
            "stage_ms_rank0": stages, "key_load_s": load_s, "data": "synthetic (unsatisfied instance)",
            "excludes": "solver, Fiat-Shamir hashing (challenges injected)"}) + "\n").encode())
    try:
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    finally:
        os._exit(0)


if __name__ == "__main__":
    main()
