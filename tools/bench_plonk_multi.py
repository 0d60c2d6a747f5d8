#!/usr/bin/env python3
"""PLONK prove on N GPUs (BASELINE configs[3]: BLS12-381, 2^22 gates), one process per GPU:

   torchrun --nproc-per-node N tools/bench_plonk_multi.py [--curve bls12-381] [--log2n 22] [--steps 3]

gnark_b200/plonk.py with shard=(rank, world): every KZG commitment is a point-range-sharded MSM (all_gather of
the partial digests), coset i of the quotient is evaluated on rank i mod N (one all_reduce of the disjoint quarters),
the O(n) stages are replicated.  Satisfied instance + trapdoor SRS (oracle/plonk_fast.py); the proof of the last
timed step is checked by the verifier's equations on rank 0.  The sharded prover's parity is also pinned on the CPU over
gloo (tests/test_dist.py::test_sharded_plonk_gloo).  One JSON line on rank 0 (wall clock around Prove, max over ranks)."""
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
    # the same SATISFIED instance on every rank (oracle/plonk_fast.py), trapdoor SRS built on this rank's GPU
    from oracle import plonk_fast
    inst = plonk_fast.satisfied_instance(c, args.log2n, seed=22)
    cols = {"ql": inst.ql, "qr": inst.qr, "qm": inst.qm, "qo": inst.qo, "qk": inst.qk, "l": inst.l, "r": inst.r, "o": inst.o}
    perm = inst.perm
    srs = plonk_fast.trapdoor_srs_gpu(lib, c, inst, dev=local)
    t0 = time.perf_counter()
    pk = plonk.ProvingKey.from_trace(c.curve_id, args.log2n, cols["ql"], cols["qr"], cols["qm"], cols["qo"], cols["qk"], perm,
                                     srs, dev=local, shard=(rank, world, None))
    load_s = time.perf_counter() - t0
    ic = inst.ch
    ch = plonk.Challenges(gamma=ic.gamma, beta=ic.beta, alpha=ic.alpha, zeta=ic.zeta, v=ic.v, bl=ic.bl, br=ic.br, bo=ic.bo, bz=ic.bz)
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
    verified = None
    if rank == 0:
        # the sharded prover's proof through the verifier's equations (trapdoor identity for the two openings)
        from oracle import ff
        pts = np.stack([proof.LRO[0], proof.LRO[1], proof.LRO[2], proof.Z, proof.H[0], proof.H[1], proof.H[2],
                        proof.LinearizedDigest, proof.BatchedProofH, proof.ZShiftedOpeningH])
        vals = ff.pack_elements(list(proof.BatchedClaimedValues[:6]) + [proof.ZShiftedClaimedValue], c.r, c.fr_limbs)
        verified = bool(plonk_fast.verify(c, inst, pts, vals, with_pairing=False))
    if rank == 0:
        os.write(real_stdout, (json.dumps({
            "config": f"PLONK prove 2^{args.log2n} {args.curve}, {world} GPU(s): sharded KZG commitments + coset-parallel quotient",
            "n_gpus": world, "metric": "plonk_prove_ms", "value": float(np.median(times[1:])), "first_call_ms": times[0],
            "verified": verified, "stage_ms_rank0": stages, "key_load_s": load_s, "data": "synthetic satisfied instance, trapdoor SRS",
            "excludes": "solver, Fiat-Shamir hashing (challenges injected)"}) + "\n").encode())
    # orderly exit: barrier, synchronize, tear the process group down (a watchdog only for a hung teardown)
    import threading
    dog = threading.Timer(120, lambda: os._exit(0))
    dog.daemon = True
    dog.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    if world > 1:
        dist.destroy_process_group()
    dog.cancel()


if __name__ == "__main__":
    main()
