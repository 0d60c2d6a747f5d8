#!/usr/bin/env python3
"""Sharded NTT (gnark_b200/parallel_ntt.py) on N GPUs: correctness against the single-GPU transform and timing.

   torchrun --nproc-per-node N tools/bench_sharded_ntt.py [--curve bls12-381] [--log2n 24] [--steps 10]

Every rank builds the same seeded input, takes its CYCLIC shard, runs forward (-> SLICED evaluations) and the
inverse back; rank 0 also runs the whole transform on its own GPU (b200_ntt) and every rank compares its SLICED
shard against it.  One JSON line on rank 0: ms per forward transform (CUDA events, max over ranks)."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from gnark_b200 import lib, parallel_ntt as pn
    ap = argparse.ArgumentParser()
    ap.add_argument("--curve", default="bls12-381")
    ap.add_argument("--log2n", type=int, default=24)
    ap.add_argument("--steps", type=int, default=10)
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
    curve = {"bn254": lib.BN254, "bls12-381": lib.BLS12_381, "bls12-377": lib.BLS12_377, "bw6-761": lib.BW6_761}[args.curve]
    L = lib.CURVE_SHAPES[curve][0]
    n = 1 << args.log2n
    rs = np.random.RandomState(11)
    X = rs.randint(0, 1 << 62, size=(n, L), dtype=np.int64).astype(np.uint64)
    X[:, L - 1] &= np.uint64((1 << 56) - 1)            # < r for every curve here: valid Montgomery residues
    sd = pn.ShardedDomain(curve, args.log2n, rank, world, dev=local)
    mine = torch.from_numpy(pn.cyclic_shard(X, world, rank).view(np.int64).reshape(-1)).cuda()
    # reference on this GPU: DIF (natural -> bit-reversed) + bit reversal = natural-order evaluations
    full = torch.from_numpy(X.view(np.int64).reshape(-1).copy()).cuda()
    dom = lib.Domain(curve, args.log2n, dev=local)
    dom.ntt_async(full, inverse=False, decimation=lib.DIF)
    lib.vec_bit_reverse(local, curve, full, args.log2n)
    lib.sync(local)
    want = pn.sliced_shard(full.cpu().numpy().view(np.uint64).reshape(n, L), world, rank)
    del full
    got = sd.forward(mine.clone())
    ok = bool(np.array_equal(got.cpu().numpy().view(np.uint64).reshape(-1, L), want))
    back = sd.inverse(got)
    ok &= bool(torch.equal(back, mine))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    bufs = [mine.clone() for _ in range(args.steps + 2)]
    for b in bufs[:2]:
        sd.forward(b)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    with torch.cuda.stream(sd.stream):
        e0.record()
    for b in bufs[2:]:
        sd.forward(b)
    with torch.cuda.stream(sd.stream):
        e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    # single-GPU time for the same size, for the ratio
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    y = torch.from_numpy(X.view(np.int64).reshape(-1).copy()).cuda()
    dom.ntt_async(y); lib.sync(local)
    stream = torch.cuda.Stream(); lib.set_stream(local, stream.cuda_stream)
    with torch.cuda.stream(stream):
        t0.record()
        for _ in range(args.steps):
            dom.ntt_async(y)
        t1.record()
    torch.cuda.synchronize()
    single_ms = t0.elapsed_time(t1) / args.steps
    t = torch.tensor([ms, 0.0 if ok else 1.0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        os.write(real_stdout, (json.dumps({
            "config": f"{args.curve} NTT 2^{args.log2n}, CYCLIC -> SLICED over {world} GPU(s), one all-to-all",
            "n_gpus": world, "ms_per_transform": float(t[0]), "single_gpu_ms": single_ms, "correct": float(t[1]) == 0.0,
            "exchange_bytes_per_rank": (n // world) * L * 8 * (world - 1) // world}) + "\n").encode())
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
