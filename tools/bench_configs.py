#!/usr/bin/env python3
"""The other BASELINE.json configurations, measured the same way as bench.py (CUDA events on the
launching stream, max over ranks), one JSON line per config on rank 0:

  configs[1'] BN254 G1 MSM, 2^24 points total, point-range sharded over the ranks
              (north_star target: >= 1e8 scalar-muls/s at 8 x B200)
  configs[4]  BW6-761 G1 MSM, 2^24 points total, sharded (8 x 2^21)
  configs[3]  PLONK 2^22 BLS12-381 building blocks on one GPU: NTT 2^22, NTT 2^24, KZG-commit MSM 2^22,
              fused constraint kernel - primitive timings x the per-proof counts of SURVEY.md §8d

  torchrun --nproc-per-node N tools/bench_configs.py [--total-log 24] [--steps 5]
Bases: 2^14 distinct known-dlog points per rank (oracle fixed-base), tiled; correctness of every
kernel at these sizes is pinned by tests/ (known-dlog at 2^20); here the first tile is checked.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def rand_fr(rs, c, count):
    L = c.fr_limbs
    a = rs.randint(0, 1 << 62, size=(count, L), dtype=np.int64).astype(np.uint64)
    a[:, L - 1] &= np.uint64((1 << (c.r.bit_length() - 64 * (L - 1) - 1)) - 1)
    return a


def main():
    import torch
    import torch.distributed as dist
    from gnark_b200 import lib
    from oracle import corelib, derive, ec, ff
    from oracle.params import CURVES

    ap = argparse.ArgumentParser()
    ap.add_argument("--total-log", type=int, default=24)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--skip-plonk", action="store_true")
    ap.add_argument("--plonk-log", type=int, default=22)
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
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    lib.set_stream(local, stream.cuda_stream)

    def emit(obj):
        if rank == 0:
            os.write(real_stdout, (json.dumps(obj) + "\n").encode())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def sharded_msm_bench(cname, total_log):
        c = CURVES[cname]
        rs = np.random.RandomState(100 + rank)
        n_total = 1 << total_log
        n = n_total // world
        small = 1 << 14
        ks = rand_fr(rs, c, small)
        base = derive.subgroup_point(c, 1)
        pts_small = corelib.fixed_base(c, 1, ec.pack_points(c, 1, [base]), ks)
        pts = np.tile(pts_small, (n // small, 1))
        sc = rand_fr(rs, c, n)
        t0 = time.perf_counter()
        table = lib.Table(c.curve_id, 1, pts, dev=local, precomp=True)
        load_s = time.perf_counter() - t0
        info = table.info()
        # check on the first tile (known dlog)
        got = ec.from_jac(ff.Fp(c.p), ec.unpack_points(c, 1, table.msm(sc[:small].copy(), n=small), ncoords=3)[0])
        want = ec.scalar_mul(ff.Fp(c.p), corelib.fr_dot(c, ks, sc[:small].copy()), base)
        assert got == want, f"{cname}: first-tile MSM mismatch"
        d_sc = torch.from_numpy(sc.view(np.int64)).cuda()
        K = args.steps
        limbs = 3 * c.fp_limbs
        d_outs = torch.zeros((K, limbs), dtype=torch.int64, device="cuda")

        def run(k):
            for i in range(k):
                table.msm_pipelined(d_sc, d_outs[i], n=n)
            table.join()
            if world > 1:
                parts = [torch.empty_like(d_outs) for _ in range(world)]
                dist.all_gather(parts, d_outs)
                for i in range(k):
                    acc = parts[0][i].cpu().numpy().view(np.uint64).copy()
                    for p in parts[1:]:
                        lib.point_add_jac(c.curve_id, 1, acc, p[i].cpu().numpy().view(np.uint64))
        run(2)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(K); e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        prof = table.msm_profile(d_sc, d_outs[0], n=n)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t[0])
        table.free()
        emit({"config": f"{cname} G1 MSM 2^{total_log} points total, point-range shard x{world}", "n_gpus": world,
              "points_per_gpu": n, "metric": "scalar-muls/s", "value": n_total * K / (ms / 1e3), "ms_per_msm": ms / K,
              "steps": K, "window_bits": info["window_bits"], "windows": info["n_windows"],
              "table_bytes_per_gpu": info["device_bytes"], "table_load_s": load_s, "stage_ms_rank0": prof,
              "scaling": "strong (fixed total)", "data": "synthetic (2^14 distinct known-dlog bases per rank, tiled)"})

    sharded_msm_bench("bn254", args.total_log)
    sharded_msm_bench("bw6-761", args.total_log)

    if rank == 0 and not args.skip_plonk:
        c = CURVES["bls12-381"]
        rs = np.random.RandomState(7)
        L = c.fr_limbs
        out = {"config": "PLONK 2^22 BLS12-381 building blocks, 1 GPU", "n_gpus": 1}

        def time_it(fn, reps=5):
            fn(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                fn()
            e1.record(); torch.cuda.synchronize()
            return e0.elapsed_time(e1) / reps
        for logn in (22, 24):
            d = lib.Domain(c.curve_id, logn, dev=local)
            x = torch.from_numpy(rand_fr(rs, c, 1 << logn).view(np.int64)).cuda()
            out[f"ntt_2^{logn}_ms"] = time_it(lambda: d.ntt_async(x, inverse=False, decimation=lib.DIF))
            out[f"ntt_2^{logn}_coset_inverse_ms"] = time_it(lambda: d.ntt_async(x, inverse=True, decimation=lib.DIT, on_coset=True))
            if logn == 22:
                # fused constraint kernel on one coset
                polys = {k: torch.from_numpy(rand_fr(rs, c, 1 << logn).view(np.int64)).cuda()
                         for k in ("l", "r", "o", "z", "s1", "s2", "s3", "ql", "qr", "qm", "qo", "qk")}
                res = torch.zeros((4 << logn) * L, dtype=torch.int64, device="cuda")
                one = ff.pack_elements([5], c.r, L)
                g = ff.pack_elements([c.mult_gen], c.r, L)
                w4 = ff.pack_elements([pow(c.root_of_unity, 1 << (c.two_adicity - logn - 2), c.r)], c.r, L)
                bl = {"l": rand_fr(rs, c, 2), "r": rand_fr(rs, c, 2), "o": rand_fr(rs, c, 2), "z": rand_fr(rs, c, 3)}
                out["quotient_constraints_one_coset_ms"] = time_it(
                    lambda: lib.plonk_constraints_coset(d, g, w4, polys, one, one, one, bl, 1, 4, res))
                del polys, res
            d.free()
            del x
        # KZG commit = MSM 2^22 on the SRS
        small = 1 << 14
        ks = rand_fr(rs, c, small)
        pts = np.tile(corelib.fixed_base(c, 1, ec.pack_points(c, 1, [c.g1]), ks), ((1 << 22) // small, 1))
        table = lib.Table(c.curve_id, 1, pts, dev=local, precomp=True)
        d_sc = torch.from_numpy(rand_fr(rs, c, 1 << 22).view(np.int64)).cuda()
        d_o = torch.zeros(3 * c.fp_limbs, dtype=torch.int64, device="cuda")

        def msm4():
            for _ in range(4):
                table.msm_pipelined(d_sc, d_o, n=1 << 22)
            table.join()
        out["kzg_commit_msm_2^22_ms"] = time_it(msm4, reps=2) / 4
        table.free()
        # ~108 NTT(2^22) + 1 NTT(2^24) + 10 MSM(~2^22) + 4 coset constraint passes
        out["estimated_prover_arithmetic_ms"] = (108 * out["ntt_2^22_ms"] + out["ntt_2^24_coset_inverse_ms"]
                                                 + 10 * out["kzg_commit_msm_2^22_ms"]
                                                 + 4 * out["quotient_constraints_one_coset_ms"])
        out["note"] = ("sum of primitive timings x per-proof counts (SURVEY.md §8d config 4); the O(n) scans, "
                       "Fiat-Shamir and host orchestration of a full PLONK prover are not included")
        emit(out)
        # the real thing: one b200_plonk_prove call per proof (plonk_host.cu), synthetic unsatisfied instance -
        # same work as a real proof (the prover does not test satisfiability), timing only; parity of the
        # pipeline is pinned at small sizes by the tests
        try:
            logn = args.plonk_log
            n = 1 << logn
            rs2 = np.random.RandomState(9)
            cols = {k: rand_fr(rs2, c, n) for k in ("ql", "qr", "qm", "qo", "qk", "l", "r", "o")}
            perm = rs2.permutation(3 * n).astype(np.int64)
            small = 1 << 14
            srs = np.tile(corelib.fixed_base(c, 1, ec.pack_points(c, 1, [c.g1]), rand_fr(rs2, c, small)),
                          ((n + 3) // small + 1, 1))[:n + 3].copy()
            t0 = time.perf_counter()
            key = lib.PlonkKey(c.curve_id, logn, cols["ql"], cols["qr"], cols["qm"], cols["qo"], cols["qk"], perm, srs,
                               dev=local)
            load_s = time.perf_counter() - t0
            one = lambda k: rand_fr(rs2, c, k)
            chal = [one(1) for _ in range(5)] + [one(2), one(2), one(2), one(3)]
            times = []
            for _ in range(4):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                key.prove(cols["l"], cols["r"], cols["o"], *chal)
                times.append(1e3 * (time.perf_counter() - t0))
            key.free()
            emit({"config": f"PLONK prove 2^{logn} BLS12-381, 1 GPU, one b200_plonk_prove call per proof",
                  "n_gpus": 1, "metric": "plonk_prove_ms", "value": float(np.median(times[1:])), "first_call_ms": times[0],
                  "key_load_s": load_s, "includes": "H2D of L,R,O, 60 NTTs of 2^n, 4 fused constraint passes, iNTT 4n, "
                  "10 KZG commitments (MSM), grand product, evaluations, opening quotients, D2H of 10 points + 7 values",
                  "excludes": "solver, Fiat-Shamir hashing (challenges injected)", "data": "synthetic (unsatisfied instance)"})
        except Exception as e:
            emit({"config": "PLONK prove", "error": repr(e)})
    # no NCCL teardown: at 8 ranks destroy_process_group stalled after both results had been emitted
    # (round-1 run: 600 s lost to the timeout); barrier, flush and leave
    try:
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    finally:
        sys.stdout.flush(); sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
