#!/usr/bin/env python3
"""bench.py - BN254 G1 MSM throughput (BASELINE.json metric, config 2) on N B200s.

  python bench.py [--gpus N] [--steps K] [--warmup W]          # this repo's CUDA path
  python bench.py --impl reference [...]                        # CPU arm (oracle port, all host cores)
  torchrun --nproc-per-node N bench.py --gpus N ...             # one rank per GPU

One "step" = one full MSM (decompose -> sort -> bucket accumulate -> reduce) over
2^20 random scalars per GPU against a device-resident base table (PinToGPU
semantics of the reference, backend/accelerated/icicle/groth16/bn254/icicle.go:185-261).
N > 1: the (scalar, base) index range is sharded over ranks (weak scaling: 2^20 per
GPU), each rank produces one partial point, one NCCL all_gather of N Jacobian
points, host-side group adds (the reference sums its chunk results the same way,
icicle.go:383-411).  Prints ONE JSON line on rank 0.
"""

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "bn254_g1_msm_scalar_muls_per_sec"
UNIT = "scalar-muls/s"
LOG_N = 20
SEED = 0x6E61726B00000002  # config 2 seed
GROTH16_LEG_LIMIT_S = 420
WHOLE_RUN_LIMIT_S = 1500
NCU_TRAFFIC_BYTES = 2.324e9  # k_msm_accumulate at 2^20, one ncu --set full capture (profiles/r01_ncu_accumulate_summary.md)


# --------------------------------------------------------------------------------------
# workload (synthetic, seeded).  Uses the oracle ONLY to materialise known-discrete-log
# bases and to check the result before timing (checker role) and for the CPU arm.
# --------------------------------------------------------------------------------------
def make_workload(n, seed, rank=0):
    """bases: 2^16 distinct known-discrete-log points k_i * G (same on every rank), tiled to n;
    scalars: n uniform field elements per rank.  expected = (sum s_i k_i) * G."""
    from oracle import corelib, ec, ff
    from oracle.params import BN254 as C

    def rand_fr(rs, count):
        # uniform on [0, 2^253) (r ~ 2^253.6); the raw limbs are used as Montgomery residues, which are
        # themselves uniform field elements: no conversion needed
        a = rs.randint(0, 1 << 62, size=(count, 4), dtype=np.int64).astype(np.uint64)
        a[:, 3] &= np.uint64((1 << 61) - 1)
        return a
    small = min(n, 1 << 16)
    ks_small = rand_fr(np.random.RandomState(seed & 0x7FFFFFFF), small)
    pts_small = corelib.fixed_base(C, 1, ec.pack_points(C, 1, [C.g1]), ks_small)
    reps = n // small
    pts = np.tile(pts_small, (reps, 1))
    sc_m = rand_fr(np.random.RandomState((seed + rank + 1) & 0x7FFFFFFF), n)
    dot = corelib.fr_dot(C, np.tile(ks_small, (reps, 1)), sc_m)      # sum k_i s_i mod r
    expected = ec.scalar_mul(ff.Fp(C.p), dot, C.g1)
    return C, pts, sc_m, expected


def jac_to_affine(C, jac):
    from oracle import ec, ff
    return ec.from_jac(ff.Fp(C.p), ec.unpack_points(C, 1, jac, ncoords=3)[0])


# --------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        self.idx = gpu_index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
                for k, nme in enumerate(names):
                    if r[3 + k].lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


IMAD_WIDE_PER_ADD = 1360          # 10 products x (64 + 64 + 8) 32x32->64 multiplier operations, BN254
IMAD_WIDE_LANES_PER_CLK_SM = 32   # measured issue rate of IMAD.WIDE (half-rate pipe), profiles/r01_microbench_pipes.txt
N_SMS = 148


def multiplier_roofline(entries, kernel_ms, clocks):
    mhz = (clocks or {}).get("sm_mhz") or (clocks or {}).get("sm_max_mhz") or 1965.0
    peak = N_SMS * IMAD_WIDE_LANES_PER_CLK_SM * mhz * 1e6 / 1e12
    achieved = entries * IMAD_WIDE_PER_ADD / (kernel_ms / 1e3) / 1e12
    return {"bound": "int32 multiplier issue", "achieved": achieved, "peak": peak, "unit": "T IMAD.WIDE/s",
            "frac": achieved / peak, "sm_mhz": mhz,
            "model": "bucket entries x 1360 IMAD.WIDE / accumulate time; peak = 148 SMs x 32 lanes/clk x SM clock under load"}


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# --------------------------------------------------------------------------------------
def cpu_msm(C, pts, sc, c, threads):
    """the CPU arm: signed-digit Pippenger, affine buckets with batched additions for the full windows and
    extended-Jacobian for the short top window, (window x point-range) tasks over all threads - the structure
    of gnark-crypto's MultiExp, restated in oracle/c/oracle.cpp"""
    from oracle import corelib
    return corelib.msm(C, 1, pts, sc, c=c, nthreads=threads, batch_affine=True)


def host_threads():
    """threads the CPU arm may really use: the affinity mask and the cgroup CPU quota, not just the core count"""
    t = os.cpu_count() or 1
    try:
        t = min(t, len(os.sched_getaffinity(0)))
    except Exception:
        pass
    try:
        q, per = open("/sys/fs/cgroup/cpu.max").read().split()
        if q != "max":
            t = min(t, max(1, int(float(q) / float(per) + 0.5)))
    except Exception:
        pass
    return max(1, t)


def cpu_best_config(C, pts, sc, threads):
    """pick window size and thread count the way gnark-crypto picks its window (by problem size / cores): a quick
    sweep.  On SMT hosts the integer-multiplier-bound inner loop often runs faster on one thread per core, so half
    the hardware threads is tried as well; the faster configuration is the one timed."""
    best = (None, threads, 1e9)
    cands = [threads] if threads < 16 else [threads, threads // 2]
    for th in cands:
        for c in (11, 12, 13, 14, 15, 16, 17, 18):
            t0 = time.perf_counter()
            cpu_msm(C, pts, sc, c, th)
            dt = time.perf_counter() - t0
            if dt < best[2]:
                best = (c, th, dt)
    return best[0], best[1]


def cpu_msm_rate(C, pts, sc, sample_n, reps, threads):
    """oracle port (restatement of gnark-crypto's MultiExp algorithm, NOT gnark-crypto) on host cores."""
    p, s = pts[:sample_n], sc[:sample_n]
    c, th = cpu_best_config(C, p, s, threads)
    t0 = time.perf_counter()
    for _ in range(reps):
        cpu_msm(C, p, s, c, th)
    dt = (time.perf_counter() - t0) / reps
    return sample_n / dt, dt, th, c


def run_reference(args):
    """--impl reference: the reference's CPU implementation of this path.  gnark (Go) cannot be
    built here and its arithmetic lives in the absent gnark-crypto module, so this arm times the
    C++ restatement of the same algorithm (oracle/c/oracle.cpp) on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = 1 << LOG_N
    threads = host_threads()
    sample_n = n if threads >= 16 else n >> 2
    C, pts, sc, expected = make_workload(sample_n, SEED)
    cw, threads = cpu_best_config(C, pts, sc, threads)
    assert jac_to_affine(C, cpu_msm(C, pts, sc, cw, threads)) == expected
    for _ in range(args.warmup):
        cpu_msm(C, pts, sc, cw, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_msm(C, pts, sc, cw, threads)
    dt = time.perf_counter() - t0
    value = sample_n * args.steps / dt
    sample = f"{args.steps} x BN254 G1 MSM of 2^{int(np.log2(sample_n))} points per step on {threads} host threads"
    emit({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32x8 (254-bit Montgomery)",
        "data": "synthetic",
        "config": {"workload": f"BN254 G1 MSM 2^{LOG_N} random scalars/bases (BASELINE configs[1]); CPU arm sample 2^{int(np.log2(sample_n))}"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "algorithm": f"Pippenger c={cw}, batch-affine buckets (restatement of gnark-crypto MultiExp)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })


# --------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    from gnark_b200 import lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun --nproc-per-node N for --gpus N")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib.load()
    lib.init([local])
    stream = torch.cuda.Stream()               # a real (non-null) stream: kernels, events and NCCL all on it
    torch.cuda.set_stream(stream)
    lib.set_stream(local, stream.cuda_stream)

    n = 1 << LOG_N
    C, pts, sc, expected = make_workload(n, SEED, rank)
    table = lib.Table(lib.BN254, 1, pts, dev=local, precomp=True)
    info = table.info()
    d_sc = torch.from_numpy(sc.view(np.int64)).cuda()
    d_out = torch.zeros(12, dtype=torch.int64, device="cuda")
    h_sc = torch.from_numpy(sc.view(np.int64)).pin_memory()
    h_out = np.zeros(12, dtype=np.uint64)

    # correctness on the timed input before timing (known-dlog oracle)
    table.msm_async(d_sc, d_out, n=n)
    torch.cuda.synchronize()
    got = jac_to_affine(C, d_out.cpu().numpy().view(np.uint64))
    assert got == expected, "GPU MSM does not match the known-discrete-log oracle"

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def combine(d_partial):
        """N partial points -> one (all_gather over NCCL, host adds); returns on every rank."""
        if world == 1:
            return d_partial
        parts = [torch.empty_like(d_partial) for _ in range(world)]
        dist.all_gather(parts, d_partial)
        acc = parts[0].cpu().numpy().view(np.uint64).copy()
        for p in parts[1:]:
            lib.point_add_jac(lib.BN254, 1, acc, p.cpu().numpy().view(np.uint64))
        return acc

    # ---- device-resident throughput (value) ------------------------------------------
    # K independent MSMs issued back to back; the latency-bound reduction tail of MSM i overlaps
    # the sort/accumulate of MSM i+1 (b200_msm_pipelined), results joined before the end event.
    K = args.steps
    d_outs = torch.zeros((K, 12), dtype=torch.int64, device="cuda")

    def run_steps(k):
        for i in range(k):
            table.msm_pipelined(d_sc, d_outs[i], n=n)
        table.join()
        if world == 1:
            return d_outs[:k]
        parts = [torch.empty_like(d_outs) for _ in range(world)]
        dist.all_gather(parts, d_outs)          # one exchange of world x K x 96 B over NVLink
        return parts

    def fold(parts, i):
        acc = parts[0][i].cpu().numpy().view(np.uint64).copy()
        for p in parts[1:]:
            lib.point_add_jac(lib.BN254, 1, acc, p[i].cpu().numpy().view(np.uint64))
        return acc

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()       # samples cover warm-up + timed region + e2e region (same kernels throughout)
    run_steps(max(args.warmup, 3))
    t_load = time.perf_counter()
    while time.perf_counter() - t_load < 0.6:   # >= 0.6 s under load so nvidia-smi (100 ms period) sees it
        run_steps(4)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    res = run_steps(K)
    if world > 1:
        totals = [fold(res, i) for i in range(K)]   # host-side group adds inside the timed region
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    if world == 1:
        assert jac_to_affine(C, res[K - 1].cpu().numpy().view(np.uint64)) == expected

    # ---- end to end through the C ABI with host buffers (e2e) --------------------------
    for _ in range(2):
        table.msm(h_sc, n=n)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        part = table.msm(h_sc, n=n)      # H2D scalars (pinned) + MSM + D2H result, synchronous
        if world > 1:
            combine(torch.from_numpy(part.view(np.int64)).cuda())
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    # optional (not yet hardware-validated, off by default): the same K MSMs submitted back to back through
    # b200_msm_submit - upload of step i+1, compute of step i and tail/download of step i-1 overlap
    e2e_submit = None
    if args.e2e_submit and world == 1:
        try:
            h_outs = torch.zeros((args.steps, 12), dtype=torch.int64).pin_memory()
            for _ in range(2):
                table.msm_submit(h_sc, h_outs[0], n=n)
            lib.sync(local)
            t0 = time.perf_counter()
            for i in range(args.steps):
                table.msm_submit(h_sc, h_outs[i], n=n)
            lib.sync(local)
            dt = time.perf_counter() - t0
            ok = jac_to_affine(C, h_outs[args.steps - 1].numpy().view(np.uint64)) == expected
            e2e_submit = {"value": n * args.steps / dt, "unit": UNIT, "ms_per_step": 1e3 * dt / args.steps, "correct": bool(ok),
                          "note": "b200_msm_submit: K MSMs from pinned host scalars, H2D / compute / tail+D2H overlapped"}
        except Exception as e:
            e2e_submit = {"error": repr(e)}
    clocks = sampler.stop() if rank == 0 else None

    # ---- stage profile of the dominant kernel (accumulate) ------------------------------
    prof = []
    for _ in range(5):
        prof.append(table.msm_profile(d_sc, d_out, n=n))
    stage_ms = {k: float(np.median([p[k] for p in prof])) for k in prof[0]}

    if world > 1:
        t = torch.tensor([ms, e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_s = float(t[0]), float(t[1])
    table.free()
    if rank != 0:
        if not args.no_groth16:
            arm_watchdog(GROTH16_LEG_LIMIT_S)
            try:
                groth16_leg(local, pts, n, rank, world)
            except Exception:
                pass
        finish(world)

    total = n * world
    value = total * args.steps / (ms / 1e3)
    e2e_value = total * args.steps / e2e_s
    peak, peak_src = measured_peak_gbs()
    W = info["n_windows"]
    alg_bytes = n * W * (64 + 4)           # W x (sizeof(affine) + 4 B index) per scalar-mul
    acc_ms = stage_ms["accumulate"]
    achieved = alg_bytes / (acc_ms / 1e3) / 1e9
    threads = host_threads()
    sample_n = n if threads >= 32 else 1 << 18   # many-core hosts need the full problem to scale
    if world == 1:
        cpu_rate, cpu_dt, cpu_th, cpu_c = cpu_msm_rate(C, pts, sc, sample_n, 3, threads)
        cpu_baseline = {"value": cpu_rate, "unit": UNIT, "cores": cpu_th, "kind": "port",
                        "sample": f"3 x BN254 G1 MSM of 2^{int(np.log2(sample_n))} points of the same workload "
                                  f"({cpu_dt:.2f} s each; window c={cpu_c} and thread count {cpu_th} of {threads} "
                                  f"usable picked by a sweep)"}
    else:
        cpu_baseline = {"value": None, "unit": UNIT, "cores": threads, "kind": "port",
                        "sample": "timed on rank 0 at N=1 only (see the N=1 line)"}
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32x8 (254-bit Montgomery integers)", "data": "synthetic",
        "config": {
            "workload": f"BN254 G1 MSM 2^{LOG_N} random scalars/bases per GPU (BASELINE configs[1]), bases device-resident",
            "points_per_gpu": n, "window_bits": info["window_bits"], "windows": W, "precomputed_table": bool(info["precomp"]),
            "table_bytes": info["device_bytes"], "parallelism": f"point-range shard x{world}" if world > 1 else "single GPU",
            "l2": "no flush: per-step working set (table %.2f GiB + scalars 32 MiB + 256 MiB sort buffers) exceeds the 126 MB L2" % (info["device_bytes"] / 2**30),
        },
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": n * 32 * world, "d2h_bytes_per_step": 96 * world,
                "ms_per_step": 1e3 * e2e_s / args.steps,
                "note": "b200_msm_g1 with pinned host scalars; bases resident (PinToGPU)"},
        # own kernels per MSM: decompose, bucket_offsets, task_counts, accumulate, combine, combine_heavy,
        # reduce_chunks, set_sum, finish (the radix sort and the scan are CUB launches on top of these)
        "gpu_launches": 9 * args.steps,
        "roofline": {"bound": "hbm", "kernel": "k_msm_accumulate", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": NCU_TRAFFIC_BYTES, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": acc_ms,
                     "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum, profiles/r01_ncu_accumulate_summary.md",
                     "binding_unit": {"unit": "sm__pipe_fmaheavy (IMAD.WIDE issue)", "pct_of_peak": 86.5,
                                      "source": "profiles/r01_ncu_accumulate_summary.md"},
                     # the same kernel against the ceiling that binds it, computed live: one XYZZ mixed addition per
                     # bucket entry = 10 Montgomery products x 136 IMAD.WIDE (static SASS count, profiles/r01_sass_stats.md);
                     # the pipe issues 32 IMAD.WIDE lanes / clk / SM (tools/microbench.cu, profiles/r01_microbench_pipes.txt)
                     "multiplier": multiplier_roofline(n * W, acc_ms, clocks),
                     "note": "integer-multiplier bound, not HBM bound: ~1360 IMAD.WIDE per gathered 68 B (DESIGN.md)"},
        "stage_ms": stage_ms,
        "e2e_submit": e2e_submit,
        "cpu_baseline": cpu_baseline,
        "clocks": clocks,
    }
    emit(out) if args.no_groth16 else None
    if args.no_groth16:
        finish(world)
    def headline_only():
        out["groth16"] = {"error": "secondary leg did not finish within %d s" % GROTH16_LEG_LIMIT_S}
        emit(out)
    dog = arm_watchdog(GROTH16_LEG_LIMIT_S, headline_only)
    try:
        out["groth16"] = groth16_leg(local, pts, n, 0, world)
    except Exception as e:  # the headline metric must still be printed
        out["groth16"] = {"error": repr(e)}
    dog.cancel()
    emit(out)
    arm_watchdog(60)          # the line is out; never hang in the closing barrier
    finish(world)


def arm_watchdog(seconds, last_words=None, code=0):
    """a stuck secondary leg (a rank lost inside a collective) must not cost the headline line: after
    `seconds` run last_words() (rank 0: print what is already measured) and leave"""
    def fire():
        try:
            if last_words is not None:
                last_words()
        finally:
            sys.stderr.write("bench.py: watchdog fired after %d s\n" % seconds)
            sys.stderr.flush()
            os._exit(code)
    t = threading.Timer(seconds, fire)
    t.daemon = True
    t.start()
    return t


def finish(world):
    """orderly exit without NCCL teardown (destroy_process_group can stall at exit on some multi-rank boxes)"""
    import torch
    import torch.distributed as dist
    try:
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    finally:
        sys.stdout.flush(); sys.stderr.flush()
        os._exit(0)


def groth16_leg(dev, g1_pts, n, rank=0, world=1):
    """Secondary BASELINE metric: Groth16 prove ms at 2^20 R1CS (configs[2]), BN254, 1 GPU.
    From "A,B,C,W on the host" to "3 proof points on the host" (solver excluded, as in
    SURVEY.md §8d config 3): 7 NTT(2^20) + 4 G1 MSM + 1 G2 MSM + host-side assembly.
    Synthetic key: G1 tables reuse the benchmark's 2^20 known-dlog points, G2.B tiles 2^16
    oracle-generated points; synthetic (unsatisfied) solution vectors - timing only, the
    pipeline's parity is pinned by tests/test_gpu_groth16.py at small sizes."""
    from gnark_b200 import groth16 as g16
    from oracle import corelib, ec, ff
    from oracle.params import BN254 as C
    import torch
    import torch.distributed as dist
    rs = np.random.RandomState(5)

    def rand_fr(count):
        a = rs.randint(0, 1 << 62, size=(count, 4), dtype=np.int64).astype(np.uint64)
        a[:, 3] &= np.uint64((1 << 61) - 1)
        return a
    # bases are identical on every rank (make_workload), so every rank holds the SAME synthetic key to shard
    nb_wires, nb_public = n + 2, 2
    g2_small = corelib.fixed_base(C, 2, ec.pack_points(C, 2, [C.g2]), rand_fr(1 << 16))
    g2_b = np.tile(g2_small, (n // (1 << 16) + 1, 1))[:nb_wires].copy()
    g1 = np.concatenate([g1_pts, g1_pts[:2]])
    pk = g16.ProvingKey.from_arrays(
        g16.BN254, n, g1[0], g1[1], g1[2], g1[:nb_wires], g1[:nb_wires], g1[:n - 1], g1[:nb_wires - nb_public],
        g2_small[0], g2_small[1], g2_b, np.zeros(nb_wires, dtype=np.uint8), np.zeros(nb_wires, dtype=np.uint8),
        nb_public)
    opts = [g16.WithDeviceID(dev), g16.WithSharding(rank, world)]
    t0 = time.perf_counter()
    pk.setup_device_pointers(g16.NewConfig(*opts))
    setup_s = time.perf_counter() - t0
    sol_pageable = g16.R1CSSolution(W=rand_fr(nb_wires), A=rand_fr(n - 1), B=rand_fr(n - 1), C=rand_fr(n - 1))
    # the solver's output vectors live in C-owned pinned buffers (b200_host_alloc)
    keep = [torch.from_numpy(v.view(np.int64)).pin_memory() for v in (sol_pageable.W, sol_pageable.A, sol_pageable.B, sol_pageable.C)]
    sol = g16.R1CSSolution(*[k.numpy().view(np.uint64) for k in keep])
    times = []
    for i in range(10):
        cur = sol if i < 6 else sol_pageable
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        g16.ProveSolution(pk, cur, *opts)
        dt = 1e3 * (time.perf_counter() - t0)
        if world > 1:       # max over ranks
            t = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t[0])
        times.append(dt)
    pk.free_gpu_resources()
    return {"metric": "groth16_prove_ms", "n_constraints": n - 1, "curve": "bn254", "n_gpus": world,
            "parallelism": "every MSM table point-range sharded x%d, computeH replicated, 1 all_gather of 5 points" % world,
            "prove_ms_median": float(np.median(times[1:6])), "prove_ms_min": float(min(times[1:6])),
            "prove_ms_pageable_median": float(np.median(times[6:])),
            "first_call_ms": times[0], "key_load_s": setup_s,
            "includes": "H2D of W,A,B,C (4 x 32 MiB, pinned host buffers; pageable variant reported beside it), "
                        "computeH (7 NTT), 5 MSM, D2H, host assembly",
            "excludes": "R1CS solver (CPU, out of scope)"}


_REAL_STDOUT = None


def emit(obj):
    """the ONE JSON line on the real stdout (fd 1 is pointed at stderr while the bench runs, so that
    library chatter such as NCCL's version banner cannot pollute it)"""
    line = (json.dumps(obj) + "\n").encode()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, line)


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-groth16", action="store_true", help="skip the secondary Groth16 2^20 prove-time leg")
    ap.add_argument("--e2e-submit", action="store_true", help="also time the asynchronous host-buffer path (b200_msm_submit)")
    args = ap.parse_args()
    arm_watchdog(WHOLE_RUN_LIMIT_S, code=1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
