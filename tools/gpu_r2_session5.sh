#!/usr/bin/env bash
# Round 2, GPU session 5 (one B200): tree-shaped set sum, 3 blocks/SM for BN254 G2, L2 fill granularity 64 B, warp-level
# synchronisation in the NTT pass - parity, racecheck, bench, stage times, the chunk / window knobs under the new
# pipeline, ncu captures for profiles/.  Outputs: gpurun_out/s5_*.
set -u
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
L=$OUT/s5_session.log
: > $L
t0=$(date +%s)
lap() { echo "== [$(( $(date +%s) - t0 )) s] $*" | tee -a $L; }

lap "1. parity suite"
timeout 1500 python -m pytest tests -q -m gpu -rfEs -p no:cacheprovider 2>&1 | tail -60 > $OUT/s5_pytest.log
tail -25 $OUT/s5_pytest.log | tee -a $L

lap "2. racecheck / memcheck of the kernels with warp-level synchronisation (small sizes)"
timeout 600 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_ntt.py -q -x -k "bn254 and (all_modes_small or 12 or 13)" -p no:cacheprovider 2>&1 | tail -8 | tee -a $L
timeout 600 compute-sanitizer --tool racecheck --print-limit 5 python tools/run_msm.py bn254 1 14 1 2>&1 | tail -6 | tee -a $L
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python tools/run_msm.py bn254 2 14 1 2>&1 | tail -6 | tee -a $L

lap "3. bench N=1, all legs"
timeout 900 python bench.py --steps 20 --warmup 3 > $OUT/s5_bench_n1.json 2> $OUT/s5_bench_n1.err
tail -c 2500 $OUT/s5_bench_n1.json | tee -a $L
tail -5 $OUT/s5_bench_n1.err | tee -a $L

lap "4. stage times (defaults)"
for cfg in "bn254 1 20" "bn254 2 20" "bls12-381 1 20" "bls12-381 1 22" "bls12-381 2 20" "bw6-761 1 18" "bn254 1 24"; do
  set -- $cfg
  timeout 400 python tools/sweep_msm.py $1 $2 $3 --reps 5 >> $OUT/s5_defaults.jsonl 2>> $OUT/s5_err.log
done
cut -c1-420 $OUT/s5_defaults.jsonl | tee -a $L

lap "5. knobs under the new pipeline: chunk, window"
timeout 400 python tools/sweep_msm.py bn254 1 20 --reps 10 --set GB200_MSM_CHUNK=4,8,16 > $OUT/s5_knobs.jsonl 2>> $OUT/s5_err.log
timeout 400 python tools/sweep_msm.py bn254 1 20 --reps 10 --set GB200_MSM_WINDOW=16,17 >> $OUT/s5_knobs.jsonl 2>> $OUT/s5_err.log
timeout 400 python tools/sweep_msm.py bls12-381 1 22 --reps 5 --set GB200_MSM_WINDOW=16,17 >> $OUT/s5_knobs.jsonl 2>> $OUT/s5_err.log
cut -c1-420 $OUT/s5_knobs.jsonl | tee -a $L
timeout 400 python tools/sweep_ntt.py --curve bn254 --logs 20,22,24 --tiles 8,9 > $OUT/s5_ntt.jsonl 2>> $OUT/s5_err.log
timeout 400 python tools/sweep_ntt.py --curve bls12-381 --logs 22 --tiles 8,9 >> $OUT/s5_ntt.jsonl 2>> $OUT/s5_err.log
timeout 400 python tools/sweep_ntt.py --curve bw6-761 --logs 20 --tiles 8,9 >> $OUT/s5_ntt.jsonl 2>> $OUT/s5_err.log
cat $OUT/s5_ntt.jsonl | tee -a $L

lap "6. ncu"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/s5_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-strong --no-groth16 --no-plonk --no-cpu > $OUT/s5_ncu_bench.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_msm_accumulate -c 1 -f -o $OUT/s5_ncu_accumulate_bn254_g1 \
    python tools/run_msm.py bn254 1 20 1 > $OUT/s5_ncu_g1.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_msm_accumulate -c 1 -f -o $OUT/s5_ncu_accumulate_bn254_g2 \
    python tools/run_msm.py bn254 2 20 1 > $OUT/s5_ncu_g2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_ntt_pass -c 3 -f -o $OUT/s5_ncu_ntt_pass \
    python -c "
import numpy as np, torch
from gnark_b200 import lib
lib.load(); lib.init([0])
d = lib.Domain(lib.BN254, 20)
x = torch.randint(0, 1 << 60, ((1 << 20) * 4,), dtype=torch.int64, device='cuda')
torch.cuda.synchronize()
d.ntt_async(x); lib.sync(0)
" > $OUT/s5_ncu_ntt.log 2>&1
lap "done"
ls -la $OUT | grep s5_ | tee -a $L
