#!/usr/bin/env bash
# Round 2, GPU session 2 (one B200): the pruned library with the new defaults (operands by value, lazy Fp2, 2^9 NTT
# tiles): parity, bench, the remaining tuning questions (task length under pipelining, NTT tiles below 2^9, persistent
# accumulate on top of by-value operands), ncu captures for profiles/.  Outputs: gpurun_out/s2_*.
set -u
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
L=$OUT/s2_session.log
: > $L
t0=$(date +%s)
lap() { echo "== [$(( $(date +%s) - t0 )) s] $*" | tee -a $L; }

lap "1. parity suite"
timeout 1500 python -m pytest tests -q -m gpu -rfEs -p no:cacheprovider 2>&1 | tail -60 > $OUT/s2_pytest.log
tail -25 $OUT/s2_pytest.log | tee -a $L

lap "2. bench N=1, all legs"
timeout 900 python bench.py --steps 20 --warmup 3 > $OUT/s2_bench_n1.json 2> $OUT/s2_bench_n1.err
tail -c 2500 $OUT/s2_bench_n1.json | tee -a $L
tail -5 $OUT/s2_bench_n1.err | tee -a $L

lap "3. new defaults on the four configurations + knobs"
for cfg in "bn254 1 20" "bn254 2 20" "bls12-381 1 20" "bw6-761 1 18" "bls12-381 1 22"; do
  set -- $cfg
  timeout 400 python tools/sweep_msm.py $1 $2 $3 --reps 5 --set GB200_MSM_PERSISTENT=0,1 >> $OUT/s2_defaults.jsonl 2>> $OUT/s2_err.log
done
timeout 400 python tools/sweep_msm.py bn254 1 20 --reps 10 --set GB200_MSM_TASK_LEN=16,24,32,40,48,64 > $OUT/s2_tasklen.jsonl 2>> $OUT/s2_err.log
timeout 400 python tools/sweep_msm.py bn254 2 20 --reps 5 --set GB200_MSM_TASK_LEN=16,32,64 >> $OUT/s2_tasklen.jsonl 2>> $OUT/s2_err.log
timeout 400 python tools/sweep_msm.py bls12-381 1 20 --reps 5 --set GB200_MSM_TASK_LEN=16,32,64 >> $OUT/s2_tasklen.jsonl 2>> $OUT/s2_err.log
cat $OUT/s2_defaults.jsonl $OUT/s2_tasklen.jsonl | cut -c1-360 | tee -a $L

lap "4. NTT tiles"
timeout 400 python tools/sweep_ntt.py --curve bn254 --logs 20,22,24 --tiles 10,9,8,7 > $OUT/s2_ntt.jsonl 2>> $OUT/s2_err.log
timeout 400 python tools/sweep_ntt.py --curve bls12-381 --logs 22 --tiles 10,9,8,7 >> $OUT/s2_ntt.jsonl 2>> $OUT/s2_err.log
timeout 400 python tools/sweep_ntt.py --curve bw6-761 --logs 20 --tiles 10,9,8,7 >> $OUT/s2_ntt.jsonl 2>> $OUT/s2_err.log
cat $OUT/s2_ntt.jsonl | tee -a $L

lap "5. ncu"
# launch list of one bench step (shares), full captures of the three dominant kernels
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/s2_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-strong --no-groth16 --no-plonk --no-cpu > $OUT/s2_ncu_bench.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_msm_accumulate -c 1 -f -o $OUT/s2_ncu_accumulate_bn254_g1 \
    python tools/run_msm.py bn254 1 20 1 > $OUT/s2_ncu_g1.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_msm_accumulate -c 1 -f -o $OUT/s2_ncu_accumulate_bn254_g2 \
    python tools/run_msm.py bn254 2 20 1 > $OUT/s2_ncu_g2.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_msm_accumulate -c 1 -f -o $OUT/s2_ncu_accumulate_bls381_g1 \
    python tools/run_msm.py bls12-381 1 20 1 > $OUT/s2_ncu_bls.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_ntt_pass -c 3 -f -o $OUT/s2_ncu_ntt_pass \
    python -c "
import numpy as np, torch
from gnark_b200 import lib
lib.load(); lib.init([0])
d = lib.Domain(lib.BN254, 20)
x = torch.randint(0, 1 << 60, ((1 << 20) * 4,), dtype=torch.int64, device='cuda')
torch.cuda.synchronize()
d.ntt_async(x); lib.sync(0)
" > $OUT/s2_ncu_ntt.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/s2_launches_plonk.csv \
    python tools/run_plonk.py bls12-381 20 1 > $OUT/s2_ncu_plonk.log 2>&1
lap "6. PLONK stage times at 2^22 (no profiler)"
timeout 600 python tools/run_plonk.py bls12-381 22 4 2>&1 | tee -a $L
lap "done"
ls -la $OUT | tee -a $L
