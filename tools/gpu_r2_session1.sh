#!/usr/bin/env bash
# Round 2, GPU session 1 (one B200): parity suite with every guard removed, bench N=1 with all legs, A/B of every
# compile-time variant and run-time knob left over from round 1, ncu captures.  Outputs: gpurun_out/s1_*.
set -u
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
L=$OUT/s1_session.log
: > $L
t0=$(date +%s)
lap() { echo "== [$(( $(date +%s) - t0 )) s] $*" | tee -a $L; }

lap "1. parity suite (experiments enabled)"
GB200_RUN_EXPERIMENTS=1 timeout 1500 python -m pytest tests -q -m gpu -rfEs -p no:cacheprovider 2>&1 | tail -150 > $OUT/s1_pytest.log
tail -30 $OUT/s1_pytest.log | tee -a $L

lap "2. bench N=1, all legs"
timeout 900 python bench.py --steps 20 --warmup 3 > $OUT/s1_bench_n1.json 2> $OUT/s1_bench_n1.err
tail -c 3000 $OUT/s1_bench_n1.json | tee -a $L
tail -5 $OUT/s1_bench_n1.err | tee -a $L

lap "3. compile-time variants"
run_variant() {  # tag cfg...
  tag=$1; shift
  for cfg in "$@"; do
    set -- $cfg
    if [ "$tag" = default ]; then
      timeout 240 python tools/sweep_msm.py $1 $2 $3 --reps 5 | sed "s/^{/{\"lib\": \"$tag\", /" >> $OUT/s1_variants.jsonl 2>> $OUT/s1_err.log
    else
      GB200_LIB=$PWD/gnark_b200/lib/libgnark_b200_$tag.so timeout 240 python tools/sweep_msm.py $1 $2 $3 --reps 5 \
        | sed "s/^{/{\"lib\": \"$tag\", /" >> $OUT/s1_variants.jsonl 2>> $OUT/s1_err.log
    fi
  done
}
run_variant default "bn254 1 20" "bn254 2 20" "bls12-381 1 20" "bw6-761 1 18"
run_variant sqrxlazy5 "bn254 1 20"
run_variant sqrxlazy "bn254 1 20" "bn254 2 20" "bls12-381 1 20"
run_variant kara8 "bn254 1 20" "bn254 2 20"
run_variant prefetch "bn254 1 20" "bn254 2 20" "bls12-381 1 20"
run_variant inl12 "bls12-381 1 20"
run_variant kara "bls12-381 1 20" "bw6-761 1 18"
run_variant byval "bn254 2 20" "bls12-381 1 20" "bw6-761 1 18"
run_variant lazy "bn254 2 20"
run_variant inlfp2 "bn254 2 20"
run_variant lazyinl "bn254 2 20"
cut -c1-330 $OUT/s1_variants.jsonl | tee -a $L

lap "4. run-time knobs"
timeout 400 python tools/sweep_msm.py bn254 1 20 --reps 5 --set GB200_MSM_WINDOW=15,16,17,18,19,20 > $OUT/s1_knob_window.jsonl 2>> $OUT/s1_err.log
timeout 400 python tools/sweep_msm.py bn254 1 20 --reps 5 --set GB200_MSM_PERSISTENT=0,1,2 > $OUT/s1_knob_persistent.jsonl 2>> $OUT/s1_err.log
timeout 400 python tools/sweep_msm.py bn254 2 20 --reps 5 --set GB200_MSM_PERSISTENT=0,1,2 >> $OUT/s1_knob_persistent.jsonl 2>> $OUT/s1_err.log
timeout 400 python tools/sweep_msm.py bls12-381 1 20 --reps 5 --set GB200_MSM_PERSISTENT=0,1,2 >> $OUT/s1_knob_persistent.jsonl 2>> $OUT/s1_err.log
timeout 400 python tools/sweep_msm.py bw6-761 1 18 --reps 5 --set GB200_MSM_PERSISTENT=0,1,2 >> $OUT/s1_knob_persistent.jsonl 2>> $OUT/s1_err.log
for cfg in "bn254 2 20" "bls12-381 1 20" "bw6-761 1 18"; do
  set -- $cfg
  timeout 400 python tools/sweep_msm.py $1 $2 $3 --reps 5 --set GB200_MSM_SMEM_ACC=0,1 >> $OUT/s1_knob_smem.jsonl 2>> $OUT/s1_err.log
done
timeout 400 python tools/sweep_msm.py bn254 1 20 --reps 5 --set GB200_MSM_HYBRID=0,12,25,38,50 > $OUT/s1_knob_hybrid.jsonl 2>> $OUT/s1_err.log
timeout 400 python tools/sweep_msm.py bls12-381 1 20 --reps 5 --set GB200_MSM_HYBRID=0,25,38,50,62 >> $OUT/s1_knob_hybrid.jsonl 2>> $OUT/s1_err.log
timeout 400 python tools/sweep_msm.py bn254 1 20 --reps 5 --set GB200_MSM_BATCH_AFFINE=0,2,4,5,6 > $OUT/s1_knob_ba.jsonl 2>> $OUT/s1_err.log
timeout 400 python tools/sweep_msm.py bn254 2 20 --reps 5 --set GB200_MSM_BATCH_AFFINE=0,3,5 >> $OUT/s1_knob_ba.jsonl 2>> $OUT/s1_err.log
timeout 400 python tools/sweep_msm.py bls12-381 1 20 --reps 5 --set GB200_MSM_BATCH_AFFINE=0,3,5 >> $OUT/s1_knob_ba.jsonl 2>> $OUT/s1_err.log
timeout 400 python tools/sweep_msm.py bn254 1 20 --reps 5 --set GB200_MSM_TASK_LEN=32,64,128 > $OUT/s1_knob_tasklen.jsonl 2>> $OUT/s1_err.log
timeout 600 python tools/sweep_msm.py bls12-381 1 22 --reps 3 --set GB200_MSM_WINDOW=16,18,20 > $OUT/s1_knob_bls_2p22.jsonl 2>> $OUT/s1_err.log
cat $OUT/s1_knob_*.jsonl | cut -c1-330 | tee -a $L

lap "5. NTT tiles / radix-8"
timeout 400 python tools/sweep_ntt.py --curve bn254 --logs 20,22,24 > $OUT/s1_ntt.jsonl 2>> $OUT/s1_err.log
timeout 400 python tools/sweep_ntt.py --curve bls12-381 --logs 22 >> $OUT/s1_ntt.jsonl 2>> $OUT/s1_err.log
cat $OUT/s1_ntt.jsonl | tee -a $L

lap "6. ncu: NTT pass (full), G2 accumulate (full)"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_ntt_pass -c 2 -f -o $OUT/s1_ncu_ntt_pass \
    python -c "
import numpy as np, torch
from gnark_b200 import lib
lib.load(); lib.init([0])
d = lib.Domain(lib.BN254, 20)
x = torch.randint(0, 1 << 60, ((1 << 20) * 4,), dtype=torch.int64, device='cuda')
torch.cuda.synchronize()
d.ntt_async(x); lib.sync(0)
" > $OUT/s1_ncu_ntt.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_msm_accumulate -c 1 -f -o $OUT/s1_ncu_g2_accumulate \
    python tools/run_msm.py bn254 2 20 1 > $OUT/s1_ncu_g2.log 2>&1
lap "done"
ls -la $OUT | tee -a $L
