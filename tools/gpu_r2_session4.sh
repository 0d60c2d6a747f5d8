#!/usr/bin/env bash
# Round 2, GPU session 4 (one B200): the lane-cooperative MSM tail and the wave-depth task length on hardware -
# parity suite, bench, stage times, the two remaining experiment knobs (L2 fill granularity, 3 blocks/SM for the
# BN254 G2 accumulate), ncu of the new tail kernels.  Outputs: gpurun_out/s4_*.
set -u
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
L=$OUT/s4_session.log
: > $L
t0=$(date +%s)
lap() { echo "== [$(( $(date +%s) - t0 )) s] $*" | tee -a $L; }

lap "1. parity suite"
timeout 1500 python -m pytest tests -q -m gpu -rfEs -p no:cacheprovider 2>&1 | tail -60 > $OUT/s4_pytest.log
tail -25 $OUT/s4_pytest.log | tee -a $L

lap "2. bench N=1, all legs"
timeout 900 python bench.py --steps 20 --warmup 3 > $OUT/s4_bench_n1.json 2> $OUT/s4_bench_n1.err
tail -c 2500 $OUT/s4_bench_n1.json | tee -a $L
tail -5 $OUT/s4_bench_n1.err | tee -a $L

lap "3. stage times with the new tail (defaults)"
for cfg in "bn254 1 20" "bn254 2 20" "bls12-381 1 20" "bls12-381 1 22" "bw6-761 1 18" "bn254 1 24"; do
  set -- $cfg
  timeout 400 python tools/sweep_msm.py $1 $2 $3 --reps 5 >> $OUT/s4_defaults.jsonl 2>> $OUT/s4_err.log
done
cut -c1-420 $OUT/s4_defaults.jsonl | tee -a $L

lap "4. experiments: task length on the wide curves, L2 fill granularity, 3 blocks/SM for BN254 G2"
timeout 400 python tools/sweep_msm.py bw6-761 1 18 --reps 5 --set GB200_MSM_TASK_LEN=16,24,32,48,64 > $OUT/s4_tasklen.jsonl 2>> $OUT/s4_err.log
timeout 400 python tools/sweep_msm.py bls12-381 1 22 --reps 5 --set GB200_MSM_TASK_LEN=32,48,64,96 >> $OUT/s4_tasklen.jsonl 2>> $OUT/s4_err.log
timeout 400 python tools/sweep_msm.py bn254 1 24 --reps 3 --set GB200_MSM_TASK_LEN=32,48,64,96 >> $OUT/s4_tasklen.jsonl 2>> $OUT/s4_err.log
for cfg in "bn254 1 20" "bn254 2 20" "bls12-381 1 22"; do
  set -- $cfg
  timeout 400 python tools/sweep_msm.py $1 $2 $3 --reps 10 --set GB200_MSM_ACC_STREAM=0,1 >> $OUT/s4_accstream.jsonl 2>> $OUT/s4_err.log
done
cut -c1-300 $OUT/s4_accstream.jsonl | tee -a $L
timeout 400 python tools/sweep_ntt.py --curve bn254 --logs 20,22,24 --tiles 8,9 > $OUT/s4_ntt.jsonl 2>> $OUT/s4_err.log
timeout 400 python tools/sweep_ntt.py --curve bls12-381 --logs 22 --tiles 8,9 >> $OUT/s4_ntt.jsonl 2>> $OUT/s4_err.log
cat $OUT/s4_ntt.jsonl | tee -a $L
for g in 32 64 128; do
  GB200_L2_FETCH=$g timeout 300 python tools/sweep_msm.py bn254 1 20 --reps 10 2>> $OUT/s4_err.log | sed "s/^{/{\"l2_fetch\": $g, /" >> $OUT/s4_l2fetch.jsonl
done
GB200_LIB=$PWD/gnark_b200/lib/libgnark_b200_accw3.so timeout 300 python tools/sweep_msm.py bn254 2 20 --reps 5 2>> $OUT/s4_err.log | sed 's/^{/{"lib": "accw3", /' > $OUT/s4_accw3.jsonl
cat $OUT/s4_tasklen.jsonl $OUT/s4_l2fetch.jsonl $OUT/s4_accw3.jsonl | cut -c1-420 | tee -a $L

lap "5. ncu"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/s4_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-strong --no-groth16 --no-plonk --no-cpu > $OUT/s4_ncu_bench.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k "regex:k_msm_(combine|reduce_chunks|set_sum|finish)" -c 5 -f -o $OUT/s4_ncu_tail_bn254_g1 \
    python tools/run_msm.py bn254 1 20 1 > $OUT/s4_ncu_tail.log 2>&1
for g in 32 64; do
  GB200_L2_FETCH=$g timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_msm_accumulate -c 1 --csv \
      --log-file $OUT/s4_ncu_l2fetch_$g.csv python tools/run_msm.py bn254 1 20 1 > $OUT/s4_ncu_l2_$g.log 2>&1
done
tail -3 $OUT/s4_ncu_l2fetch_32.csv $OUT/s4_ncu_l2fetch_64.csv | cut -c1-300 | tee -a $L
lap "done"
ls -la $OUT | tail -30 | tee -a $L
