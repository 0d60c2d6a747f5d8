#!/usr/bin/env bash
# Round 2, GPU session 7 (one B200), after the carry fix of mont_reduce_wide (lazily reduced Fp2 product): the Groth16
# diagnosis that found it, the full parity suite (StatisticalZK, the known-failure regressions), bench N=1, G2 stage
# times, the randomised soak, the Groth16 step profile, ncu launch list + G2 accumulate capture.  Outputs: gpurun_out/s7_*.
set -u
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
L=$OUT/s7_session.log
: > $L
t0=$(date +%s)
lap() { echo "== [$(( $(date +%s) - t0 )) s] $*" | tee -a $L; }

lap "0. Groth16 under bench.py's configurations"
timeout 300 python tools/diag_groth16.py 20 2>&1 | tail -12 | tee -a $L

lap "1. parity suite"
timeout 1500 python -m pytest tests -q -m gpu -rfEs -p no:cacheprovider --durations=6 2>&1 | tail -40 > $OUT/s7_pytest.log
tail -25 $OUT/s7_pytest.log | tee -a $L

lap "2. bench N=1, all legs"
timeout 900 python bench.py --steps 20 --warmup 3 > $OUT/s7_bench_n1.json 2> $OUT/s7_bench_n1.err
echo "rc=$?" | tee -a $L
python - <<'PY' | tee -a $L
import json
d = json.load(open("gpurun_out/s7_bench_n1.json"))
print("value %.4g e2e %.4g ms/step %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]))
print("stage_ms", d["stage_ms"])
for k in ("strong", "groth16", "plonk"):
    v = d.get(k, {})
    print(k, {a: b for a, b in v.items() if not isinstance(b, (dict, list)) and len(str(b)) < 60})
print("cpu_baseline", d.get("cpu_baseline"))
PY
tail -3 $OUT/s7_bench_n1.err | tee -a $L

lap "3. stage times, G2 and the others"
for cfg in "bn254 1 20" "bn254 2 20" "bls12-381 2 20" "bls12-381 1 22"; do
  set -- $cfg
  timeout 400 python tools/sweep_msm.py $1 $2 $3 --reps 5 >> $OUT/s7_msm.jsonl 2>> $OUT/s7_err.log
done
cut -c1-420 $OUT/s7_msm.jsonl | tee -a $L

lap "4. randomised parity soak"
timeout 300 python tools/fuzz_gpu.py --seconds ${S7_FUZZ_S:-150} --seed 7 > $OUT/s7_fuzz.jsonl 2> $OUT/s7_fuzz.err
echo "rc=$?" | tee -a $L
cat $OUT/s7_fuzz.jsonl | tee -a $L
tail -3 $OUT/s7_fuzz.err | tee -a $L

lap "5. Groth16 step profile"
timeout 300 python tools/run_groth16.py bn254 20 3 2>&1 | tail -30 | tee -a $L

lap "6. ncu"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/s7_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-strong --no-groth16 --no-plonk --no-cpu > $OUT/s7_ncu_bench.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_msm_accumulate -c 1 -f -o $OUT/s7_ncu_accumulate_bn254_g2 \
    python tools/run_msm.py bn254 2 20 1 > $OUT/s7_ncu_g2.log 2>&1
lap "done"
ls -la $OUT | grep s7_ | tee -a $L
