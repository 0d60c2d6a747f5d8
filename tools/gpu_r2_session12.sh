#!/usr/bin/env bash
# Round 2, GPU session 12 (one B200): final state - full parity suite, bench N=1, PLONK proof times with the pipelined
# commitments of the L/R/O, H1..3 and linearise rounds, launch lists.  Outputs: gpurun_out/s12_*.
set -u
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
L=$OUT/s12_session.log
: > $L
t0=$(date +%s)
lap() { echo "== [$(( $(date +%s) - t0 )) s] $*" | tee -a $L; }
lap "1. parity suite"
timeout 1500 python -m pytest tests -q -m gpu -rfEs -p no:cacheprovider --durations=5 2>&1 | tail -30 > $OUT/s12_pytest.log
tail -22 $OUT/s12_pytest.log | tee -a $L
lap "2. bench N=1, all legs"
timeout 900 python bench.py --steps 20 --warmup 3 > $OUT/s12_bench_n1.json 2> $OUT/s12_bench_n1.err
echo "rc=$?" | tee -a $L
python - <<'PY' | tee -a $L
import json
d = json.load(open("gpurun_out/s12_bench_n1.json"))
print("value %.4g e2e %.4g ms/step %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]))
for k in ("strong", "groth16", "plonk"):
    v = d.get(k, {})
    print(k, {a: b for a, b in v.items() if not isinstance(b, (dict, list)) and len(str(b)) < 60})
print("plonk stages", d.get("plonk", {}).get("stage_ms"))
PY
tail -3 $OUT/s12_bench_n1.err | tee -a $L
lap "3. PLONK proof times (pageable inputs)"
timeout 300 python tools/run_plonk.py bls12-381 22 4 2>&1 | tail -5 | tee -a $L
timeout 300 python tools/run_plonk.py bn254 20 4 2>&1 | tail -5 | tee -a $L
lap "4. PLONK under compute-sanitizer (small size: memcheck of the pipelined commitments)"
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python tools/run_plonk.py bn254 10 1 2>&1 | tail -6 | tee -a $L
lap "5. launch lists"
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $OUT/s12_launches_plonk.csv \
    python tools/run_plonk.py bls12-381 22 2 > $OUT/s12_plonk_ncu.log 2>&1
python tools/launch_totals.py $OUT/s12_launches_plonk.csv --from k_msm_decompose --nth -10 --back 16 | head -14 | tee -a $L
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/s12_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-strong --no-groth16 --no-plonk --no-cpu > $OUT/s12_ncu_bench.log 2>&1
lap "done"
