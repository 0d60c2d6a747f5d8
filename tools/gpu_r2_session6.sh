#!/usr/bin/env bash
# Round 2, GPU session 6 (one B200): the full parity suite incl. the new full-size Groth16 proof (2^20, Verify),
# bench.py N=1 with the verified Groth16 leg, and the randomised GPU parity soak.  Outputs: gpurun_out/s6_*.
set -u
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
L=$OUT/s6_session.log
: > $L
t0=$(date +%s)
lap() { echo "== [$(( $(date +%s) - t0 )) s] $*" | tee -a $L; }

lap "1. parity suite"
timeout 1500 python -m pytest tests -q -m gpu -rfEs -p no:cacheprovider --durations=8 2>&1 | tail -40 > $OUT/s6_pytest.log
tail -30 $OUT/s6_pytest.log | tee -a $L

lap "2. bench N=1, all legs"
timeout 900 python bench.py --steps 20 --warmup 3 > $OUT/s6_bench_n1.json 2> $OUT/s6_bench_n1.err
echo "rc=$?" | tee -a $L
tail -c 1500 $OUT/s6_bench_n1.json | tee -a $L
tail -5 $OUT/s6_bench_n1.err | tee -a $L

lap "3. randomised parity soak"
timeout 400 python tools/fuzz_gpu.py --seconds ${S6_FUZZ_S:-240} --seed 6 > $OUT/s6_fuzz.jsonl 2> $OUT/s6_fuzz.err
echo "rc=$?" | tee -a $L
cat $OUT/s6_fuzz.jsonl | tee -a $L
tail -5 $OUT/s6_fuzz.err | tee -a $L
lap "done"
