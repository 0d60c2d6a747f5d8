#!/usr/bin/env bash
# Round 2, GPU session 15 (one B200): A/B of the radix-4 NTT pass (GB200_NTT_RADIX4=1: two stages per shared-memory
# round trip) - parity of every NTT test with it, then ms per transform with and without.  Outputs: gpurun_out/s15_*.
set -u
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
L=$OUT/s15_session.log
: > $L
t0=$(date +%s)
lap() { echo "== [$(( $(date +%s) - t0 )) s] $*" | tee -a $L; }
lap "1. NTT parity with GB200_NTT_RADIX4=1"
GB200_NTT_RADIX4=1 timeout 600 python -m pytest tests -q -m gpu -p no:cacheprovider -k "ntt or compute_h or quotient" 2>&1 | tail -4 | tee -a $L
lap "2. timing"
for r4 in 0 1; do
  for cfg in "bn254 20,22,24" "bls12-381 22,24" "bw6-761 20"; do
    set -- $cfg
    for tiles in 8 9; do
      GB200_NTT_RADIX4=$r4 timeout 300 python tools/sweep_ntt.py --curve $1 --logs $2 --tiles $tiles --reps 30 2>>$OUT/s15_err.log | sed "s/^{/{\"radix4\": $r4, /" >> $OUT/s15_ntt.jsonl
    done
  done
done
cat $OUT/s15_ntt.jsonl | cut -c1-140 | tee -a $L
lap "done"
