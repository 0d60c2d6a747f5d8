#!/usr/bin/env bash
# Round 2, GPU session 11 (one B200): PLONK non-MSM work - pre-scaled blinding coefficients, L1 denominators cached in
# the domain handle, constraint kernel at 4 x 128 threads per SM.  Parity of the PLONK tests (cache on and off), proof
# times, launch list.  Outputs: gpurun_out/s11_*.
set -u
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
L=$OUT/s11_session.log
: > $L
t0=$(date +%s)
lap() { echo "== [$(( $(date +%s) - t0 )) s] $*" | tee -a $L; }
lap "1. PLONK parity (denominator cache on)"
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider -k "plonk or quotient or vec_ops or scans" 2>&1 | tail -4 | tee -a $L
lap "2. PLONK parity (GB200_PLONK_DEN_CACHE=0)"
GB200_PLONK_DEN_CACHE=0 timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider -k "plonk_prove or quotient" 2>&1 | tail -4 | tee -a $L
lap "3. proof times, BLS12-381 2^22 and BN254 2^20"
timeout 300 python tools/run_plonk.py bls12-381 22 4 2>&1 | tail -6 | tee -a $L
timeout 300 python tools/run_plonk.py bn254 20 4 2>&1 | tail -6 | tee -a $L
lap "4. launch list of one proof"
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $OUT/s11_launches_plonk.csv \
    python tools/run_plonk.py bls12-381 22 2 > $OUT/s11_plonk_ncu.log 2>&1
python tools/launch_totals.py $OUT/s11_launches_plonk.csv --from k_msm_decompose --nth -10 --back 16 | head -16 | tee -a $L
lap "done"
