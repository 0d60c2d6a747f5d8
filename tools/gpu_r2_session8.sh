#!/usr/bin/env bash
# Round 2, multi-GPU session 8 (gpurun --gpus 8): bench.py at N=8 (BW6-761 leg with a real base point) and N=4, and the
# two tests that need two devices in one process.  Every torchrun under its own timeout.  Outputs: gpurun_out/s8_*.
set -u
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
NG=$(python -c "import torch; print(torch.cuda.device_count())")
L=$OUT/s8_session.log
: > $L
t0=$(date +%s)
lap() { echo "== [$(( $(date +%s) - t0 )) s] $*" | tee -a $L; }
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
summ() {
python - "$1" <<'PY' | tee -a $L
import json, sys
d = json.load(open(sys.argv[1]))
print("N=%d value %.4g e2e %.4g ms/step %.3f clocks %s" % (d["n_gpus"], d["value"], d["e2e"]["value"], d["ms_per_step"], d.get("clocks")))
for k in ("strong", "groth16", "plonk", "bw6"):
    v = d.get(k)
    if v:
        print(k, {a: b for a, b in v.items() if not isinstance(b, (dict, list)) and len(str(b)) < 48})
        if "stage_ms_rank0" in v: print("   stage_ms_rank0", v["stage_ms_rank0"])
PY
}
lap "devices: $NG"
for n in ${S8_NS:-8 4}; do
  [ "$n" -le "$NG" ] || continue
  lap "bench N=$n"
  timeout ${S8_BENCH_TIMEOUT:-420} $TR --nproc-per-node $n --master-port $((29800 + n)) bench.py --gpus $n --steps 20 --warmup 3 \
      > $OUT/s8_bench_n$n.json 2> $OUT/s8_bench_n$n.err
  echo "rc=$?" | tee -a $L
  summ $OUT/s8_bench_n$n.json
  grep -v "^\*\*\*\|OMP_NUM_THREADS\|^$\|NCCL version" $OUT/s8_bench_n$n.err | tail -5 | tee -a $L
done
lap "tests that need two devices"
CUDA_VISIBLE_DEVICES=0,1 timeout 600 python -m pytest tests -q -m gpu -rfEs -p no:cacheprovider -k "two_devices or with_devices" 2>&1 | tail -8 | tee -a $L
lap "done"
