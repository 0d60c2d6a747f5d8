#!/usr/bin/env bash
# Round 2, multi-GPU session (gpurun --gpus N, N = 2, 4 or 8): the tests that need two devices, bench.py at every
# power of two up to N (weak + strong legs, Groth16 sharded, BW6-761 at 8), sharded PLONK and multi-GPU NTT.
# Every torchrun is wrapped in its own timeout: a hung collective must not eat the box.  Outputs: gpurun_out/s3_*.
set -u
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
NG=$(python -c "import torch; print(torch.cuda.device_count())")
L=$OUT/s3_session_n$NG.log
: > $L
t0=$(date +%s)
lap() { echo "== [$(( $(date +%s) - t0 )) s] $*" | tee -a $L; }
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"

lap "devices: $NG"
if [ "${S3_TESTS:-1}" = 1 ]; then
lap "1. tests that need two devices"
timeout 600 python -m pytest tests -q -m gpu -rfEs -p no:cacheprovider -k "two_devices or with_devices or sharded_two or comm or allreduce" 2>&1 | tail -15 | tee -a $L
fi

for n in ${S3_NS:-2 4 8}; do
  [ "$n" -le "$NG" ] || continue
  lap "2. bench N=$n"
  timeout ${S3_BENCH_TIMEOUT:-420} $TR --nproc-per-node $n --master-port $((29700 + n)) bench.py --gpus $n --steps 20 --warmup 3 ${S3_BENCH_FLAGS:-} \
      > $OUT/s3_bench_n$n.json 2> $OUT/s3_bench_n$n.err
  echo "rc=$?" | tee -a $L
  tail -c 2500 $OUT/s3_bench_n$n.json | tee -a $L
  grep -v "^\*\*\*\|OMP_NUM_THREADS\|^$" $OUT/s3_bench_n$n.err | tail -8 | tee -a $L
done

if [ "${S3_PLONK:-1}" = 1 ]; then
lap "3. sharded PLONK, N=$NG"
timeout 600 $TR --nproc-per-node $NG --master-port 29721 tools/bench_plonk_multi.py --log2n ${S3_PLONK_LOG:-22} --steps 2 \
    > $OUT/s3_plonk_n$NG.json 2> $OUT/s3_plonk_n$NG.err
echo "rc=$?" | tee -a $L
cut -c1-1500 $OUT/s3_plonk_n$NG.json | tee -a $L
grep -v "^\*\*\*\|OMP_NUM_THREADS\|^$" $OUT/s3_plonk_n$NG.err | tail -8 | tee -a $L
lap "4. multi-GPU NTT 2^24, N=$NG"
timeout 300 $TR --nproc-per-node $NG --master-port 29722 tools/bench_sharded_ntt.py --log2n 24 > $OUT/s3_ntt_n$NG.json 2> $OUT/s3_ntt_n$NG.err
echo "rc=$?" | tee -a $L
cat $OUT/s3_ntt_n$NG.json | tee -a $L
grep -v "^\*\*\*\|OMP_NUM_THREADS\|^$" $OUT/s3_ntt_n$NG.err | tail -5 | tee -a $L
fi
lap "done"
