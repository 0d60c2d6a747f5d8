#!/usr/bin/env bash
# One gpurun call that validates everything written without a GPU and A/B-tests the opt-in kernels:
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/gpu_session.sh 1 2'        # parity + bench        (~10 min)
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/gpu_session.sh 3v'         # arithmetic variants    (~8 min)
#   /usr/local/graft/bin/gpurun --timeout 1800 -- 'bash tools/gpu_session.sh 3 3b'       # knob sweeps            (~20 min)
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/gpu_session.sh 2b 4'       # other configs + ncu
#   /usr/local/graft/bin/gpurun --gpus 2 --timeout 1500 -- 'bash tools/gpu_session.sh 5' # multi-GPU
# Stages: 1 parity, 2 bench, 2b configs, 3 MSM knob sweeps, 3v compile-time variants, 3b NTT tiles, 4 ncu, 5 multi-GPU;
# no argument = all of them.  Outputs land in gpurun_out/ (copy what is worth keeping into profiles/).
set -u
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
export GB200_RUN_EXPERIMENTS=1     # tests/test_gpu_zz_late.py: run the opt-in experiment tests too
STAGES=" ${*:-1 2 2b 3 3v 3b 4 5} "
want() { [[ "$STAGES" == *" $1 "* ]]; }
: >> $OUT/session.log

if want 1; then
echo "== 1. parity suite (validated part first, then the late file)" | tee -a $OUT/session.log
timeout 900 python -m pytest tests -q -m gpu -x --deselect tests/test_gpu_zz_late.py 2>&1 | tail -5 | tee -a $OUT/session.log
timeout 600 python -m pytest tests/test_gpu_zz_late.py -q -m gpu -rxX 2>&1 | tail -40 | tee -a $OUT/session.log
fi

if want 2; then
echo "== 2. bench, 1 GPU (with the asynchronous host path)" | tee -a $OUT/session.log
timeout 600 python bench.py --steps 20 --warmup 3 --e2e-submit > $OUT/bench_n1.json 2> $OUT/bench_n1.err
tail -c 2000 $OUT/bench_n1.json | tee -a $OUT/session.log
# the Groth16 leg's pageable-input variant with threaded staging of the uploads
GB200_STAGE_THREADS=4 timeout 600 python bench.py --steps 5 --warmup 3 > $OUT/bench_n1_stage4.json 2>> $OUT/bench_n1.err
python -c "import json;d=json.load(open('$OUT/bench_n1_stage4.json'));print('groth16 with GB200_STAGE_THREADS=4:', d.get('groth16'))" | tee -a $OUT/session.log

fi
if want 2b; then
echo "== 2b. other configs on this GPU count: sharded MSMs (2^22 total), PLONK building blocks and the end-to-end PLONK prove" | tee -a $OUT/session.log
timeout 900 python tools/bench_configs.py --total-log 22 --steps 3 > $OUT/configs.jsonl 2>> $OUT/session.err
cut -c1-600 $OUT/configs.jsonl | tee -a $OUT/session.log

fi
if want 3; then
echo "== 3. MSM knob sweeps (every configuration is checked against the known-dlog oracle)" | tee -a $OUT/session.log
timeout 600 python tools/sweep_msm.py bn254 1 20 --set GB200_MSM_WINDOW=16,17,18,19,20,22 > $OUT/sweep_bn254_window.jsonl 2>> $OUT/session.err
# persistent accumulate (grid = SMs x resident blocks, tasks from an atomic counter): removes the partial last wave and
# the lanes idling behind the short last task of every bucket
for cfg in "bn254 1 20" "bn254 2 20" "bls12-381 1 20" "bw6-761 1 18"; do
  set -- $cfg
  timeout 600 python tools/sweep_msm.py $1 $2 $3 --set GB200_MSM_PERSISTENT=0,1,2 >> $OUT/sweep_persistent.jsonl 2>> $OUT/session.err
done
# accumulators in shared memory (more resident warps for the wide fields: BN254 G2 96 instead of 144 registers, BLS12-381
# G1 96 instead of 126)
for cfg in "bn254 2 20" "bls12-381 1 20" "bls12-381 2 18" "bw6-761 1 18"; do
  set -- $cfg
  timeout 600 python tools/sweep_msm.py $1 $2 $3 --set GB200_MSM_SMEM_ACC=0,1 >> $OUT/sweep_smem_acc.jsonl 2>> $OUT/session.err
done
# the PLONK config's MSM (2^22 points, BLS12-381): ten of these per proof
timeout 900 python tools/sweep_msm.py bls12-381 1 22 --reps 3 --set GB200_MSM_WINDOW=16,18,20,22 > $OUT/sweep_bls381_2p22_window.jsonl 2>> $OUT/session.err
timeout 600 python tools/sweep_msm.py bn254 1 20 --set GB200_MSM_HYBRID=0,12,25,38,50,62 > $OUT/sweep_bn254_hybrid.jsonl 2>> $OUT/session.err
timeout 600 python tools/sweep_msm.py bls12-381 1 20 --set GB200_MSM_HYBRID=0,25,38,50,62,75 > $OUT/sweep_bls381_hybrid.jsonl 2>> $OUT/session.err
timeout 600 python tools/sweep_msm.py bn254 2 20 --set GB200_MSM_WINDOW=14,16,18 > $OUT/sweep_bn254_g2_window.jsonl 2>> $OUT/session.err
timeout 600 python tools/sweep_msm.py bn254 1 20 --set GB200_MSM_BATCH_AFFINE=0,2,4,5,6 > $OUT/sweep_bn254_ba.jsonl 2>> $OUT/session.err
timeout 600 python tools/sweep_msm.py bn254 2 20 --set GB200_MSM_BATCH_AFFINE=0,3,5 > $OUT/sweep_bn254_g2_ba.jsonl 2>> $OUT/session.err
timeout 600 python tools/sweep_msm.py bls12-381 1 20 --set GB200_MSM_BATCH_AFFINE=0,3,5 > $OUT/sweep_bls381_ba.jsonl 2>> $OUT/session.err
timeout 600 python tools/sweep_msm.py bw6-761 1 18 --set GB200_MSM_HYBRID=0,30,50 > $OUT/sweep_bw6_hybrid.jsonl 2>> $OUT/session.err
timeout 600 python tools/sweep_msm.py bw6-761 1 18 --set GB200_MSM_BATCH_AFFINE=0,4 > $OUT/sweep_bw6_ba.jsonl 2>> $OUT/session.err
fi
if want 3v; then
echo "== 3v. compile-time arithmetic / calling-convention variants" | tee -a $OUT/session.log
# compile-time arithmetic variants (build before the call, they travel with the snapshot):
#   make -C gnark_b200/csrc opt                                              -> libgnark_b200_opt.so (all three)
#   make -C gnark_b200/csrc opt OPTFLAGS=-DGB200_MONT_SQR OPTNAME=sqr         -> libgnark_b200_sqr.so
#   make -C gnark_b200/csrc opt OPTFLAGS=-DGB200_FP2_LAZY OPTNAME=lazy        -> libgnark_b200_lazy.so
#   make -C gnark_b200/csrc opt OPTFLAGS=-DGB200_MONT_KARATSUBA OPTNAME=kara  -> libgnark_b200_kara.so
#   make -C gnark_b200/csrc opt OPTFLAGS=-DGB200_CALL_BYVAL OPTNAME=byval     -> libgnark_b200_byval.so
#     (out-of-line field products take operands in registers: no stack round trip per call, more registers per thread)
#   make -C gnark_b200/csrc opt OPTFLAGS=-DGB200_INLINE_LIMBS=12 OPTNAME=inl12 -> 12-limb products (BLS12-381/377 Fp) inlined like
#     the 8-limb ones: statically 110 registers instead of 126, no CALL, stack 288 B instead of 768 B, 8.6 k instructions
#   make -C gnark_b200/csrc opt OPTFLAGS=-DGB200_INLINE_FP2 OPTNAME=inlfp2     -> Fp2 product / square inlined (G2)
#   make -C gnark_b200/csrc opt OPTFLAGS="-DGB200_MONT_SQR -DGB200_XYZZ_LAZY -DGB200_ACC_MIN_BLOCKS=5" OPTNAME=sqrxlazy5
#     (BN254 G1 accumulate: -9.4 % IMAD.WIDE at 96 registers - the first candidate, profiles/r01_sass_stats.md)
#   make -C gnark_b200/csrc opt OPTFLAGS=-DGB200_ACC_PREFETCH OPTNAME=prefetch -> L2 prefetch of the next gathered point
#   or all of them:  make -j8 -C gnark_b200/csrc variants      (about an hour on 8 cores; `variants-top`: the four first
#   candidates sqrxlazy5 / prefetch / inl12 / lazy in ~10 min)
# each library runs only the configurations its flags can change (a run = table upload + precompute + 10 MSMs, ~15 s)
for lib in gnark_b200/lib/libgnark_b200_*.so; do
  [ -f "$lib" ] || continue
  tag=$(basename $lib .so | sed 's/libgnark_b200_//')
  case "$tag" in
    sqr|xlazy|sqrxlazy|sqrxlazy5|kara8) cfgs=("bn254 1 20") ;;
    lazy|inlfp2)                        cfgs=("bn254 2 20") ;;
    inl12)                              cfgs=("bls12-381 1 20") ;;
    kara)                               cfgs=("bls12-381 1 20" "bw6-761 1 18") ;;
    byval)                              cfgs=("bn254 2 20" "bls12-381 1 20" "bw6-761 1 18") ;;
    *)                                  cfgs=("bn254 1 20" "bn254 2 20" "bls12-381 1 20" "bw6-761 1 18") ;;   # opt, prefetch
  esac
  for cfg in "${cfgs[@]}"; do
    set -- $cfg
    GB200_LIB=$PWD/$lib timeout 300 python tools/sweep_msm.py $1 $2 $3 \
        | sed "s/^{/{\"lib\": \"$tag\", /" >> $OUT/sweep_optlib.jsonl 2>> $OUT/session.err
  done
done
# the default library on the same four configurations, same script, for the comparison
for cfg in "bn254 1 20" "bn254 2 20" "bls12-381 1 20" "bw6-761 1 18"; do
  set -- $cfg
  timeout 300 python tools/sweep_msm.py $1 $2 $3 | sed "s/^{/{\"lib\": \"default\", /" >> $OUT/sweep_optlib.jsonl 2>> $OUT/session.err
done
fi
cat $OUT/sweep_*.jsonl 2>/dev/null | cut -c1-400 | tee -a $OUT/session.log

if want 3b; then
echo "== 3b. NTT tile sizes" | tee -a $OUT/session.log
timeout 600 python tools/sweep_ntt.py --curve bn254 --logs 20,22,24 > $OUT/sweep_ntt.jsonl 2>> $OUT/session.err
timeout 600 python tools/sweep_ntt.py --curve bls12-381 --logs 22 >> $OUT/sweep_ntt.jsonl 2>> $OUT/session.err
cat $OUT/sweep_ntt.jsonl | tee -a $OUT/session.log

fi
if want 4; then
echo "== 4. ncu: launch list of one Groth16-sized step and full captures of the NTT pass and the G2 accumulate" | tee -a $OUT/session.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_g2.csv \
    python tools/run_msm.py bn254 2 20 1 > $OUT/ncu_g2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_msm_accumulate -c 1 -o $OUT/ncu_g2_accumulate \
    python tools/run_msm.py bn254 2 20 1 >> $OUT/ncu_g2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ntt_pass -c 2 -o $OUT/ncu_ntt_pass \
    python -c "
import numpy as np, torch
from gnark_b200 import lib
lib.load(); lib.init([0])
d = lib.Domain(lib.BN254, 22)
x = torch.randint(0, 1 << 60, ((1 << 22) * 4,), dtype=torch.int64, device='cuda')
d.ntt_async(x); lib.sync(0)
" > $OUT/ncu_ntt.log 2>&1
fi
if want 5; then
echo "== 5. (only under gpurun --gpus 2/4/8) sharded NTT and the sharded configs" | tee -a $OUT/session.log
NG=$(python -c "import torch; print(torch.cuda.device_count())")
if [ "$NG" -gt 1 ]; then
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29611 \
      tools/bench_sharded_ntt.py --log2n 24 > $OUT/sharded_ntt_n$NG.json 2>> $OUT/session.err
  cat $OUT/sharded_ntt_n$NG.json | tee -a $OUT/session.log
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29613 \
      tools/bench_plonk_multi.py --log2n 22 > $OUT/plonk_n$NG.json 2>> $OUT/session.err
  cut -c1-800 $OUT/plonk_n$NG.json | tee -a $OUT/session.log
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29612 \
      bench.py --gpus $NG --steps 20 --warmup 3 > $OUT/bench_n$NG.json 2>> $OUT/session.err
  tail -c 1500 $OUT/bench_n$NG.json | tee -a $OUT/session.log
  # where does a sharded Groth16 proof spend its time (per-stage laps of every rank on stderr)
  GB200_STEP_PROFILE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 \
      --master-port 29614 bench.py --gpus $NG --steps 3 --warmup 3 > /dev/null 2> $OUT/groth16_steps_n$NG.err
  grep "gb200 step" $OUT/groth16_steps_n$NG.err | tail -40 | tee -a $OUT/session.log
fi
fi
ls -la $OUT | tee -a $OUT/session.log
