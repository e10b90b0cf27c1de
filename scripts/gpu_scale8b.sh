#!/bin/bash
# 8-GPU box: flagship at N = 8, 4, 2, 1 (with e2e), other configs at N = 8, fused-comm correctness check
set -u
mkdir -p gpurun_out
S=gpurun_out/summary_scale8b.txt
: > $S
NG=$(nvidia-smi -L | wc -l)
echo "gpus=$NG" | tee -a $S
run() {  # name N extra-args...
  local name=$1 N=$2; shift 2
  if [ $N -gt $NG ]; then return; fi
  if [ $N -eq 1 ]; then
    timeout -s KILL 400 python bench.py --gpus 1 "$@" > gpurun_out/$name.log 2>&1
  else
    timeout -s KILL 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29700 + RANDOM % 200)) bench.py --gpus $N "$@" > gpurun_out/$name.log 2>&1
  fi
  echo "$name exit=$? $(grep '^{' gpurun_out/$name.log | tail -n 1 | cut -c1-220)" | tee -a $S
}
run scale2_8 8 --steps 100 --warmup 10
run scale2_4 4 --steps 100 --warmup 10
run scale2_2 2 --steps 100 --warmup 10
run scale2_1 1 --steps 100 --warmup 10
run scale2_multitask_8 8 --steps 30 --warmup 5 --config configs/multitask_w512.cfg
run scale2_parser_8 8 --steps 30 --warmup 5 --config configs/parser_w256.cfg
run scale2_tagger_8 8 --steps 30 --warmup 5 --config configs/tagger_w96.cfg
echo "=== comm check N=$NG" | tee -a $S
timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29911 benchmarks/comm_check.py --check > gpurun_out/comm_check2_$NG.log 2>&1
echo "exit=$?" | tee -a $S; grep -E "check" gpurun_out/comm_check2_$NG.log | tail -3 | tee -a $S
