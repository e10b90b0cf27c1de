#!/bin/bash
# 1 -> 8 GPU scaling of the flagship bench + fused comm check / sweep at 8 GPUs.
set -u
mkdir -p gpurun_out
S=gpurun_out/summary_scale8.txt
: > $S
NG=$(nvidia-smi -L | wc -l)
echo "gpus=$NG" | tee -a $S
for N in 1 2 4 8; do
  if [ $N -gt $NG ]; then continue; fi
  echo "=== bench N=$N" | tee -a $S
  if [ $N -eq 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 100 --warmup 10 > gpurun_out/scale_$N.log 2>&1
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29700+N)) bench.py --gpus $N --steps 100 --warmup 10 > gpurun_out/scale_$N.log 2>&1
  fi
  echo "exit=$? $(grep '^{' gpurun_out/scale_$N.log | tail -n 1 | cut -c1-260)" | tee -a $S
done
N=$NG
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "=== comm check (P2P) N=$N" | tee -a $S
timeout 400 $TR --master-port 29811 benchmarks/comm_check.py --check --sweep --max-bytes $((1<<30)) --out gpurun_out/comm_check_p2p_$N.json > gpurun_out/comm_p2p_$N.log 2>&1
echo "exit=$?" | tee -a $S; grep -E "check" gpurun_out/comm_p2p_$N.log | tail -2 | tee -a $S
echo "=== comm check (NVLS) N=$N" | tee -a $S
SRB_NVLS=1 timeout 400 $TR --master-port 29812 benchmarks/comm_check.py --check --sweep --max-bytes $((1<<30)) --out gpurun_out/comm_check_nvls_$N.json > gpurun_out/comm_nvls_$N.log 2>&1
echo "exit=$?" | tee -a $S; grep -E "check" gpurun_out/comm_nvls_$N.log | tail -2 | tee -a $S
echo "=== bench N=$N NVLS" | tee -a $S
SRB_NVLS=1 timeout 600 $TR --master-port 29813 bench.py --gpus $N --steps 100 --warmup 10 --no-e2e > gpurun_out/scale_nvls_$N.log 2>&1
echo "exit=$? $(grep '^{' gpurun_out/scale_nvls_$N.log | tail -n 1 | cut -c1-260)" | tee -a $S
echo "=== nccl-baseline N=$N" | tee -a $S
timeout 600 $TR --master-port 29814 bench.py --gpus $N --steps 10 --warmup 3 --impl nccl-baseline --engine eager --no-e2e > gpurun_out/scale_nccl_$N.log 2>&1
echo "exit=$? $(grep '^{' gpurun_out/scale_nccl_$N.log | tail -n 1 | cut -c1-260)" | tee -a $S
echo "=== actors + fused comm (CLI path), 2 GPU workers" | tee -a $S
timeout 600 python benchmarks/bench_rayproxy.py --workers 2 --gpu --mode sync --comm auto --steps 60 > gpurun_out/actors_fused.log 2>&1; echo "exit=$? $(grep '^{' gpurun_out/actors_fused.log | tail -n 1 | cut -c1-400)" | tee -a $S
