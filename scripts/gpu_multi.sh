#!/bin/bash
# multi-GPU validation: fused comm check, bandwidth sweep, bench at N GPUs.  usage: gpu_multi.sh N
set -u
N=${1:-2}
mkdir -p gpurun_out
S=gpurun_out/summary_multi_$N.txt
: > $S
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "=== comm check (P2P) N=$N" | tee -a $S
timeout 300 $TR --master-port 29611 benchmarks/comm_check.py --check --sweep --max-bytes $((1<<28)) --out gpurun_out/comm_check_p2p_$N.json > gpurun_out/comm_p2p_$N.log 2>&1
echo "exit=$?" | tee -a $S; grep -E "check|bytes" gpurun_out/comm_p2p_$N.log | tail -24 | tee -a $S
echo "=== comm check (NVLS) N=$N" | tee -a $S
SRB_NVLS=1 timeout 300 $TR --master-port 29612 benchmarks/comm_check.py --check --sweep --max-bytes $((1<<28)) --out gpurun_out/comm_check_nvls_$N.json > gpurun_out/comm_nvls_$N.log 2>&1
echo "exit=$?" | tee -a $S; grep -E "check|bytes|Error|error" gpurun_out/comm_nvls_$N.log | tail -24 | tee -a $S
echo "=== bench N=1" | tee -a $S
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_1.log 2>&1; echo "exit=$? $(tail -n 1 gpurun_out/bench_1.log | cut -c1-600)" | tee -a $S
echo "=== bench N=$N (fused)" | tee -a $S
timeout 600 $TR --master-port 29613 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_$N.log 2>&1; echo "exit=$? $(grep '^{' gpurun_out/bench_$N.log | tail -n 1 | cut -c1-600)" | tee -a $S
echo "=== bench N=$N (nccl-baseline)" | tee -a $S
timeout 600 $TR --master-port 29614 bench.py --gpus $N --steps 10 --warmup 3 --impl nccl-baseline --engine eager --no-e2e > gpurun_out/bench_nccl_$N.log 2>&1; echo "exit=$? $(grep '^{' gpurun_out/bench_nccl_$N.log | tail -n 1 | cut -c1-400)" | tee -a $S
tail -5 gpurun_out/bench_$N.log | cut -c1-300
