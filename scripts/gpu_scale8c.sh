#!/bin/bash
# refresh of the flagship scaling points after the halo GEMMs: N = 8 and N = 1 on the same box
set -u
mkdir -p gpurun_out
S=gpurun_out/summary_scale8c.txt
: > $S
timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29741 bench.py --gpus 8 --steps 100 --warmup 10 > gpurun_out/scale3_8.log 2>&1
echo "N=8 exit=$? $(grep '^{' gpurun_out/scale3_8.log | tail -n 1 | cut -c1-220)" | tee -a $S
timeout -s KILL 300 python bench.py --gpus 1 --steps 100 --warmup 10 > gpurun_out/scale3_1.log 2>&1
echo "N=1 exit=$? $(grep '^{' gpurun_out/scale3_1.log | tail -n 1 | cut -c1-220)" | tee -a $S
