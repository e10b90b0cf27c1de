#!/bin/bash
# tests + B-ray emulation + default bench.  usage: gpu_multi2.sh
set -u
mkdir -p gpurun_out
S=gpurun_out/summary_m2.txt
: > $S
echo "=== tests" | tee -a $S
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -p no:cacheprovider > gpurun_out/test_all.log 2>&1
echo "exit=$? $(tail -n 1 gpurun_out/test_all.log)" | tee -a $S
grep -E "^(FAILED|ERROR)" gpurun_out/test_all.log | tee -a $S
echo "=== bench default (200 steps)" | tee -a $S
timeout 600 python bench.py > gpurun_out/bench_default.log 2>&1; echo "exit=$? $(grep '^{' gpurun_out/bench_default.log | tail -n 1 | cut -c1-300)" | tee -a $S
echo "=== B-ray emulation, 2 GPU workers, quorum 2" | tee -a $S
timeout 600 python benchmarks/bench_rayproxy.py --workers 2 --gpu --quorum 2 --steps 25 > gpurun_out/bray_q2.log 2>&1; echo "exit=$? $(grep '^{' gpurun_out/bray_q2.log | tail -n 1 | cut -c1-700)" | tee -a $S
echo "=== B-ray emulation, 2 GPU workers, quorum N (sync special case)" | tee -a $S
timeout 600 python benchmarks/bench_rayproxy.py --workers 2 --gpu --quorum 4 --steps 25 > gpurun_out/bray_qn.log 2>&1; echo "exit=$? $(grep '^{' gpurun_out/bray_qn.log | tail -n 1 | cut -c1-700)" | tee -a $S
echo "=== actors + fused comm (CLI path), 2 GPU workers" | tee -a $S
timeout 600 python benchmarks/bench_rayproxy.py --workers 2 --gpu --mode sync --comm auto --steps 60 > gpurun_out/actors_fused.log 2>&1; echo "exit=$? $(grep '^{' gpurun_out/actors_fused.log | tail -n 1 | cut -c1-500)" | tee -a $S
tail -3 gpurun_out/bray_q2.log | cut -c1-300
