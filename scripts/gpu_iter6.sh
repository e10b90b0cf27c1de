#!/bin/bash
# pair-MMA (cta_group::2) bring-up: its tests first under a hard kill timeout, then the rest
set -u
mkdir -p gpurun_out
S=gpurun_out/summary8.txt
: > $S
echo "=== pair tests" | tee -a $S
timeout -s KILL 240 python -m pytest tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider -k "tc_ and 3" -x > gpurun_out/test_pair.log 2>&1
rc=$?
echo "exit=$rc $(tail -n 1 gpurun_out/test_pair.log)" | tee -a $S
grep -E "^(FAILED|ERROR)" gpurun_out/test_pair.log | head -20 | tee -a $S
if [ $rc -eq 0 ]; then
  echo "=== gemm bench" | tee -a $S
  timeout -s KILL 300 python benchmarks/gemm_bench.py --json gpurun_out/gemm_bench.json 2>&1 | tee -a $S
  for cl in 2 3; do
    echo "=== bench flagship SRB_GEMM_CLUSTER=$cl" | tee -a $S
    SRB_GEMM_CLUSTER=$cl timeout -s KILL 400 python bench.py --steps 100 --warmup 10 > gpurun_out/bench_flagship_cl$cl.log 2>&1; echo "exit=$? $(tail -n 1 gpurun_out/bench_flagship_cl$cl.log | cut -c1-300)" | tee -a $S
  done
fi
echo "=== all tests" | tee -a $S
timeout -s KILL 900 python -m pytest tests -m gpu -q -p no:cacheprovider -k "not (tc_ and 3)" > gpurun_out/test_all.log 2>&1
echo "exit=$? $(tail -n 1 gpurun_out/test_all.log)" | tee -a $S
grep -E "^(FAILED|ERROR)" gpurun_out/test_all.log | tee -a $S
