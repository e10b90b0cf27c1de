#!/bin/bash
# tests + flagship/parser bench + GEMM roofline micro-benchmark
set -u
mkdir -p gpurun_out
S=gpurun_out/summary7.txt
: > $S
echo "=== tests" | tee -a $S
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/test_all.log 2>&1
echo "exit=$? $(tail -n 1 gpurun_out/test_all.log)" | tee -a $S
grep -E "^(FAILED|ERROR)" gpurun_out/test_all.log | tee -a $S
echo "=== bench flagship" | tee -a $S
timeout 600 python bench.py --steps 100 --warmup 10 > gpurun_out/bench_flagship.log 2>&1; echo "exit=$? $(tail -n 1 gpurun_out/bench_flagship.log | cut -c1-300)" | tee -a $S
echo "=== bench parser" | tee -a $S
timeout 600 python bench.py --steps 30 --warmup 5 --config configs/parser_w256.cfg > gpurun_out/bench_parser_w256.log 2>&1; echo "exit=$? $(tail -n 1 gpurun_out/bench_parser_w256.log | cut -c1-300)" | tee -a $S
echo "=== gemm bench" | tee -a $S
timeout 600 python benchmarks/gemm_bench.py --json gpurun_out/gemm_bench.json 2>&1 | tee -a $S
cat $S > /dev/null
