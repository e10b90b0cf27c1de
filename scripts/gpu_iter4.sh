#!/bin/bash
# tests + flagship bench + ncu full captures of the two transition kernels
set -u
mkdir -p gpurun_out
S=gpurun_out/summary6.txt
: > $S
echo "=== tests" | tee -a $S
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/test_all.log 2>&1
echo "exit=$? $(tail -n 1 gpurun_out/test_all.log)" | tee -a $S
grep -E "^(FAILED|ERROR)" gpurun_out/test_all.log | tee -a $S
echo "=== bench flagship" | tee -a $S
timeout 600 python bench.py --steps 100 --warmup 10 > gpurun_out/bench_flagship.log 2>&1; echo "exit=$? $(tail -n 1 gpurun_out/bench_flagship.log | cut -c1-300)" | tee -a $S
echo "=== bench parser" | tee -a $S
timeout 600 python bench.py --steps 30 --warmup 5 --config configs/parser_w256.cfg > gpurun_out/bench_parser_w256.log 2>&1; echo "exit=$? $(tail -n 1 gpurun_out/bench_parser_w256.log | cut -c1-300)" | tee -a $S
if [ "${1:-}" = "ncu" ]; then
  echo "=== ncu full: biluo" | tee -a $S
  timeout 900 ncu --set full --import-source on --clock-control none -k regex:biluo_steps_kernel -s 3 -c 1 -f -o gpurun_out/biluo_full \
     python bench.py --steps 2 --warmup 3 --engine eager --no-e2e > gpurun_out/ncu_biluo.log 2>&1; echo "exit=$?" | tee -a $S
  echo "=== ncu full: arc" | tee -a $S
  timeout 900 ncu --set full --import-source on --clock-control none -k regex:arc_eager_steps_kernel -s 3 -c 1 -f -o gpurun_out/arc_full \
     python bench.py --steps 2 --warmup 3 --engine eager --no-e2e --config configs/parser_w256.cfg > gpurun_out/ncu_arc.log 2>&1; echo "exit=$?" | tee -a $S
  ls -la gpurun_out/*.ncu-rep | tee -a $S
fi
cat $S
