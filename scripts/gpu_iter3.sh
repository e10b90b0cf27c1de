#!/bin/bash
# tests + flagship bench + launch lists (flagship, parser, multitask; eager engine so kernels are named)
set -u
mkdir -p gpurun_out
S=gpurun_out/summary5.txt
: > $S
echo "=== tests" | tee -a $S
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/test_all.log 2>&1
echo "exit=$? $(tail -n 1 gpurun_out/test_all.log)" | tee -a $S
grep -E "^(FAILED|ERROR)" gpurun_out/test_all.log | tee -a $S
echo "=== bench flagship" | tee -a $S
timeout 600 python bench.py --steps 100 --warmup 10 > gpurun_out/bench_flagship.log 2>&1; echo "exit=$? $(tail -n 1 gpurun_out/bench_flagship.log | cut -c1-300)" | tee -a $S
for c in ${BENCH_CONFIGS:-}; do
  echo "=== bench $c" | tee -a $S
  timeout 600 python bench.py --steps 30 --warmup 5 --config configs/$c.cfg > gpurun_out/bench_$c.log 2>&1
  echo "exit=$? $(tail -n 1 gpurun_out/bench_$c.log | cut -c1-300)" | tee -a $S
done
if [ "${1:-}" = "ncu" ]; then
  echo "=== ncu launch list flagship" | tee -a $S
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 500 -c 300 --csv --log-file gpurun_out/launches.csv \
     python bench.py --steps 2 --warmup 3 --engine eager --no-e2e > gpurun_out/ncu_launch.log 2>&1; echo "exit=$?" | tee -a $S
  for c in parser_w256 multitask_w512; do
    echo "=== ncu launch list $c" | tee -a $S
    timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 450 -c 450 --csv --log-file gpurun_out/launches_$c.csv \
       python bench.py --steps 2 --warmup 3 --engine eager --no-e2e --config configs/$c.cfg > gpurun_out/ncu_launch_$c.log 2>&1; echo "exit=$?" | tee -a $S
  done
fi
cat $S
