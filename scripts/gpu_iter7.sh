#!/bin/bash
set -u
mkdir -p gpurun_out
S=gpurun_out/summary9.txt
: > $S
echo "=== tests" | tee -a $S
timeout -s KILL 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/test_all.log 2>&1
echo "exit=$? $(tail -n 1 gpurun_out/test_all.log)" | tee -a $S
grep -E "^(FAILED|ERROR)" gpurun_out/test_all.log | tee -a $S
for sd in 0 1; do
  echo "=== bench flagship SRB_SIDE_DW=$sd" | tee -a $S
  SRB_SIDE_DW=$sd timeout -s KILL 400 python bench.py --steps 100 --warmup 10 > gpurun_out/bench_flagship_side$sd.log 2>&1; echo "exit=$? $(tail -n 1 gpurun_out/bench_flagship_side$sd.log | cut -c1-300)" | tee -a $S
done
for c in parser_w256 multitask_w512 tagger_w96; do
  echo "=== bench $c" | tee -a $S
  timeout -s KILL 400 python bench.py --steps 30 --warmup 5 --config configs/$c.cfg > gpurun_out/bench_$c.log 2>&1
  echo "exit=$? $(tail -n 1 gpurun_out/bench_$c.log | cut -c1-300)" | tee -a $S
done
