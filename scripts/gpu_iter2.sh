#!/bin/bash
# tests + flagship bench + per-config benches (parser / multitask / tagger) through the engine
set -u
mkdir -p gpurun_out
S=gpurun_out/summary4.txt
: > $S
echo "=== tests" | tee -a $S
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/test_all.log 2>&1
echo "exit=$? $(tail -n 1 gpurun_out/test_all.log)" | tee -a $S
grep -E "^(FAILED|ERROR)" gpurun_out/test_all.log | tee -a $S
echo "=== bench flagship" | tee -a $S
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_flagship.log 2>&1; echo "exit=$? $(tail -n 1 gpurun_out/bench_flagship.log | cut -c1-600)" | tee -a $S
for c in tagger_w96 parser_w256 multitask_w512; do
  echo "=== bench $c" | tee -a $S
  timeout 600 python bench.py --steps 30 --warmup 5 --config configs/$c.cfg > gpurun_out/bench_$c.log 2>&1
  echo "exit=$? $(tail -n 1 gpurun_out/bench_$c.log | cut -c1-700)" | tee -a $S
done
if [ "${1:-}" = "ncu" ]; then
  echo "=== ncu launch list (multitask, eager)" | tee -a $S
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 900 -c 500 --csv --log-file gpurun_out/launches_multitask.csv \
     python bench.py --steps 2 --warmup 3 --engine eager --no-e2e --config configs/multitask_w512.cfg > gpurun_out/ncu_launch_mt.log 2>&1; echo "exit=$?" | tee -a $S
fi
cat $S
