#!/bin/bash
# quick iteration: all GPU tests + bench + launch list
set -u
mkdir -p gpurun_out
: > gpurun_out/summary3.txt
echo "=== tests" | tee -a gpurun_out/summary3.txt
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/test_all.log 2>&1
echo "exit=$? $(tail -n 1 gpurun_out/test_all.log)" | tee -a gpurun_out/summary3.txt
grep -E "^(FAILED|ERROR)" gpurun_out/test_all.log | tee -a gpurun_out/summary3.txt
echo "=== bench graph" | tee -a gpurun_out/summary3.txt
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_graph.log 2>&1; echo "exit=$? $(tail -n 1 gpurun_out/bench_graph.log | cut -c1-400)" | tee -a gpurun_out/summary3.txt
if [ "${1:-}" != "nolist" ]; then
echo "=== ncu launch list" | tee -a gpurun_out/summary3.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 500 -c 300 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 2 --warmup 3 --engine eager --no-e2e > gpurun_out/ncu_launch.log 2>&1; echo "exit=$?" | tee -a gpurun_out/summary3.txt
fi
cat gpurun_out/summary3.txt
