#!/bin/bash
# bench (graph + eager engines), launch list and one full ncu capture of the top kernel.
set -u
mkdir -p gpurun_out
: > gpurun_out/summary2.txt
echo "=== biluo test" | tee -a gpurun_out/summary2.txt
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -x -p no:cacheprovider -k "biluo or gpu_training" > gpurun_out/test_biluo2.log 2>&1
echo "exit=$? $(tail -n 1 gpurun_out/test_biluo2.log)" | tee -a gpurun_out/summary2.txt
echo "=== bench graph" | tee -a gpurun_out/summary2.txt
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_graph.log 2>&1; echo "exit=$? $(tail -n 1 gpurun_out/bench_graph.log)" | tee -a gpurun_out/summary2.txt
echo "=== bench eager" | tee -a gpurun_out/summary2.txt
timeout 600 python bench.py --steps 10 --warmup 3 --engine eager > gpurun_out/bench_eager.log 2>&1; echo "exit=$? $(tail -n 1 gpurun_out/bench_eager.log)" | tee -a gpurun_out/summary2.txt
echo "=== bench big batch" | tee -a gpurun_out/summary2.txt
timeout 600 python bench.py --steps 10 --warmup 4 --docs-per-gpu 4096 > gpurun_out/bench_graph4k.log 2>&1; echo "exit=$? $(tail -n 1 gpurun_out/bench_graph4k.log)" | tee -a gpurun_out/summary2.txt
echo "=== ncu launch list" | tee -a gpurun_out/summary2.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 400 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 2 --warmup 3 --engine eager --no-e2e > gpurun_out/ncu_launch.log 2>&1; echo "exit=$?" | tee -a gpurun_out/summary2.txt
echo "=== ncu full (window maxout GEMM)" | tee -a gpurun_out/summary2.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 30 -c 3 -o gpurun_out/prof_gemm -f \
   python bench.py --steps 2 --warmup 3 --engine eager --no-e2e > gpurun_out/ncu_full.log 2>&1; echo "exit=$?" | tee -a gpurun_out/summary2.txt
cat gpurun_out/summary2.txt
