#!/bin/bash
# round 2, call A (1 GPU): new comm kernel tests first, then the whole GPU suite, then the bench.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_env.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_comm.py -x -q 2>&1 | tail -40 > gpurun_out/r2a_test_comm.log
echo "comm tests rc=$?" >> gpurun_out/r2a_test_comm.log
timeout 1500 python -m pytest tests -m "gpu and not multigpu" -q -x --deselect tests/test_gpu_comm.py 2>&1 | tail -40 > gpurun_out/r2a_test_all.log
timeout 600 python bench.py --steps 100 --warmup 10 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
timeout 600 python bench.py --steps 100 --warmup 10 --random-batches --bucket-rows 1024 --no-own-baseline > gpurun_out/r2a_bench_random1024.json 2>> gpurun_out/r2a_bench.err
tail -5 gpurun_out/r2a_test_comm.log gpurun_out/r2a_test_all.log
cat gpurun_out/r2a_bench.json
tail -5 gpurun_out/r2a_bench.err
