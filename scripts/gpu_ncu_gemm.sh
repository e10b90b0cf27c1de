#!/bin/bash
# ncu --set full of the GEMM family inside one eager flagship step
set -u
mkdir -p gpurun_out
timeout -s KILL 900 ncu --set full --import-source on --clock-control none -k regex:gemm_kernel -s 24 -c 12 -f -o gpurun_out/gemm_full \
   python bench.py --steps 2 --warmup 3 --engine eager --no-e2e > gpurun_out/ncu_gemm.log 2>&1
echo "exit=$?"; ls -la gpurun_out/gemm_full.ncu-rep
