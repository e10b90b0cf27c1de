#!/bin/bash
# round 2, call O (1 GPU): ncu --set full of the fused GEMM+LayerNorm kernel (layer_bench, one capture)
set -u
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 12 -c 1 -f -o gpurun_out/r2o_fused_ln \
  python benchmarks/layer_bench.py --iters 4 > gpurun_out/r2o_ncu.log 2>&1
tail -3 gpurun_out/r2o_ncu.log
ls -la gpurun_out/r2o_fused_ln.ncu-rep
