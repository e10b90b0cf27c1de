#!/bin/bash
# round 2, call Y (1 GPU): GEMM roofline micro-benchmark with the final kernels + full GPU suite on HEAD
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -2
timeout 300 python benchmarks/gemm_bench.py --json gpurun_out/r2y_gemm_bench.json 2>&1 | tail -30
