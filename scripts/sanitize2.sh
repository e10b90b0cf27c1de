#!/bin/bash
# compute-sanitizer over the kernel tests that changed this round (bounded: each pass is killed after its budget)
set -u
mkdir -p gpurun_out
K1="colsum or hash_embed or softmax_xent or biluo_kernel or maxout_block"
K2="arc_eager_kernel"
K3="test_tc_gemm_plain_nt and 128-128-64"
for tool in memcheck racecheck; do
  timeout -s KILL 200 compute-sanitizer --tool $tool --error-exitcode 1 python -m pytest tests/test_gpu_kernels.py -q -x -p no:cacheprovider -k "$K1" > gpurun_out/sanitize_${tool}_elementwise.log 2>&1
  echo "$tool elementwise+biluo exit=$? $(grep -E 'ERROR SUMMARY|passed|failed' gpurun_out/sanitize_${tool}_elementwise.log | tail -2 | tr '\n' ' ')"
  timeout -s KILL 200 compute-sanitizer --tool $tool --error-exitcode 1 python -m pytest tests/test_gpu_parser.py -q -x -p no:cacheprovider -k "$K2" > gpurun_out/sanitize_${tool}_arc.log 2>&1
  echo "$tool arc exit=$? $(grep -E 'ERROR SUMMARY|passed|failed' gpurun_out/sanitize_${tool}_arc.log | tail -2 | tr '\n' ' ')"
done
timeout -s KILL 150 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_kernels.py -q -x -p no:cacheprovider -k "$K3" > gpurun_out/sanitize_memcheck_gemm.log 2>&1
echo "memcheck gemm(1,2,3) exit=$? $(grep -E 'ERROR SUMMARY|passed|failed' gpurun_out/sanitize_memcheck_gemm.log | tail -2 | tr '\n' ' ')"
