#!/bin/bash
# compute-sanitizer passes over the kernel unit tests (run on a GPU box; slow: minutes per tool).
# memcheck = OOB / misaligned, racecheck = shared-memory hazards, synccheck = barrier misuse.
# Cross-GPU flag protocols are outside what the sanitizer understands; those are covered by
# benchmarks/comm_check.py (value-checks against NCCL) and the kernels' spin-wait timeouts.
set -u
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 1 \
     python -m pytest tests/test_gpu_kernels.py -q -x -p no:cacheprovider -k "${1:-hash_embed or softmax_xent or adam_shard or tc_gemm_plain}" \
     > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool exit=$? $(grep -E 'ERROR SUMMARY|passed|failed' gpurun_out/sanitize_$tool.log | tail -2 | tr '\n' ' ')"
done
