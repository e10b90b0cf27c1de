#!/bin/bash
# Run on a GPU box via gpurun.  Each test group runs in its own process under `timeout`, so a
# hung kernel cannot eat the whole call.  Logs land in gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/nvsmi.txt 2>&1
python -c "import torch; print(torch.__version__, torch.cuda.get_device_name(0))" > gpurun_out/env.txt 2>&1
run() {  # name, timeout, pytest -k expr
  echo "=== $1" | tee -a gpurun_out/summary.txt
  timeout "$2" python -m pytest tests/test_gpu_kernels.py -q -x -p no:cacheprovider -k "$3" > "gpurun_out/test_$1.log" 2>&1
  echo "exit=$? $(tail -n 1 gpurun_out/test_$1.log)" | tee -a gpurun_out/summary.txt
}
: > gpurun_out/summary.txt
run elementwise 300 "hash_embed or softmax_xent or adam_shard"
run gemm_plain 200 "tc_gemm_plain"
run gemm_window 200 "tc_window_maxout"
run gemm_dx 200 "tc_window_dx"
run gemm_dw 200 "tc_dw"
run block_lib 300 "maxout_block and False"
run block_tc 300 "maxout_block and True"
run biluo 300 "biluo"
run e2e 400 "gpu_training"
echo "=== smoke" | tee -a gpurun_out/summary.txt
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "exit=$? $(tail -n 1 gpurun_out/smoke.log)" | tee -a gpurun_out/summary.txt
echo "=== bench" | tee -a gpurun_out/summary.txt
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1; echo "exit=$? $(tail -n 1 gpurun_out/bench.log)" | tee -a gpurun_out/summary.txt
SRB_USE_TC=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e > gpurun_out/bench_notc.log 2>&1; echo "notc exit=$? $(tail -n 1 gpurun_out/bench_notc.log)" | tee -a gpurun_out/summary.txt
cat gpurun_out/summary.txt
