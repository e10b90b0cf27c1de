#!/bin/bash
# round 2, call F (2 GPUs): new single-GPU tests, exchange timeline at 1 and 2 GPUs, 2-GPU variants
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m "gpu and not multigpu" -q 2>&1 | tail -40 > gpurun_out/r2f_test_all.log
tail -15 gpurun_out/r2f_test_all.log
timeout 300 python benchmarks/exchange_trace.py --steps 30 --out gpurun_out/r2f_trace_1gpu > gpurun_out/r2f_trace_1gpu.txt 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 benchmarks/exchange_trace.py --steps 30 --out gpurun_out/r2f_trace_2gpu > gpurun_out/r2f_trace_2gpu.txt 2>&1
SRB_NVLS=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 benchmarks/exchange_trace.py --steps 30 --out gpurun_out/r2f_trace_2gpu_p2p > gpurun_out/r2f_trace_2gpu_p2p.txt 2>&1
grep -v "^W0\|^\*\*\*\|OMP_NUM" gpurun_out/r2f_trace_2gpu.txt | tail -75
run() { name=$1; n=$2; shift 2
  if [ $n = 1 ]; then L="python"; else L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n"; fi
  env "$@" timeout 300 $L bench.py --gpus $n --steps 100 --warmup 10 --no-own-baseline --no-e2e > gpurun_out/r2f_$name.json 2> gpurun_out/r2f_$name.err
  python - "$name" <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/r2f_{n}.json").read().strip().splitlines()[-1])
    print(n, round(d["value"]), round(d["ms_per_step"],4), d["step_ms"], d["gpu_launches"], d["config"]["exchange_buckets"])
except Exception as e:
    print(n, "FAILED", e, open(f"gpurun_out/r2f_{n}.err").read()[-2000:])
PY
}
run n1 1 X=1
run n2 2 X=1
run n2_p2p 2 SRB_NVLS=0
run n2_b6 2 SRB_COMM_BUCKETS=6
