#!/bin/bash
# round 2, call D (2 GPUs): multi-GPU correctness of the bucketed exchange + gates, 2-GPU bench
set -u
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2d_topo.txt 2>&1
timeout 1500 python -m pytest tests/test_multigpu.py -x -q 2>&1 | tail -60 > gpurun_out/r2d_test_multigpu.log
tail -25 gpurun_out/r2d_test_multigpu.log
for n in 1 2; do
  if [ $n = 1 ]; then L="python"; else L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2950$n"; fi
  timeout 600 $L bench.py --gpus $n --steps 100 --warmup 10 --no-own-baseline > gpurun_out/r2d_bench_$n.json 2> gpurun_out/r2d_bench_$n.err
  python - $n <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/r2d_bench_{n}.json").read().strip().splitlines()[-1])
    print("N=",n, round(d["value"]), round(d["ms_per_step"],4), d["step_ms"], "e2e", round(d["e2e"]["value"]), d["gpu_launches"], d["config"]["exchange_buckets"])
except Exception as e:
    print(n, "FAILED", e, open(f"gpurun_out/r2d_bench_{n}.err").read()[-2500:])
PY
done
