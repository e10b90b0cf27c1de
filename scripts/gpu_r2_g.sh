#!/bin/bash
# round 2, call G (2 GPUs): re-validate after the signal / local-clear rework, trace + bench
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_comm.py tests/test_gpu_kernels.py -q -x 2>&1 | tail -5 > gpurun_out/r2g_test_1gpu.log
tail -3 gpurun_out/r2g_test_1gpu.log
timeout 900 python -m pytest tests/test_multigpu.py -x -q -k "not other_optimizers" 2>&1 | tail -30 > gpurun_out/r2g_test_multigpu.log
tail -6 gpurun_out/r2g_test_multigpu.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 benchmarks/exchange_trace.py --steps 30 --out gpurun_out/r2g_trace_2gpu > gpurun_out/r2g_trace_2gpu.txt 2>&1
grep -v "^W0\|^\*\*\*\|OMP_NUM\|grad[0-9]" gpurun_out/r2g_trace_2gpu.txt | tail -32
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2g_trace_2gpu_rank1.json")); m=d["median_us"]
print("rank1", d["owned_per_bucket"])
for k,v in m.items():
    if not k.startswith("grad"): print(f"{v:9.1f} {k}")
PY
run() { name=$1; n=$2; shift 2
  if [ $n = 1 ]; then L="python"; else L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n"; fi
  env "$@" timeout 300 $L bench.py --gpus $n --steps 100 --warmup 10 --no-own-baseline > gpurun_out/r2g_$name.json 2> gpurun_out/r2g_$name.err
  python - "$name" <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/r2g_{n}.json").read().strip().splitlines()[-1])
    print(n, round(d["value"]), round(d["ms_per_step"],4), d["step_ms"], "e2e", round(d["e2e"]["value"]), d["gpu_launches"], d["config"]["exchange_buckets"])
except Exception as e:
    print(n, "FAILED", e, open(f"gpurun_out/r2g_{n}.err").read()[-2000:])
PY
}
run n1 1 X=1
run n2 2 X=1
run n2_p2p 2 SRB_NVLS=0
