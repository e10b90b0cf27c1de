#!/bin/bash
# round 2, call H (8 GPUs): 8-rank correctness, step timeline, scaling points
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multigpu.py -x -q -k "8" 2>&1 | tail -30 > gpurun_out/r2h_test_multigpu8.log
tail -6 gpurun_out/r2h_test_multigpu8.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 benchmarks/exchange_trace.py --steps 30 --out gpurun_out/r2h_trace_8gpu > gpurun_out/r2h_trace_8gpu.txt 2>&1
grep -v "^W0\|^\*\*\*\|OMP_NUM\|grad[0-9]" gpurun_out/r2h_trace_8gpu.txt | tail -36
python - <<'PY'
import json
for r in range(8):
    try:
        d=json.load(open(f"gpurun_out/r2h_trace_8gpu_rank{r}.json")); m=d["median_us"]
        print("rank",r, d["owned_per_bucket"], {k:round(v) for k,v in m.items() if k.startswith("b3.") or k in ("step_end","backward_enqueued_done")})
    except Exception as e: print(r, e)
PY
run() { name=$1; n=$2; shift 2
  if [ $n = 1 ]; then L="python"; else L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n"; fi
  env "$@" timeout 300 $L bench.py --gpus $n --steps 100 --warmup 10 --no-own-baseline > gpurun_out/r2h_$name.json 2> gpurun_out/r2h_$name.err
  python - "$name" <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/r2h_{n}.json").read().strip().splitlines()[-1])
    print(n, round(d["value"]), round(d["ms_per_step"],4), d["step_ms"], "e2e", round(d["e2e"]["value"]), d["gpu_launches"], d["config"]["exchange_buckets"])
except Exception as e:
    print(n, "FAILED", e, open(f"gpurun_out/r2h_{n}.err").read()[-2000:])
PY
}
run n1 1 X=1
run n8 8 X=1
run n8_p2p 8 SRB_NVLS=0
run n4 4 X=1
