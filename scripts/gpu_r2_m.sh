#!/bin/bash
# round 2, call M (8 GPUs): scaling points N=1,2,4,8 on ONE box with the fixed bench timing,
# plus the other configs at N=8 and the 8-rank correctness tests
set -u
mkdir -p gpurun_out
run() { name=$1; n=$2; shift 2
  if [ $n = 1 ]; then L="python"; else L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2962$n"; fi
  timeout 300 $L bench.py --gpus $n --steps 100 --warmup 10 "$@" > gpurun_out/r2m_$name.json 2> gpurun_out/r2m_$name.err
  python - "$name" <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/r2m_{n}.json").read().strip().splitlines()[-1])
    print(n, round(d["value"]), round(d["ms_per_step"],4), d["step_ms"], "e2e", round(d["e2e"]["value"]), d["gpu_launches"], d["config"].get("exchange_buckets"), d.get("vs_own_nccl_baseline"), d["clocks"])
except Exception as e:
    print(n, "FAILED", e, open(f"gpurun_out/r2m_{n}.err").read()[-2000:])
PY
}
run n1 1 --no-own-baseline
run n2 2 --no-own-baseline
run n4 4 --no-own-baseline
run n8 8
run tagger_n8 8 --config configs/tagger_w96.cfg --no-own-baseline
run parser_n8 8 --config configs/parser_w256.cfg --no-own-baseline
run multitask_n8 8 --config configs/multitask_w512.cfg --no-own-baseline
timeout 600 python -m pytest tests/test_multigpu.py -x -q -k "8" 2>&1 | tail -8 > gpurun_out/r2m_test_multigpu8.log
tail -4 gpurun_out/r2m_test_multigpu8.log
