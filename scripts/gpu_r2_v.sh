#!/bin/bash
# round 2, call V (1 GPU): two collate workers / ordered prefetch: engine tests + end-to-end numbers
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parser.py tests/test_gpu_kernels.py -q -x -k "engine or lag or training or worker" 2>&1 | tail -3
run() { name=$1; shift
  timeout 300 python bench.py "$@" --steps 200 --warmup 10 --no-own-baseline > gpurun_out/r2v_$name.json 2> gpurun_out/r2v_$name.err
  python - "$name" <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/r2v_{n}.json").read().strip().splitlines()[-1])
    print(n, round(d["value"]), round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"],4), d["gpu_launches"]//d["steps"])
except Exception as e:
    print(n, "FAILED", e, open(f"gpurun_out/r2v_{n}.err").read()[-2500:])
PY
}
run flagship
run tagger_w96 --config configs/tagger_w96.cfg
run multitask_w512 --config configs/multitask_w512.cfg
