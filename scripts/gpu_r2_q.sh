#!/bin/bash
# round 2, call Q (1 GPU): LayerNorm vec kernels for every width that is a multiple of 8: numerics + benches
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -12
run() { name=$1; shift
  timeout 300 python bench.py "$@" --steps 100 --warmup 10 --no-own-baseline > gpurun_out/r2q_$name.json 2> gpurun_out/r2q_$name.err
  python - "$name" <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/r2q_{n}.json").read().strip().splitlines()[-1])
    print(n, round(d["value"]), round(d["ms_per_step"],4), d["step_ms"], "e2e", round(d["e2e"]["value"]), d["gpu_launches"]//d["steps"])
except Exception as e:
    print(n, "FAILED", e, open(f"gpurun_out/r2q_{n}.err").read()[-2500:])
PY
}
run flagship
run tagger_w96 --config configs/tagger_w96.cfg
run parser_w256 --config configs/parser_w256.cfg
run multitask_w512 --config configs/multitask_w512.cfg


