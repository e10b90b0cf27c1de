#!/bin/bash
# round 2, call I (1 GPU): full single-GPU suite (new: width 96, accumulate/stream engine), all configs
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m "gpu and not multigpu" -q 2>&1 | tail -40 > gpurun_out/r2i_test_all.log
tail -25 gpurun_out/r2i_test_all.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
run() { name=$1; shift
  timeout 400 python bench.py --steps 60 --warmup 8 --no-own-baseline "$@" > gpurun_out/r2i_$name.json 2> gpurun_out/r2i_$name.err
  python - "$name" <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/r2i_{n}.json").read().strip().splitlines()[-1])
    print(n, round(d["value"]), round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]), d["gpu_launches"]//d["steps"], d["config"]["exchange_buckets"])
except Exception as e:
    print(n, "FAILED", e, open(f"gpurun_out/r2i_{n}.err").read()[-2500:])
PY
}
run flagship
run tagger_w96 --config configs/tagger_w96.cfg
run parser_w256 --config configs/parser_w256.cfg
run multitask_w512 --config configs/multitask_w512.cfg
run ner_w256 --config configs/ner_w256.cfg
