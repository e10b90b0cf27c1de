#!/bin/bash
# round 2, call K (1 GPU): programmatic dependent launch on the whole training chain
set -u
mkdir -p gpurun_out
run() { name=$1; shift
  env "$@" timeout 400 python bench.py --steps 100 --warmup 10 --no-own-baseline > gpurun_out/r2k_$name.json 2> gpurun_out/r2k_$name.err
  python - "$name" <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/r2k_{n}.json").read().strip().splitlines()[-1])
    print(n, round(d["value"]), round(d["ms_per_step"],4), d["step_ms"], "e2e", round(d["e2e"]["value"]), d["gpu_launches"]//d["steps"])
except Exception as e:
    print(n, "FAILED", e, open(f"gpurun_out/r2k_{n}.err").read()[-2500:])
PY
}
run pdl X=1
run nopdl SRB_PDL=0
run pdl_side SRB_PDL_SIDE=1
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r2k_test.log
tail -8 gpurun_out/r2k_test.log
for c in tagger_w96 parser_w256 multitask_w512; do
  timeout 400 python bench.py --config $c --steps 60 --warmup 10 --no-own-baseline > gpurun_out/r2k_$c.json 2> gpurun_out/r2k_$c.err
  python - "$c" <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/r2k_{n}.json").read().strip().splitlines()[-1])
    print(n, round(d["value"]), round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]))
except Exception as e:
    print(n, "FAILED", e, open(f"gpurun_out/r2k_{n}.err").read()[-1500:])
PY
done
