#!/bin/bash
# round 2, call U (2 GPUs): multi-GPU tests and an N=2 bench with the final kernels
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multigpu.py -x -q 2>&1 | tail -5
for c in flagship tagger_w96; do
  if [ $c = flagship ]; then CF=""; else CF="--config configs/$c.cfg"; fi
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29661 bench.py --gpus 2 $CF --steps 100 --warmup 10 --no-own-baseline > gpurun_out/r2u_$c.json 2> gpurun_out/r2u_$c.err
  python - "$c" <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/r2u_{n}.json").read().strip().splitlines()[-1])
    print(n, round(d["value"]), round(d["ms_per_step"],4), d["step_ms"], "e2e", round(d["e2e"]["value"]))
except Exception as e:
    print(n, "FAILED", e, open(f"gpurun_out/r2u_{n}.err").read()[-2500:])
PY
done
