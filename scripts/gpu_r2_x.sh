#!/bin/bash
# round 2, call X (2 GPUs): final N=2 bench sanity (adaptive collate depth under torchrun)
set -u
mkdir -p gpurun_out
for c in flagship tagger_w96; do
  if [ $c = flagship ]; then CF=""; else CF="--config configs/$c.cfg"; fi
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29671 bench.py --gpus 2 $CF --steps 100 --warmup 5 --no-own-baseline > gpurun_out/r2x_$c.json 2> gpurun_out/r2x_$c.err
  python - "$c" <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/r2x_{n}.json").read().strip().splitlines()[-1])
    print(n, round(d["value"]), round(d["ms_per_step"],4), d["step_ms"], "e2e", round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"],4))
except Exception as e:
    print(n, "FAILED", e, open(f"gpurun_out/r2x_{n}.err").read()[-2500:])
PY
done
