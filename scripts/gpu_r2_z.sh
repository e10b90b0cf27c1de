#!/bin/bash
# round 2, call Z (4 GPUs): N=4 flagship point with the final code
set -u
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29681 bench.py --gpus 4 --steps 100 --warmup 5 --no-own-baseline > gpurun_out/r2z_n4.json 2> gpurun_out/r2z_n4.err
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r2z_n4.json").read().strip().splitlines()[-1])
    print("n4", round(d["value"]), round(d["ms_per_step"],4), d["step_ms"], "e2e", round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"],4))
except Exception as e:
    print("n4 FAILED", e, open("gpurun_out/r2z_n4.err").read()[-2000:])
PY
