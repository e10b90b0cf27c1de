#!/bin/bash
# round 2, call S (1 GPU): forward window GEMM with the weights resident in shared memory
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x -k "maxout or window or fused_layernorm or training or engine or width_96" 2>&1 | tail -4
run() { name=$1; shift
  env "$@" timeout 300 python bench.py --steps 100 --warmup 10 --no-own-baseline > gpurun_out/r2s_$name.json 2> gpurun_out/r2s_$name.err
  python - "$name" <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/r2s_{n}.json").read().strip().splitlines()[-1])
    print(n, round(d["value"]), round(d["ms_per_step"],4), d["step_ms"], "e2e", round(d["e2e"]["value"]), d["gpu_launches"]//d["steps"])
except Exception as e:
    print(n, "FAILED", e, open(f"gpurun_out/r2s_{n}.err").read()[-2500:])
PY
}
run bres X=1
run nobres SRB_GEMM_BRES=0
SRB_FUSED_LN=0 timeout 200 python benchmarks/layer_bench.py 2>&1 | tail -1
SRB_GEMM_BRES=0 timeout 200 python benchmarks/layer_bench.py 2>&1 | tail -1
