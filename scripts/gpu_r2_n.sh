#!/bin/bash
# round 2, call N (1 GPU): fused GEMM + LayerNorm epilogue: numerics, then the flagship step with / without
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -x -k "fused_layernorm or maxout_block" 2>&1 | tail -15
run() { name=$1; shift
  env "$@" timeout 300 python bench.py --steps 100 --warmup 10 --no-own-baseline > gpurun_out/r2n_$name.json 2> gpurun_out/r2n_$name.err
  python - "$name" <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/r2n_{n}.json").read().strip().splitlines()[-1])
    print(n, round(d["value"]), round(d["ms_per_step"],4), d["step_ms"], "e2e", round(d["e2e"]["value"]), d["gpu_launches"]//d["steps"])
except Exception as e:
    print(n, "FAILED", e, open(f"gpurun_out/r2n_{n}.err").read()[-2500:])
PY
}
run fused X=1
run plain SRB_FUSED_LN=0
timeout 200 python benchmarks/layer_bench.py 2>&1 | tail -2
