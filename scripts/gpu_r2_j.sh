#!/bin/bash
# round 2, call J (1 GPU): block-per-doc BILUO kernel: tests, bench, launch list
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x -k "biluo or training or worker or engine" 2>&1 | tail -15 > gpurun_out/r2j_test.log
tail -8 gpurun_out/r2j_test.log
run() { name=$1; shift
  env "$@" timeout 400 python bench.py --steps 100 --warmup 10 --no-own-baseline > gpurun_out/r2j_$name.json 2> gpurun_out/r2j_$name.err
  python - "$name" <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/r2j_{n}.json").read().strip().splitlines()[-1])
    print(n, round(d["value"]), round(d["ms_per_step"],4), d["step_ms"], "e2e", round(d["e2e"]["value"]), d["gpu_launches"]//d["steps"])
except Exception as e:
    print(n, "FAILED", e, open(f"gpurun_out/r2j_{n}.err").read()[-2500:])
PY
}
run block X=1
run warp SRB_BILUO_BLOCK=0
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 160 --csv --log-file gpurun_out/r2j_launches.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-own-baseline > gpurun_out/r2j_ncu.log 2>&1
python scripts/launch_summary.py gpurun_out/r2j_launches.csv > gpurun_out/r2j_launch_summary.txt 2>&1
head -24 gpurun_out/r2j_launch_summary.txt
