#!/bin/bash
# round 2, call B (1 GPU): exchange-overlap variants at 1 GPU + launch list of the graph step
set -u
mkdir -p gpurun_out
B="python bench.py --steps 100 --warmup 10 --no-e2e --no-own-baseline"
run() { name=$1; shift; env "$@" timeout 300 $B > gpurun_out/r2b_$name.json 2> gpurun_out/r2b_$name.err; python - "$name" <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/r2b_{n}.json").read().strip().splitlines()[-1])
    print(n, round(d["value"]), round(d["ms_per_step"],4), d["step_ms"], d["gpu_launches"], d["config"]["exchange_buckets"])
except Exception as e:
    print(n, "FAILED", e, open(f"gpurun_out/r2b_{n}.err").read()[-800:])
PY
}
run default X=1
run prio0 SRB_COMM_PRIO=0
run nooverlap SRB_COMM_OVERLAP=0
run buckets1 SRB_COMM_BUCKETS=1
run buckets10 SRB_COMM_BUCKETS=10
run buckets3 SRB_COMM_BUCKETS=3
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 200 --csv --log-file gpurun_out/r2b_launches.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-own-baseline > gpurun_out/r2b_ncu.log 2>&1
python scripts/launch_summary.py gpurun_out/r2b_launches.csv > gpurun_out/r2b_launch_summary.txt 2>&1 | tail -3
