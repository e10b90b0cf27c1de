#!/bin/bash
# round 2, call C (1 GPU): tests of the reworked exchange + head glue, bench variants, launch list
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_comm.py -x -q 2>&1 | tail -30 > gpurun_out/r2c_test_comm.log
timeout 1500 python -m pytest tests -m "gpu and not multigpu" -q -x --deselect tests/test_gpu_comm.py 2>&1 | tail -40 > gpurun_out/r2c_test_all.log
tail -4 gpurun_out/r2c_test_comm.log gpurun_out/r2c_test_all.log
B="python bench.py --steps 100 --warmup 10 --no-e2e --no-own-baseline"
run() { name=$1; shift; env "$@" timeout 300 $B > gpurun_out/r2c_$name.json 2> gpurun_out/r2c_$name.err; python - "$name" <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/r2c_{n}.json").read().strip().splitlines()[-1])
    print(n, round(d["value"]), round(d["ms_per_step"],4), d["step_ms"], d["gpu_launches"], d["config"]["exchange_buckets"])
except Exception as e:
    print(n, "FAILED", e, open(f"gpurun_out/r2c_{n}.err").read()[-1500:])
PY
}
run default X=1
run nooverlap SRB_COMM_OVERLAP=0
run buckets1 SRB_COMM_BUCKETS=1
run buckets3 SRB_COMM_BUCKETS=3
run buckets10 SRB_COMM_BUCKETS=10
run prio0 SRB_COMM_PRIO=0
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 200 --csv --log-file gpurun_out/r2c_launches.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-own-baseline > gpurun_out/r2c_ncu.log 2>&1
python scripts/launch_summary.py gpurun_out/r2c_launches.csv > gpurun_out/r2c_launch_summary.txt 2>&1
head -30 gpurun_out/r2c_launch_summary.txt
