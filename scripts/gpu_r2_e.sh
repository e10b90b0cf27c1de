#!/bin/bash
# round 2, call E (2 GPUs): re-validate the exchange after the signal/wait split + 2-GPU variants
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multigpu.py -x -q -k "not other_optimizers and not tagger" 2>&1 | tail -30 > gpurun_out/r2e_test_multigpu.log
tail -6 gpurun_out/r2e_test_multigpu.log
run() { name=$1; n=$2; shift 2
  if [ $n = 1 ]; then L="python"; else L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n"; fi
  env "$@" timeout 300 $L bench.py --gpus $n --steps 100 --warmup 10 --no-own-baseline --no-e2e > gpurun_out/r2e_$name.json 2> gpurun_out/r2e_$name.err
  python - "$name" <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/r2e_{n}.json").read().strip().splitlines()[-1])
    print(n, round(d["value"]), round(d["ms_per_step"],4), d["step_ms"], d["gpu_launches"], d["config"]["exchange_buckets"])
except Exception as e:
    print(n, "FAILED", e, open(f"gpurun_out/r2e_{n}.err").read()[-2000:])
PY
}
run n1 1 X=1
run n2 2 X=1
run n2_prio0 2 SRB_COMM_PRIO=0
run n2_b2 2 SRB_COMM_BUCKETS=2
run n2_b8 2 SRB_COMM_BUCKETS=8
run n2_terminal 2 SRB_GATE_ALWAYS=1
run n2_nooverlap 2 SRB_COMM_OVERLAP=0
run n2_p2p 2 SRB_NVLS=0
B2="--random-batches"; true
