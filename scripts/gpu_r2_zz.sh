#!/bin/bash
# round 2, last call (1 GPU): kernel list of one flagship step with the final code
set -u
mkdir -p gpurun_out
timeout 80 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 240 --csv --log-file gpurun_out/r2zz_launches.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-own-baseline > gpurun_out/r2zz_ncu.log 2>&1
python scripts/launch_summary.py gpurun_out/r2zz_launches.csv > gpurun_out/r2zz_launch_summary.txt 2>&1
cat gpurun_out/r2zz_launch_summary.txt | head -28
