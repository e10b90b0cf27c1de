#!/bin/bash
# end-of-round verification (1 GPU): all GPU tests, smoke(), default bench (with the own-NCCL arm), the
# reference arm, the other configs; everything the driver runs at round end plus the per-config points
set -u
mkdir -p gpurun_out
S=gpurun_out/r2_final_verification.txt
: > $S
echo "=== tests (pytest tests -m gpu)" | tee -a $S
timeout -s KILL 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2f_test_all.log 2>&1
echo "exit=$? $(tail -n 1 gpurun_out/r2f_test_all.log)" | tee -a $S
grep -E "^(FAILED|ERROR)" gpurun_out/r2f_test_all.log | tee -a $S
echo "=== smoke()" | tee -a $S
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee -a $S
echo "=== python bench.py (defaults)" | tee -a $S
timeout -s KILL 600 python bench.py > gpurun_out/r2f_bench_default.json 2> gpurun_out/r2f_bench_default.err; echo "exit=$?" | tee -a $S
tail -n 1 gpurun_out/r2f_bench_default.json | tee -a $S
echo "=== python bench.py --impl reference" | tee -a $S
timeout -s KILL 300 python bench.py --impl reference 2>/dev/null | tail -1 | tee -a $S
for c in tagger_w96 parser_w256 multitask_w512 ner_w256; do
  echo "=== bench $c" | tee -a $S
  timeout -s KILL 400 python bench.py --steps 100 --warmup 10 --no-own-baseline --config configs/$c.cfg > gpurun_out/r2f_bench_$c.json 2> gpurun_out/r2f_bench_$c.err
  echo "exit=$? $(tail -n 1 gpurun_out/r2f_bench_$c.json | cut -c1-330)" | tee -a $S
done
echo "=== kernels of one flagship step (ncu gpu__time_duration; library kernels listed explicitly)" | tee -a $S
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 300 --csv --log-file gpurun_out/r2f_launches.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-own-baseline > gpurun_out/r2f_ncu.log 2>&1
python scripts/launch_summary.py gpurun_out/r2f_launches.csv > gpurun_out/r2f_launch_summary.txt 2>&1
head -30 gpurun_out/r2f_launch_summary.txt | tee -a $S
