#!/bin/bash
# end-of-round verification: build check, all GPU tests, smoke(), default bench + the other configs
set -u
mkdir -p gpurun_out
S=gpurun_out/summary_final.txt
: > $S
echo "=== tests" | tee -a $S
timeout -s KILL 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/test_all.log 2>&1
echo "exit=$? $(tail -n 1 gpurun_out/test_all.log)" | tee -a $S
grep -E "^(FAILED|ERROR)" gpurun_out/test_all.log | tee -a $S
echo "=== smoke" | tee -a $S
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee -a $S
echo "=== bench default" | tee -a $S
timeout -s KILL 600 python bench.py > gpurun_out/bench_default.log 2>&1; echo "exit=$? $(tail -n 1 gpurun_out/bench_default.log | cut -c1-260)" | tee -a $S
echo "=== bench --impl reference" | tee -a $S
python bench.py --impl reference | tee -a $S
for c in tagger_w96 parser_w256 multitask_w512 ner_w256; do
  echo "=== bench $c" | tee -a $S
  timeout -s KILL 400 python bench.py --steps 50 --warmup 5 --config configs/$c.cfg > gpurun_out/bench_$c.log 2>&1
  echo "exit=$? $(tail -n 1 gpurun_out/bench_$c.log | cut -c1-260)" | tee -a $S
done
