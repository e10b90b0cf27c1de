#!/bin/bash
# round 2, call P (1 GPU): arena-init kernel numerics, launch lists of all four configs, and one
# `ncu --set full` capture per hot kernel of the flagship step (exports go to profiles/ afterwards)
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parser.py -q -x -k "biluo or arc or parser or tagger or softmax or training or worker or width_96" 2>&1 | tail -4
for c in flagship tagger_w96 parser_w256 multitask_w512; do
  if [ $c = flagship ]; then CF=""; else CF="--config configs/$c.cfg"; fi
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 400 --csv --log-file gpurun_out/r2p_launches_$c.csv python bench.py $CF --steps 3 --warmup 3 --no-e2e --no-own-baseline > gpurun_out/r2p_ncu_$c.log 2>&1
  python scripts/launch_summary.py gpurun_out/r2p_launches_$c.csv > gpurun_out/r2p_launch_summary_$c.txt 2>&1
  echo "== $c"; head -4 gpurun_out/r2p_launch_summary_$c.txt; grep -v "srb::" gpurun_out/r2p_launch_summary_$c.txt | tail -n +2 | head -8
done
cap() { name=$1; pat=$2; skip=$3
  timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:$pat -s $skip -c 1 -f -o gpurun_out/r2p_$name \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-own-baseline > gpurun_out/r2p_cap_$name.log 2>&1
  ls -la gpurun_out/r2p_$name.ncu-rep 2>&1 | awk '{print $5, $9}'
}
cap gemm_fwd gemm_kernelILi192ELi0ELi1ELi2ELb1ELb1 20
cap gemm_dx gemm_kernelILi128ELi2ELi0ELi2ELb1ELb1 20
cap gemm_dw gemm_kernelILi256ELi1ELi2ELi2ELb1ELb0 20
cap ln_fwd maxout_ln_fwd_vec_kernel 20
cap ln_bwd maxout_ln_bwd_vec_kernel 20
cap embed_fwd hash_embed_fwd_kernel 3
cap embed_bwd hash_embed_bwd_sorted_kernel 3
cap biluo biluo_block_kernel 3
cap bucket_update bucket_update_kernel 8
cap bucket_reduce bucket_reduce_kernel 8
