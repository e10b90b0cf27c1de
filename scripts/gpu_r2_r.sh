#!/bin/bash
# round 2, call R (4 GPUs): BASELINE config 5 sweep at N=4 with the B-ray (host-staged, pickled) column
set -u
mkdir -p gpurun_out
SRB_NVLS=0 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29641 benchmarks/comm_check.py --sweep --out gpurun_out/r2r_sweep4_p2p.json > gpurun_out/r2r_sweep4_p2p.log 2>&1
tail -3 gpurun_out/r2r_sweep4_p2p.log | cut -c1-300
SRB_NVLS=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29642 benchmarks/comm_check.py --sweep --bray-max-bytes 0 --out gpurun_out/r2r_sweep4_nvls.json > gpurun_out/r2r_sweep4_nvls.log 2>&1
tail -3 gpurun_out/r2r_sweep4_nvls.log | cut -c1-300
