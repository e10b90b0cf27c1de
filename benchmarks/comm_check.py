#!/usr/bin/env python
"""Multi-GPU check + bandwidth sweep for the peer-memory collectives (run under torchrun).

1. ``--check``: fused reduce-scatter + Adam + all-gather kernel vs the library path
   (``reduce_scatter_tensor`` -> per-key TorchOps Adam -> ``all_gather_into_tensor``) on a
   synthetic flat layout with the key-size distribution of the flagship model; several steps,
   bit-for-bit identical weights on all ranks required, tolerance vs the reference.
2. ``--sweep``: BASELINE.json config 5 - reduce-scatter / all-gather bandwidth 1 KB - 1 GB for
   our P2P kernels (and the NVLS multimem variants with SRB_NVLS=1) vs NCCL.  Device-timed
   (CUDA events), max over ranks, algorithmic bus bandwidth = (W-1)/W * bytes / t.

    torchrun --nproc-per-node 2 --master-addr 127.0.0.1 benchmarks/comm_check.py --check --sweep
"""
from __future__ import annotations

import argparse
import json
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch
import torch.distributed as dist


def run_check(rank, world, dev, steps=4):
    """The correctness check lives with the tests (``tests/mgpu_worker.py``, also run by
    ``pytest -m multigpu``): bucketed fused exchange vs NCCL reduce-scatter + per-key reference
    optimizer + all-gather on the flagship key-size distribution."""
    sys.path.insert(0, str(ROOT / "tests"))
    import mgpu_worker

    args = argparse.Namespace(steps=steps, scale=1, buckets=6, opt="adam", balance="lpt", pipe="ner")
    return mgpu_worker.case_exchange(rank, world, dev, args)


def time_op(fn, iters, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def run_sweep(rank, world, dev, max_bytes):
    import torch.distributed._symmetric_memory as symm_mem
    from spacy_ray_b200.ops.b200_ops import load_extension

    load_extension()
    rows = []
    sizes = [1 << p for p in range(10, 31) if (1 << p) <= max_bytes]
    max_elems = max(sizes) // 4
    buf = symm_mem.empty(max_elems, dtype=torch.float32, device=dev)
    hdl = symm_mem.rendezvous(buf, dist.group.WORLD)
    flags = symm_mem.empty(1024, dtype=torch.int32, device=dev)
    hf = symm_mem.rendezvous(flags, dist.group.WORLD)
    flags.zero_()
    buf.normal_()
    torch.cuda.synchronize()
    dist.barrier()
    use_nvls = os.environ.get("SRB_NVLS", "0") == "1"
    mc = int(hdl.multicast_ptr or 0) if use_nvls else 0
    ptrs = [int(p) for p in hdl.buffer_ptrs]
    fptrs = [int(p) for p in hf.buffer_ptrs]
    epoch = torch.zeros(1, dtype=torch.int32, device=dev)
    bar = torch.zeros(1, dtype=torch.int32, device=dev)
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    for nbytes in sizes:
        elems = nbytes // 4
        shard = max(4, (elems // world) // 4 * 4)
        local = torch.zeros(shard, dtype=torch.float32, device=dev)
        grid = max(1, min(2 * sms, (shard // 16 + 255) // 256))
        # grid must stay constant for the barrier counter: reset the counters per size
        epoch.zero_(); bar.zero_(); flags.zero_()
        torch.cuda.synchronize(); dist.barrier()
        iters = 20 if nbytes <= (1 << 24) else 5

        def ours_rs():
            torch.ops.srb.p2p_collective(0, ptrs, fptrs, mc, local, epoch, bar, err, shard, rank, grid, 10.0)

        def ours_ag():
            torch.ops.srb.p2p_collective(1, ptrs, fptrs, mc, local, epoch, bar, err, shard, rank, grid, 10.0)

        full = buf[: shard * world]
        out_n = torch.empty(shard, dtype=torch.float32, device=dev)

        def nccl_rs():
            dist.reduce_scatter_tensor(out_n, full)

        def nccl_ag():
            dist.all_gather_into_tensor(full, out_n)

        # correctness of our RS against NCCL at this size
        ours_rs(); nccl_rs(); torch.cuda.synchronize()
        ok = torch.allclose(local, out_n, rtol=1e-4, atol=1e-4)
        t = {"ours_rs": time_op(ours_rs, iters), "nccl_rs": time_op(nccl_rs, iters)}
        epoch.zero_(); bar.zero_(); flags.zero_(); torch.cuda.synchronize(); dist.barrier()
        t["ours_ag"] = time_op(ours_ag, iters)
        t["nccl_ag"] = time_op(nccl_ag, iters)
        moved = (world - 1) / world * shard * world * 4
        rows.append({"bytes": shard * world * 4, "rs_matches_nccl": bool(ok),
                     **{k + "_us": v * 1e3 for k, v in t.items()},
                     **{k + "_GBps": moved / (v * 1e-3) / 1e9 for k, v in t.items()}})
        assert int(err.item()) == 0, f"collective timed out (code {int(err.item())})"
    return rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--sweep", action="store_true")
    ap.add_argument("--max-bytes", type=int, default=1 << 30)
    ap.add_argument("--out", default="gpurun_out/comm_check.json")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local_rank)
    dev = torch.device(f"cuda:{local_rank}")
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
    result = {"world": world, "nvls_env": os.environ.get("SRB_NVLS", "0")}
    if args.check:
        result["fused_step"] = run_check(rank, world, dev)
    if args.sweep:
        result["sweep"] = run_sweep(rank, world, dev, args.max_bytes)
    if rank == 0:
        Path(args.out).parent.mkdir(parents=True, exist_ok=True)
        Path(args.out).write_text(json.dumps(result, indent=1))
        print(json.dumps({k: v for k, v in result.items() if k != "sweep"}))
        for row in result.get("sweep", []):
            print({k: (round(v, 1) if isinstance(v, float) else v) for k, v in row.items()})
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
