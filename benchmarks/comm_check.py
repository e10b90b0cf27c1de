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
    """Per-call device time: (median, max) over ``iters`` individually timed calls, each the MAX over
    ranks.  (Round 1 reported the mean of a back-to-back loop: one slow call - 220 / 400 us at 32 / 64 MB
    on the NVLS variant - moved the whole row.)"""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for e0, e1 in evs:
        e0.record()
        fn()
        e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) for e0, e1 in evs], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    t = t.sort().values
    return float(t[len(t) // 2].item()), float(t[-1].item())


class BRay:
    """B-ray column: OUR emulation of what the reference does with the same bytes - one pickled message
    per (sender, receiver) pair staged through host memory (``/root/reference/spacy_ray/proxies.py:71-75,
    102-104`` over Ray's object store): device -> host copy, ``pickle.dumps`` (protocol 5), a CPU transport
    (gloo over loopback here; Ray uses its object store + gRPC), ``pickle.loads``, host -> device, and for
    gradients the add on the owner.  Wall-clock, max over ranks."""

    def __init__(self, rank, world, dev):
        self.rank, self.world, self.dev = rank, world, dev
        self.g = dist.new_group(backend="gloo")

    def _exchange(self, out_slices, on_recv):
        import pickle

        import numpy as np

        sends, recvs = [], []
        for p, sl in out_slices:
            blob = pickle.dumps(sl.cpu().numpy(), protocol=5)
            sends.append((p, torch.from_numpy(np.frombuffer(blob, dtype=np.uint8).copy())))
        n = sends[0][1].numel() if sends else 0
        for p, _ in out_slices:
            rb = torch.empty(n, dtype=torch.uint8)
            recvs.append((p, rb, dist.irecv(rb, p, group=self.g)))
        reqs = [dist.isend(t, p, group=self.g) for p, t in sends]
        for p, rb, r in recvs:
            r.wait()
            on_recv(p, torch.from_numpy(pickle.loads(rb.numpy().tobytes())).to(self.dev))
        for r in reqs:
            r.wait()

    def time(self, kind, full, local, shard, iters):
        import time as _t

        peers = [p for p in range(self.world) if p != self.rank]

        def rs():        # every rank pushes the owner's slice of its gradient; the owner adds them up
            self._exchange([(p, full[p * shard:(p + 1) * shard]) for p in peers], lambda p, t: local.add_(t))

        def ag():        # the owner pushes its updated shard to every peer
            self._exchange([(p, local) for p in peers],
                           lambda p, t: full[p * shard:(p + 1) * shard].copy_(t))

        fn = rs if kind == "rs" else ag
        sync = torch.cuda.synchronize if self.dev.type == "cuda" else (lambda: None)
        fn()
        sync()
        dist.barrier(group=self.g)
        t0 = _t.perf_counter()
        for _ in range(iters):
            fn()
        sync()
        dt = torch.tensor([(_t.perf_counter() - t0) / iters * 1e3], dtype=torch.float64)
        dist.all_reduce(dt, op=dist.ReduceOp.MAX, group=self.g)
        return float(dt.item())


def run_sweep(rank, world, dev, max_bytes, bray_max=1 << 26):
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
    bray = BRay(rank, world, dev) if bray_max > 0 else None
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
        t, tmax = {}, {}
        t["ours_rs"], tmax["ours_rs"] = time_op(ours_rs, iters)
        t["nccl_rs"], tmax["nccl_rs"] = time_op(nccl_rs, iters)
        epoch.zero_(); bar.zero_(); flags.zero_(); torch.cuda.synchronize(); dist.barrier()
        t["ours_ag"], tmax["ours_ag"] = time_op(ours_ag, iters)
        t["nccl_ag"], tmax["nccl_ag"] = time_op(nccl_ag, iters)
        moved = (world - 1) / world * shard * world * 4
        row = {"bytes": shard * world * 4, "rs_matches_nccl": bool(ok),
               **{k + "_us": v * 1e3 for k, v in t.items()},
               **{k + "_max_us": v * 1e3 for k, v in tmax.items()},
               **{k + "_GBps": moved / (v * 1e-3) / 1e9 for k, v in t.items()}}
        if bray is not None and nbytes <= bray_max:
            bi = 5 if nbytes <= (1 << 20) else 2
            row["bray_rs_us"] = bray.time("rs", full, local, shard, bi) * 1e3
            row["bray_ag_us"] = bray.time("ag", full, local, shard, bi) * 1e3
        rows.append(row)
        assert int(err.item()) == 0, f"collective timed out (code {int(err.item())})"
    return rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--sweep", action="store_true")
    ap.add_argument("--max-bytes", type=int, default=1 << 30)
    ap.add_argument("--bray-max-bytes", type=int, default=1 << 26,
                    help="largest payload for the host-staged B-ray emulation column (0 = off)")
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
        result["sweep"] = run_sweep(rank, world, dev, args.max_bytes, args.bray_max_bytes)
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
