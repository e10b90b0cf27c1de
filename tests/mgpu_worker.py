#!/usr/bin/env python
"""Per-rank body of the multi-GPU tests (``test_multigpu.py`` launches it under
``python -m torch.distributed.run``; it can also be run by hand).  Each case prints one JSON line
from rank 0 and exits non-zero on failure.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29631 \
        tests/mgpu_worker.py --case exchange
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import torch
import torch.distributed as dist


def flagship_like_layout(world: int, scale: int = 1, balance: str = "lpt"):
    """Key sizes of the flagship model (tok2vec w256 d8 + NER head), divided like the real thing."""
    from spacy_ray_b200.parallel.sync_proxy import ALIGN, FlatLayout, _round_up

    sizes = [("E", 5000 * 256), ("E", 1000 * 256), ("E", 2500 * 256), ("E", 2500 * 256), ("W", 768 * 1024), ("b", 768),
             ("G", 256), ("b", 256)]
    for _ in range(8):
        sizes += [("W", 768 * 768), ("b", 768), ("G", 256), ("b", 256)]
    sizes += [("W", 64 * 256), ("b", 64), ("W", 3 * 64 * 2 * 64), ("b", 128), ("pad", 384), ("W", 73 * 64), ("b", 73)]
    sizes = [(n, max(1, s // scale)) for n, s in sizes]
    keys = [(i + 1, n) for i, (n, _s) in enumerate(sizes)]
    numel = {k: s for k, (_n, s) in zip(keys, sizes)}
    load = [0] * world
    owner = {}
    if balance == "lpt":
        for k in sorted(keys, key=lambda k: (-numel[k], k[0])):
            r = min(range(world), key=lambda j: (load[j], j))
            owner[k] = r
            load[r] += numel[k]
    else:                                   # reference-style: consecutive runs, leftovers to the last rank
        n = max(1, len(keys) // world)
        for i, k in enumerate(keys):
            owner[k] = min(i // n, world - 1)
    order, offset, shape, shard_len = [], {}, {}, []
    for r in range(world):
        pos = 0
        for k in keys:
            if owner[k] == r:
                order.append(k)
                offset[k] = pos
                shape[k] = (numel[k],)
                pos += _round_up(numel[k], ALIGN)
        shard_len.append(pos)
    cap = _round_up(max(max(shard_len), ALIGN), ALIGN)
    starts = [r * cap for r in range(world)]
    for k in order:
        offset[k] += starts[owner[k]]
    backward_order = list(reversed(keys))   # head first, embedding tables last
    return FlatLayout(order, owner, offset, numel, shape, starts, shard_len, cap, world), backward_order


def make_optimizer(kind: str):
    from spacy_ray_b200.training.optimizer import Optimizer

    kw = dict(L2=0.01, grad_clip=1.0)
    if kind == "radam":
        return Optimizer(0.01, use_radam=True, **kw)
    if kind == "sgd":
        return Optimizer(0.05, use_adam=False, **kw)
    if kind == "adam_avg":
        return Optimizer(0.01, use_averages=True, **kw)
    return Optimizer(0.01, **kw)


def case_exchange(rank, world, dev, args):
    """Bucketed fused exchange vs NCCL reduce-scatter + per-key reference optimizer + all-gather."""
    from spacy_ray_b200.ops.torch_ops import TorchOps
    from spacy_ray_b200.parallel.fused_comm import FusedSymmComm
    from spacy_ray_b200.parallel.sync_proxy import ShardedSyncProxy

    layout, order = flagship_like_layout(world, scale=args.scale, balance=args.balance)
    opt, ref_opt = make_optimizer(args.opt), make_optimizer(args.opt)
    ref_opt.ops = TorchOps(str(dev), dtype=torch.float32)
    comm = FusedSymmComm(rank, world, layout, dev, optimizer=opt, timeout_s=10.0, n_buckets=args.buckets)
    proxy = ShardedSyncProxy(layout, opt, rank=rank, world_size=world, device=dev, comm=comm,
                             param_dtype=torch.bfloat16, buffers=comm.buffers)
    comm.bind(proxy)
    g0 = torch.Generator(device="cpu").manual_seed(123)
    init = torch.randn(layout.total, generator=g0) * 0.1
    for k in layout.keys:                     # padding stays zero
        o, n = layout.offset[k], layout.numel[k]
        proxy.param_flat[o:o + n] = init[o:o + n].to(dev).bfloat16()
    s0 = layout.shard_start[rank]
    owned = layout.owned_keys(rank)
    for k in owned:
        o, n = layout.offset[k], layout.numel[k]
        proxy.master[o - s0:o - s0 + n] = proxy.param_flat[o:o + n].float()
    ref_w = {k: proxy.master[layout.offset[k] - s0:layout.offset[k] - s0 + layout.numel[k]].clone() for k in owned}
    ref_param = proxy.param_flat.clone()
    torch.cuda.synchronize()
    dist.barrier()
    worst = 0.0
    for step in range(args.steps):
        gg = torch.Generator(device="cpu").manual_seed(1000 * step + rank)
        grad = torch.zeros(layout.total)
        for i, k in enumerate(layout.keys):
            o, n = layout.offset[k], layout.numel[k]
            sc = 3.0 if i % 7 == 0 else (1e-4 if i % 7 == 1 else 0.02)     # above / far below the clip threshold
            grad[o:o + n] = torch.randn(n, generator=gg) * sc
        grad = grad.to(dev)
        # ---- reference: NCCL reduce-scatter + per-key optimizer + all-gather
        shard = torch.empty(layout.shard_cap, device=dev)
        dist.reduce_scatter_tensor(shard, grad.clone())
        mine = torch.zeros(layout.shard_cap, device=dev)
        for k in owned:
            o, n = layout.offset[k] - s0, layout.numel[k]
            ref_opt(k, ref_w[k], shard[o:o + n].clone())
            mine[o:o + n] = ref_w[k]
        dist.all_gather_into_tensor(ref_param, mine.bfloat16())
        # ---- ours: gradients arrive key by key ("backward pass"), buckets fire as they complete
        proxy.begin_step(overlap=True)
        for k in order:
            o, n = layout.offset[k], layout.numel[k]
            proxy.inc_grad(k[0], k[1], grad[o:o + n])
        proxy.step()
        proxy.quiesce()
        torch.cuda.synchronize()
        comm.check()
        dist.barrier()                        # every rank's remote zero stores have landed
        err = (proxy.param_flat.float() - ref_param.float()).abs().max().item()
        merr = max((proxy.master[layout.offset[k] - s0:layout.offset[k] - s0 + layout.numel[k]] - ref_w[k]).abs().max().item()
                   for k in owned) if owned else 0.0
        worst = max(worst, err, merr)
        chk = proxy.param_flat.float().sum().reshape(1).double()
        allc = [torch.zeros_like(chk) for _ in range(world)]
        dist.all_gather(allc, chk)
        assert all(float(c) == float(allc[0]) for c in allc), "ranks disagree on weights"
    # the accumulators are cleared at the start of the NEXT step (once the owners have published)
    proxy.begin_step(overlap=False)
    torch.cuda.synchronize()
    dist.barrier()
    assert float(proxy.grad_flat.abs().sum()) == 0.0, "gradient buffer not cleared"
    tol = 1e-2 if args.opt == "sgd" else 4e-3          # one bf16 ulp of weights ~0.5 is 2e-3
    assert worst < tol, f"max abs err vs NCCL path {worst}"
    return {"case": "exchange", "steps": args.steps, "max_abs_err_vs_nccl_path": worst, "world": world,
            "nvls": bool(comm.grad_mc), "buckets": comm.plan.n, "opt": args.opt, "balance": args.balance,
            "params": int(sum(layout.numel.values())), "launches": comm.launches}


def case_gate(rank, world, dev, args):
    """A consumer can never read a stale weight: the LAST rank publishes late (a test hook holds its
    weight stores back by 50 ms per bucket); every other rank immediately runs a gated
    tcgen05 GEMM and a gated hash-embed on weights that rank owns.  Results must equal the ones
    computed after a full quiesce, bit for bit - and differ from what the OLD weights give."""
    from spacy_ray_b200.ops.b200_ops import B200Ops, EPI_STORE, MODE_KK
    from spacy_ray_b200.parallel.fused_comm import FusedSymmComm
    from spacy_ray_b200.parallel.sync_proxy import ShardedSyncProxy

    ops = B200Ops(str(dev))
    layout, order = flagship_like_layout(world, scale=1, balance="lpt")
    opt = make_optimizer("adam")
    comm = FusedSymmComm(rank, world, layout, dev, optimizer=opt, timeout_s=20.0, n_buckets=6, ops=ops)
    ops.gate_provider = comm
    proxy = ShardedSyncProxy(layout, opt, rank=rank, world_size=world, device=dev, comm=comm,
                             param_dtype=torch.bfloat16, buffers=comm.buffers)
    comm.bind(proxy)
    g0 = torch.Generator(device="cpu").manual_seed(5)
    proxy.param_flat.copy_((torch.randn(layout.total, generator=g0) * 0.1).to(dev).bfloat16())
    s0 = layout.shard_start[rank]
    proxy.master.copy_(proxy.param_flat[s0:s0 + layout.shard_cap].float())
    torch.cuda.synchronize()
    dist.barrier()
    late = world - 1
    # a (768, 768) weight and an embedding table owned by the late rank
    wkey = next(k for k in layout.keys if layout.owner[k] == late and layout.numel[k] == 768 * 768)
    ekey = next((k for k in layout.keys if layout.owner[k] == late and k[1] == "E"), None)
    X = torch.randn(512, 768, device=dev, generator=torch.Generator(device=dev).manual_seed(1)).bfloat16()
    attrs = torch.randint(0, 1 << 40, (512, 4), dtype=torch.int64, device=dev)
    mask = torch.ones(512, 1, device=dev)

    def forward():
        W = proxy.get_param(*wkey).view(768, 768)
        out = torch.empty(512, 768, dtype=torch.bfloat16, device=dev)
        ops.tc_gemm(X, W, out, mode=MODE_KK, epi=EPI_STORE, block_n=256, M=512, N=768, K=768, gate=ops._gate(W))
        emb = None
        if ekey is not None:
            E = proxy.get_param(*ekey).view(-1, 256)
            emb = ops.multi_hash_embed(attrs, mask, [E], [3], [0])
        return out, emb

    results = {}
    for step in range(3):
        old_out, old_emb = forward()                       # weights of the previous epoch
        gg = torch.Generator(device="cpu").manual_seed(77 * step + rank)
        grad = (torch.randn(layout.total, generator=gg) * 0.5).to(dev)
        proxy.begin_step(overlap=False)                     # (clears the accumulators of the previous epoch)
        for k in layout.keys:
            o, n = layout.offset[k], layout.numel[k]
            proxy.inc_grad(k[0], k[1], grad[o:o + n].view(layout.shape[k]))
        # from step 1 on the late rank sits 50 ms between its reduce phase and its weight stores:
        # everybody else has long finished its own exchange kernels and launched the next forward
        comm.test_delay_us = 50_000 if (rank == late and step > 0) else 0
        proxy.step()
        new_out, new_emb = forward()                       # gated: must see the NEW weights
        proxy.quiesce()
        torch.cuda.synchronize()
        comm.check()
        dist.barrier()
        chk_out, chk_emb = forward()
        torch.cuda.synchronize()
        assert torch.equal(new_out, chk_out), f"rank {rank} step {step}: gated GEMM read stale weights"
        assert not torch.equal(new_out, old_out), "the exchange did not change the weights"
        if new_emb is not None:
            assert torch.equal(new_emb, chk_emb), f"rank {rank} step {step}: gated hash-embed read stale weights"
            assert not torch.equal(new_emb, old_emb)
        results[step] = True
    return {"case": "gate", "world": world, "steps": len(results), "late_rank": late, "nvls": bool(comm.grad_mc)}


def case_timeout(rank, world, dev, args):
    """A peer that never arrives must produce error code 1 within the timeout, not a hang."""
    from spacy_ray_b200.parallel.fused_comm import FusedSymmComm
    from spacy_ray_b200.parallel.sync_proxy import ShardedSyncProxy

    layout, order = flagship_like_layout(world, scale=64)
    opt = make_optimizer("adam")
    comm = FusedSymmComm(rank, world, layout, dev, optimizer=opt, timeout_s=1.5, n_buckets=3)
    proxy = ShardedSyncProxy(layout, opt, rank=rank, world_size=world, device=dev, comm=comm,
                             param_dtype=torch.bfloat16, buffers=comm.buffers)
    comm.bind(proxy)
    proxy.step()                                           # everyone: fine
    torch.cuda.synchronize()
    comm.check()
    dist.barrier()
    t0 = time.time()
    raised = None
    if rank != world - 1:                                  # the last rank "dies" (never launches its exchange)
        proxy.step()
        torch.cuda.synchronize()
        try:
            comm.check()
        except RuntimeError as e:
            raised = str(e)
    dt = time.time() - t0
    dist.barrier()
    if rank != world - 1:
        owns = len(layout.owned_keys(rank)) > 0
        if owns:
            assert raised is not None and "code 1" in raised, f"expected a timeout error, got {raised!r}"
        assert dt < 15.0, f"took {dt:.1f}s"
    return {"case": "timeout", "world": world, "seconds": round(dt, 2), "error": raised}


def case_train(rank, world, dev, args):
    """End to end: Worker + engine.Trainer on every rank; weights identical across ranks after
    every step, loss goes down, nothing timed out, and the step contains no NCCL kernel."""
    import numpy as np
    from conftest import multi_cfg
    from spacy_ray_b200.config import Config
    from spacy_ray_b200.engine import Trainer
    from spacy_ray_b200.worker import Worker

    text = multi_cfg([args.pipe], width=64, depth=2, n_docs=1024, max_len=16, hidden=64)
    w = Worker(Config().from_str(text, interpolate=False), rank=rank, num_workers=world, use_gpu=rank, mode="sync")
    w.set_proxy(None)
    assert w.proxy.comm.name == "fused"
    exs = list(w.train_corpus(w.nlp))
    tr = Trainer(w.nlp, w.proxy, exs, docs_per_batch=64, dropout=0.1, prefetch=False)
    hist = []
    for step in range(args.steps):
        ids = (np.arange(64, dtype=np.int64) + 64 * (step * world + rank)) % len(exs)
        hist.append(tr.train_step(np.sort(ids), lag=0))
        if step % 10 == 0 or step == args.steps - 1:
            w.proxy.quiesce()
            torch.cuda.synchronize()
            chk = w.proxy.param_flat.float().abs().sum().reshape(1).double()
            allc = [torch.zeros_like(chk) for _ in range(world)]
            dist.all_gather(allc, chk)
            assert all(float(c) == float(allc[0]) for c in allc), f"ranks disagree on weights at step {step}"
    w.proxy.comm.check()
    tr.close()
    assert all(h == h for h in hist), hist
    assert sum(hist[-5:]) < 0.7 * sum(hist[:5]), hist[::5]
    return {"case": "train", "world": world, "pipe": args.pipe, "first": hist[0], "last": hist[-1],
            "buckets": w.proxy.comm.plan.n, "graphs": len(tr._graphs)}


CASES = {"exchange": case_exchange, "gate": case_gate, "timeout": case_timeout, "train": case_train}


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", required=True, choices=sorted(CASES))
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--scale", type=int, default=1)
    ap.add_argument("--buckets", type=int, default=6)
    ap.add_argument("--opt", default="adam")
    ap.add_argument("--balance", default="lpt")
    ap.add_argument("--pipe", default="ner")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    try:
        out = CASES[args.case](rank, world, dev, args)
    except Exception as e:          # make the failing rank visible in the launcher's output
        import traceback

        print(f"[rank {rank}] FAILED: {e}\n{traceback.format_exc()}", flush=True)
        os._exit(1)
    if rank == 0:
        print("MGPU_RESULT " + json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
