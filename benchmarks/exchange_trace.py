#!/usr/bin/env python
"""Device-side timeline of one training step: where the bucketed exchange runs relative to the
backward pass (``%globaltimer`` stamps written by the exchange kernels themselves and by one-thread
stamp kernels at phase boundaries; SURVEY.md 5.1 "tile-level overlap traces").

    python benchmarks/exchange_trace.py --steps 40                         # 1 GPU
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 \
        benchmarks/exchange_trace.py --steps 40 --out gpurun_out/trace_8gpu

Writes <out>_rank<r>.json (median over the traced steps, all times in us relative to the step start)
and prints a table per rank-0.  Not a benchmark: tracing adds atomics and stamp kernels to the step.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
os.environ["SRB_COMM_TRACE"] = "1"

import numpy as np
import torch
import torch.distributed as dist


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--docs-per-gpu", type=int, default=1024)
    ap.add_argument("--out", default="gpurun_out/exchange_trace")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    import bench
    from spacy_ray_b200.config import Config
    from spacy_ray_b200.engine import Trainer
    from spacy_ray_b200.worker import Worker

    bargs = argparse.Namespace(docs_per_gpu=args.docs_per_gpu, width=256, depth=8, min_len=8, max_len=40, dropout=0.1)
    cfg = Config().from_str(bench.flagship_config(bargs, rank), interpolate=False)
    worker = Worker(cfg, rank=rank, num_workers=world, use_gpu=local, mode="sync")
    worker.set_proxy(None)
    comm = worker.proxy.comm
    exs = list(worker.train_corpus(worker.nlp))
    B = args.docs_per_gpu
    tr = Trainer(worker.nlp, worker.proxy, exs, docs_per_batch=B, dropout=0.1, bucket_rows=128)
    batches = tr.batches(args.warmup + args.steps, seed=rank, tokens_per_batch=B * 24)
    base = 32 * 8
    rows = []
    for i, ids in enumerate(batches):
        tr.train_step(ids, lag=0)
        if i < args.warmup:
            continue
        torch.cuda.synchronize()
        t = comm.trace.cpu().numpy().astype(np.int64)
        t0 = int(t[base + 0])
        rec = {"backward_enqueued_done": (int(t[base + 1]) - t0) / 1e3, "step_end": (int(t[base + 2]) - t0) / 1e3}
        for b in range(comm.plan.n):
            w = t[b * 8: b * 8 + 6]
            names = ["signal", "peers_seen", "reduce_start", "reduce_end", "update_start", "published"]
            for nm, v in zip(names, w):
                if 0 < int(v) < (1 << 62):
                    rec[f"b{b}.{nm}"] = (int(v) - t0) / 1e3
        for j, key in sorted(comm.trace_keys.items()):
            v = int(t[base + 8 + j])
            if v > 0:
                rec[f"grad{j:02d}.{key[1]}@{key[0]}"] = (v - t0) / 1e3
        rows.append(rec)
    tr.close()
    keys = sorted({k for r in rows for k in r}, key=lambda k: np.median([r[k] for r in rows if k in r]))
    med = {k: float(np.median([r[k] for r in rows if k in r])) for k in keys}
    out = {"rank": rank, "world": world, "steps": len(rows), "buckets": comm.plan.n,
           "bucket_keys": [[f"{k[1]}@{k[0]}" for k in ks] for ks in comm.plan.buckets],
           "owned_per_bucket": [r[1] - r[0] for r in comm.tables["ranges"]], "median_us": med}
    Path(args.out).parent.mkdir(parents=True, exist_ok=True)
    Path(f"{args.out}_rank{rank}.json").write_text(json.dumps(out, indent=1))
    if rank == 0:
        print(f"# step timeline, rank 0 of {world} (median of {len(rows)} steps, us after step start)")
        for k in keys:
            print(f"{med[k]:10.1f}  {k}")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
