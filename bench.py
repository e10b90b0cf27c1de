#!/usr/bin/env python
"""Headline benchmark: docs/sec training en tok2vec(HashEmbed+Maxout width 256, depth 8)+NER.

Contract (driver): ``python bench.py --gpus N --steps K --warmup W`` (N>1 under torchrun)
prints ONE JSON line from rank 0.  ``value`` = whole-job docs/s, device-timed with CUDA
events over exactly K steps, max over ranks.  ``e2e`` = the same metric through the public
API (``nlp.update`` + ``proxy.step``) including, every step, the pinned-host -> device copy
of that step's inputs and the device -> host read of the loss.

``--impl reference`` must run the unmodified reference from ``baseline/_ref``: it cannot be
imported here (spaCy / thinc / Ray are not installable offline - DESIGN.md), so that arm
reports ``unavailable``.  Two *in-repo* comparison arms exist instead and are labelled as
ours: ``--impl nccl-baseline`` (NCCL reduce-scatter/all-gather + cuBLAS GEMMs + unfused
torch Adam) and ``--impl rayproxy-emu`` (per-key host-staged async proxy protocol).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

BASELINE_PUBLISHED = None   # the reference publishes no numbers


def flagship_config(args, rank: int) -> str:
    n_docs = args.docs_per_gpu * 8
    return f"""
[system]
seed = 0

[nlp]
lang = "en"
pipeline = ["ner"]

[components]

[components.ner]
factory = "ner"

[components.ner.model]
@architectures = "spacy.TransitionBasedParser.v2"
state_type = "ner"
extra_state_tokens = false
hidden_width = 64
maxout_pieces = 2
use_upper = true

[components.ner.model.tok2vec]
@architectures = "spacy.Tok2Vec.v2"

[components.ner.model.tok2vec.embed]
@architectures = "spacy.MultiHashEmbed.v2"
width = {args.width}
attrs = ["NORM","PREFIX","SUFFIX","SHAPE"]
rows = [5000,1000,2500,2500]
include_static_vectors = false

[components.ner.model.tok2vec.encode]
@architectures = "spacy.MaxoutWindowEncoder.v2"
width = {args.width}
depth = {args.depth}
window_size = 1
maxout_pieces = 3

[corpora]

[corpora.train]
@readers = "spacy_ray_b200.SyntheticCorpus.v1"
n_docs = {n_docs}
seed = {1000 + rank}
min_len = {args.min_len}
max_len = {args.max_len}
n_ent_labels = 18
tasks = ["ner"]

[corpora.dev]
@readers = "spacy_ray_b200.SyntheticCorpus.v1"
n_docs = 64
seed = 7
min_len = {args.min_len}
max_len = {args.max_len}
n_ent_labels = 18
tasks = ["ner"]

[training]
dropout = {args.dropout}
max_steps = 0
"""


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "nccl-baseline", "rayproxy-emu"])
    ap.add_argument("--docs-per-gpu", type=int, default=1024)
    ap.add_argument("--width", type=int, default=256)
    ap.add_argument("--depth", type=int, default=8)
    ap.add_argument("--min-len", type=int, default=8)
    ap.add_argument("--max-len", type=int, default=40)
    ap.add_argument("--dropout", type=float, default=0.1)
    ap.add_argument("--no-e2e", dest="e2e", action="store_false")
    ap.add_argument("--comm", default="auto")
    ap.add_argument("--config", default=None,
                    help="train a pipeline from a .cfg file (configs/*.cfg) instead of the flagship tok2vec+NER; "
                         "its [corpora.train] must be a SyntheticCorpus (n_docs/seed are overridden per rank)")
    ap.add_argument("--engine", default="graph", choices=["graph", "eager"],
                    help="graph = CUDA-graph replay of the whole step; eager = same kernels launched from Python")
    args = ap.parse_args()

    if args.impl == "reference":
        print(json.dumps({
            "impl": "reference",
            "unavailable": "baseline/_ref/spacy_ray installs only with --no-deps; import fails: spacy, thinc, ray "
                           "(ray<1.0 has no cp312 wheel) are absent from /opt/wheelhouse and there is no network",
        }))
        return 0

    import torch

    if not torch.cuda.is_available():
        print(json.dumps({"error": "bench.py needs a CUDA device", "impl": args.impl}))
        return 1
    if args.warmup < 3:
        args.warmup = 3
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if world != args.gpus and world > 1:
        args.gpus = world
    torch.cuda.set_device(local_rank)
    import torch.distributed as dist

    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))

    from spacy_ray_b200.config import Config
    from spacy_ray_b200.utils.timing import ClockSampler
    from spacy_ray_b200.worker import Worker

    fused = args.impl == "ours"
    if args.impl == "nccl-baseline":
        os.environ["SRB_USE_TC"] = "0"          # cuBLAS GEMMs, library collectives, per-key torch Adam
    mode = "async" if args.impl == "rayproxy-emu" else "sync"
    comm = args.comm
    if args.impl == "nccl-baseline":
        comm = "local" if world == 1 else "dist"
    if args.impl == "rayproxy-emu" and world > 1:
        print(json.dumps({"impl": args.impl, "unavailable": "rayproxy-emu runs under the actor runtime: "
                          "use benchmarks/bench_rayproxy.py"}))
        return 0
    if args.config:
        cfg = Config().from_str(Path(args.config).read_text(), interpolate=False)
        cfg["corpora"]["train"]["n_docs"] = args.docs_per_gpu * 8
        cfg["corpora"]["train"]["seed"] = 1000 + rank
        cfg["training"]["max_steps"] = 0
    else:
        cfg = Config().from_str(flagship_config(args, rank), interpolate=False)
    worker = Worker(cfg, rank=rank, num_workers=world, use_gpu=local_rank, mode="sync", comm=comm,
                    fused_ops=True)
    worker.set_proxy(None)
    nlp, proxy = worker.nlp, worker.proxy
    from spacy_ray_b200.ops import get_current_ops

    ops = get_current_ops()
    examples = list(worker.train_corpus(nlp))
    B = args.docs_per_gpu
    n_params = sum(proxy.layout.numel.values())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms: float) -> float:
        if world == 1:
            return ms
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    from spacy_ray_b200.engine import Trainer

    use_graphs = args.engine == "graph"
    trainer = Trainer(nlp, proxy, examples, docs_per_batch=B, dropout=args.dropout, use_graphs=use_graphs)
    n_total = args.warmup + args.steps
    id_batches = trainer.batches(2 * n_total + 2, seed=rank)

    # ---------------- device-timed: inputs already on the device, no host reads ----------------
    # (each step's packed input block is staged on the device beforehand and moved into the
    #  static input buffer with a D2D copy inside the timed region)
    dev_inputs = []
    for ids in id_batches[:n_total]:
        trainer.prepare(ids)
        stage, err = trainer._q_out.get()
        assert err is None, err
        dev_inputs.append((stage["buf"].to(trainer.device), stage["rows"], stage["docs"], stage["words"]))
    torch.cuda.synchronize()

    def device_step(item):
        packed, rows, _d, _w = item
        trainer.dev_buf.copy_(packed, non_blocking=True)
        return trainer._run(rows)

    device_step(dev_inputs[0])                       # first step runs eagerly on every rank
    trainer.capture_buckets([it[1] for it in dev_inputs] +
                            [trainer.rows_for(ids) for ids in id_batches[n_total:]])
    for item in dev_inputs[1: args.warmup]:
        device_step(item)
    barrier()
    launches0 = ops.launches + getattr(proxy.comm, "launches", 0)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank, period_s=0.05) as clocks:
        ev0.record()
        for item in dev_inputs[args.warmup: n_total]:
            device_step(item)
        ev1.record()
        barrier()
    ms = max_over_ranks(float(ev0.elapsed_time(ev1)))
    launches = ops.launches + getattr(proxy.comm, "launches", 0) - launches0
    docs = sum_over_ranks(float(sum(it[2] for it in dev_inputs[args.warmup: n_total])))
    words = sum_over_ranks(float(sum(it[3] for it in dev_inputs[args.warmup: n_total])))
    value = docs / (ms / 1e3)
    if hasattr(proxy.comm, "check"):
        proxy.comm.check()

    # ---------------- end to end through the public API -----------------------------------------
    # Trainer.train_step(): native collate into pinned memory (prefetch thread) -> ONE H2D copy
    # -> CUDA-graph replay of the whole step -> D2H copy of the per-head losses.  Every step; the host
    # reads each loss one step late (it is logging data) so the device never waits for the host.
    e2e = None
    if args.e2e:
        rest = id_batches[n_total: 2 * n_total + 1]
        trainer.prepare(rest[0])
        for i in range(args.warmup):
            trainer.prepare(rest[i + 1])
            trainer.train_step()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        docs_local = 0
        e0.record()
        for i in range(args.warmup, n_total):
            trainer.prepare(rest[i + 1])          # prefetch the NEXT batch while this one runs
            loss_val = trainer.train_step()       # H2D + step + async D2H(loss); returns the previous step's
            docs_local += trainer.last["docs"]
        loss_val = trainer.flush_loss()           # the last step's loss is read inside the timed region too
        e1.record()
        barrier()
        ms_e = max_over_ranks(float(e0.elapsed_time(e1)))
        docs_e = sum_over_ranks(float(docs_local))
        e2e = {"value": docs_e / (ms_e / 1e3), "unit": "docs/s", "h2d_bytes_per_step": int(trainer.h2d_bytes_per_step),
               "d2h_bytes_per_step": 4 * len(trainer.loss_names), "ms_per_step": ms_e / args.steps, "last_loss": loss_val,
               "api": "spacy_ray_b200.engine.Trainer.train_step"}
        if hasattr(proxy.comm, "check"):
            proxy.comm.check()
    trainer.close()

    if rank == 0:
        mean_len = words / max(docs, 1)
        out = {
            "metric": "docs/sec (whole box, device-timed, max over ranks) "
                      + ("en tok2vec+NER" if not args.config else "+".join(nlp.pipe_names)),
            "value": value, "unit": "docs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic (SyntheticCorpus, random-init weights)",
            "impl": args.impl, "engine": args.engine,
            "config": {
                "model": (f"en tok2vec(MultiHashEmbed+MaxoutWindowEncoder width={args.width} depth={args.depth})+NER "
                          f"(TransitionBasedParser hidden=64, 18 entity labels)") if not args.config
                else f"{args.config} pipeline={nlp.pipe_names}",
                "global_batch": int(B * world), "docs_per_gpu": B, "seq_len": round(mean_len, 2),
                "words_per_sec": words / (ms / 1e3), "params": int(n_params),
                "parallelism": f"dp{world} + optimizer sharding by parameter ownership ({proxy.comm.name})",
                "l2": "per-step inputs + activations exceed the 126 MB L2 (fresh batch every step)",
                "optimizer": "Adam (thinc semantics, per-tensor clip 1.0, wd 0.01), fp32 master",
            },
            "clocks": clocks.summary(),
            "e2e": e2e,
            "gpu_launches": int(launches),
        }
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
