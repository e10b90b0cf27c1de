#!/usr/bin/env python
"""B-ray: OUR emulation of the reference's Ray-proxy dataflow.

The reference cannot be run here (spaCy/thinc/Ray are not installable), so this measures
the same *protocol* on the same hardware with our own code, and must always be labelled as
our emulation, never as "spacy-ray's numbers":

* one message per parameter tensor: gradient unicast to the owner
  (``/root/reference/spacy_ray/proxies.py:102-104``), updated parameter unicast to each of
  the N-1 peers (``proxies.py:71-75``);
* device -> host staging per message (what pickling a GPU array through Ray's object store
  costs) and host -> device on adoption;
* per-key unfused Adam on the owner, lazily at the next read (``proxies.py:126-128``);
* quorum 2 (reference default, ``proxies.py:33``) or N (synchronous, ``worker.py:151-155``).

Workers are actor processes on the built-in runtime (``parallel/actors.py``); throughput is
wall-clock docs/s summed over workers (there is no device-side notion of a step here).

    python benchmarks/bench_rayproxy.py --workers 2 --gpu --quorum 2 --steps 30
"""
from __future__ import annotations

import argparse
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--workers", type=int, default=2)
    ap.add_argument("--gpu", action="store_true")
    ap.add_argument("--quorum", type=int, default=2)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--docs-per-batch", type=int, default=1024)
    ap.add_argument("--width", type=int, default=256)
    ap.add_argument("--depth", type=int, default=8)
    ap.add_argument("--mode", default="async", choices=["async", "sync"])
    ap.add_argument("--comm", default="auto")
    args = ap.parse_args()

    import bench as flagship
    from spacy_ray_b200.config import Config
    from spacy_ray_b200.train_cli import ray_train

    ns = argparse.Namespace(docs_per_gpu=args.docs_per_batch, width=args.width, depth=args.depth, min_len=8,
                            max_len=40, dropout=0.1)
    text = flagship.flagship_config(ns, 0)
    text += f"""
max_epochs = 0
eval_frequency = 100000
patience = 0

[training.batcher]
@batchers = "spacy.batch_by_sequence.v1"
size = {args.docs_per_batch}
"""
    cfg = Config().from_str(text, interpolate=False)
    cfg["training"]["max_steps"] = args.steps
    stats = ray_train(cfg, num_workers=args.workers, use_gpu=0 if args.gpu else -1, mode=args.mode,
                      quorum=args.quorum if args.mode == "async" else None, comm=args.comm, shard_data=True)
    total = sum(s["docs_per_sec"] for s in stats)
    used = sum(s["grads_used"] or 0 for s in stats)
    disc = sum(s["grads_discarded"] or 0 for s in stats)
    out = {
        "impl": f"rayproxy-emu ({args.mode}, quorum={args.quorum})" if args.mode == "async" else f"actors+{args.comm}",
        "label": "OUR emulation of the reference protocol - not a spacy-ray measurement",
        "metric": "docs/sec (wall clock, summed over workers) en tok2vec+NER", "value": total, "unit": "docs/s",
        "n_workers": args.workers, "device": "B200" if args.gpu else "cpu", "steps_per_worker": args.steps,
        "docs_per_batch": args.docs_per_batch,
        "percent_grads_used": (used / (used + disc)) if (used + disc) else None,
        "msgs_sent_per_worker": [s["msgs_sent"] for s in stats],
        "bytes_sent_per_worker": [s["bytes_sent"] for s in stats],
        "per_worker_docs_per_sec": [s["docs_per_sec"] for s in stats],
    }
    print(json.dumps(out))
    return 0


if __name__ == "__main__":
    sys.exit(main())
