#!/usr/bin/env python
"""Render benchmarks/exchange_trace.py output (per-rank JSON) as one markdown timeline.

    python scripts/render_trace.py gpurun_out/r2h_trace_8gpu profiles/r2_exchange_trace_8gpu.md
"""
import json
import sys
from pathlib import Path

prefix, out = sys.argv[1], sys.argv[2]
ranks = []
r = 0
while Path(f"{prefix}_rank{r}.json").exists():
    ranks.append(json.loads(Path(f"{prefix}_rank{r}.json").read_text()))
    r += 1
if not ranks:
    sys.exit(f"no {prefix}_rank*.json")
d0 = ranks[0]
W = d0["world"]
lines = [f"# Step timeline from %globaltimer stamps - {W} GPU(s), flagship tok2vec w256 d8 + NER, 1024 docs / GPU",
         "",
         f"`benchmarks/exchange_trace.py` ({d0['steps']} traced steps, median per stamp; microseconds after the rank's own step",
         "start).  Stamps: one-thread kernels on the main stream at step start / after the backward pass was enqueued-and-run /",
         "after the exchange joined, one after every parameter's gradient completed (`gradNN`), and the exchange kernels'",
         "own stamps per bucket: `signal` (grad-ready flag sent), `peers_seen` (all ranks' flags in), first reduce CTA start /",
         "last reduce CTA end, first update CTA start, `published` (last update CTA: weights stored everywhere + flag).",
         "Tracing adds ~50 stamp kernels and atomics to the step, so absolute times are slower than the benchmark's.",
         "", f"Buckets (keys in backward-completion order): " + "; ".join(
             f"b{i} = {len(ks)} keys" for i, ks in enumerate(d0["bucket_keys"])), ""]
# rank 0 full timeline with the backward stamps thinned to W-type keys
lines += ["## rank 0", "", "| us | event |", "|---:|---|"]
for k, v in d0["median_us"].items():
    if k.startswith("grad") and ".W@" not in k and ".E@" not in k:
        continue
    lines.append(f"| {v:.1f} | {k} |")
lines += ["", "## exchange kernels per rank (work items owned per bucket; us after own step start)", "",
          "| rank | owned items | " + " | ".join(f"b{b} signal / peers / reduce / update..published" for b in range(d0["buckets"])) +
          " | backward done | step end |", "|---:|---|" + "---|" * d0["buckets"] + "---:|---:|"]
for d in ranks:
    m = d["median_us"]
    cells = []
    for b in range(d["buckets"]):
        g = lambda n: m.get(f"b{b}.{n}")
        if g("reduce_start") is None:
            cells.append(f"{g('signal'):.0f} / - (owns nothing)" if g("signal") is not None else "-")
        else:
            cells.append(f"{g('signal'):.0f} / {g('peers_seen'):.0f} / {g('reduce_start'):.0f}-{g('reduce_end'):.0f} / "
                         f"{g('update_start'):.0f}-{g('published'):.0f}")
    lines.append(f"| {d['rank']} | {d['owned_per_bucket']} | " + " | ".join(cells) +
                 f" | {m.get('backward_enqueued_done', 0):.0f} | {m.get('step_end', 0):.0f} |")
lines += ["", "Reading: every bucket but the last is reduced, updated and published while the backward pass is still",
          "running (its `published` stamp precedes `backward done`); only the embedding-table bucket, whose gradients are the",
          "last thing the backward pass produces, is exposed.  There is no NCCL kernel in the step."]
Path(out).write_text("\n".join(lines) + "\n")
print("\n".join(lines[:60]))
