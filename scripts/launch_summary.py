#!/usr/bin/env python
"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` log by kernel."""
import collections, csv, re, sys
path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/launches.csv"
lines = [l for l in open(path) if not l.startswith("==")]
rows = list(csv.DictReader(lines))
agg = collections.defaultdict(lambda: [0, 0.0])
for row in rows:
    name = re.sub(r"\(.*", "", re.sub(r"<.*", "", row["Kernel Name"])).replace("void ", "")[:60]
    if "gemm_kernel" in row["Kernel Name"]:
        m = re.search(r"gemm_kernel<([^>]*)>", row["Kernel Name"])
        name = "srb::gemm_kernel<%s>" % (m.group(1) if m else "")
    v = float(row["Metric Value"].replace(",", "")); u = row["Metric Unit"]
    v = v / 1000 if u == "ns" else (v * 1000 if u == "ms" else v)
    agg[name][0] += 1; agg[name][1] += v
tot = sum(v[1] for v in agg.values())
print(f"{len(rows)} launches, total {tot:.0f} us")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{v[1]:10.1f} us {v[0]:5d} {v[1]/v[0]:8.1f}/call {v[1]/tot*100:5.1f}%  {k}")
