#!/usr/bin/env python
"""Per-kernel SASS mnemonic counts of the built extension:
    cuobjdump -sass spacy_ray_b200/ops/_srb_cuda.so | python scripts/sass_summary.py > profiles/...txt
"""
import collections
import re
import subprocess
import sys

KEYS = ("UTCHMMA", "UTMALDG", "UTCBAR", "UTCATOMSWS", "LDTM", "UBLKCP", "SYNCS", "MULTIMEM", "RED.E", "ATOMG", "ATOM.E",
        "LDG.E.128", "LDG.E.64", "STG.E.128", "STG.E.64", "LDS.128", "LDS.64", "SHFL", "MUFU", "HMMA", "FFMA", "MEMBAR",
        "UCGABAR", "ERRBAR", "LDGMC", "STG.E.64.STRONG.SYS", "LDG.E.128.STRONG.SYS")


def main() -> None:
    fn = None
    counts = collections.OrderedDict()
    for line in sys.stdin:
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = m.group(1)
            counts[fn] = collections.Counter()
            continue
        if fn is None:
            continue
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        counts[fn]["_total"] += 1
        for k in KEYS:
            if op.startswith(k):
                counts[fn][k] += 1
                if k == "UTCHMMA" and ".2CTA" in op:
                    counts[fn]["UTCHMMA.2CTA"] += 1
    names = list(counts)
    dem = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    print("# SASS mnemonic counts per kernel of spacy_ray_b200/ops/_srb_cuda.so (sm_100a), from `cuobjdump -sass`")
    print("# UTCHMMA = tcgen05.mma (.2CTA = cta_group::2), UTMALDG = TMA tensor load, UTCBAR = tcgen05.commit, LDTM = tcgen05.ld,")
    print("# UTCATOMSWS = tcgen05.alloc/dealloc, SYNCS = mbarrier ops, LDGMC = NVLS multimem.ld_reduce (multimem.st compiles to STG...STRONG.SYS on the multicast address), *.STRONG.SYS = system-scope peer loads/stores, RED/ATOMG = global reductions")
    for n, d in zip(names, dem):
        c = counts[n]
        short = re.sub(r"\(.*", "", d)[:120]
        items = ", ".join(f"{k}={v}" for k, v in c.items() if k != "_total" and v)
        print(f"{short}\n    instr={c['_total']}  {items}")


if __name__ == "__main__":
    main()
