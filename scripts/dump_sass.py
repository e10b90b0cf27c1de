#!/usr/bin/env python
"""Write the raw SASS of every hot kernel of the built extension to profiles/sass/<kernel>.sass
(one file per instantiation that the flagship / parser / tagger steps actually launch) plus an index
with the mnemonics that prove the Blackwell paths (UTC*MMA = tcgen05.mma, LDTM = tcgen05.ld,
UTMALDG = TMA, LDGMC/multimem = NVLS, .SYS = peer-memory flags).  Runs without a GPU:

    python scripts/dump_sass.py [--all]
"""
import argparse
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "spacy_ray_b200" / "ops" / "_srb_cuda.so"
OUT = ROOT / "profiles" / "sass"

# demangled-name substrings of the kernels on the measured paths (template args as cuobjdump prints them)
HOT = [
    ("gemm_fwd_window_maxout_pair_halo", "gemm_kernelILi192ELi0ELi1ELi2ELb1ELb1ELb0E"),
    ("gemm_fwd_window_maxout_pair_halo_resident_b", "gemm_kernelILi192ELi0ELi1ELi2ELb1ELb1ELb1E"),
    ("gemm_fwd_window_maxout_ln_pair_halo", "gemm_kernelILi192ELi0ELi3ELi2ELb1ELb1ELb0E"),
    ("gemm_fwd_plain_maxout_pair", "gemm_kernelILi192ELi0ELi1ELi2ELb1ELb0ELb0E"),
    ("gemm_dx_window_pair_halo", "gemm_kernelILi128ELi2ELi0ELi2ELb1ELb1E"),
    ("gemm_dx_plain_pair", "gemm_kernelILi256ELi2ELi0ELi2ELb1ELb0E"),
    ("gemm_dw_splitk_pair", "gemm_kernelILi256ELi1ELi2ELi2ELb1ELb0E"),
    ("gemm_linear_store_pair", "gemm_kernelILi192ELi0ELi0ELi2ELb1ELb0E"),
    ("gemm_linear_n64", "gemm_kernelILi64ELi0ELi0ELi1ELb0ELb0E"),
    ("maxout_ln_fwd_vec", "maxout_ln_fwd_vec_kernelILi1ELi8ELi4ELi32ELb1E"),
    ("maxout_ln_fwd_vec_w96", "maxout_ln_fwd_vec_kernelILi1ELi8ELi4ELi16ELb0E"),
    ("maxout_ln_bwd_vec", "maxout_ln_bwd_vec_kernelILi3ELi8ELi2ELi32ELb1E"),
    ("maxout_ln_bwd_vec_w96", "maxout_ln_bwd_vec_kernelILi3ELi8ELi2ELi16ELb0E"),
    ("hash_embed_fwd", "hash_embed_fwd_kernel"),
    ("hash_embed_bwd_sorted_i32", "hash_embed_bwd_sorted_kernelIiE"),
    ("biluo_block", "biluo_block_kernel"),
    ("biluo_steps", "biluo_steps_kernelILi3ELi4E"),
    ("arc_eager_steps", "arc_eager_steps_kernel"),
    ("transition_scatter", "transition_scatter_kernel"),
    ("linear_softmax_xent", "linear_softmax_xent_kernel"),
    ("softmax_xent_bias", "softmax_xent_bias_kernel"),
    ("arena_init", "arena_init_kernel"),
    ("bucket_signal", "bucket_signal_kernel"),
    ("bucket_wait", "bucket_wait_kernel"),
    ("bucket_reduce", "bucket_reduce_kernel"),
    ("bucket_update", "bucket_update_kernel"),
    ("bucket_gate_zero", "bucket_gate_zero_kernel"),
    ("gate_wait", "gate_wait_kernel"),
    ("p2p_reduce_scatter", "p2p_reduce_scatter_kernel"),
    ("p2p_all_gather", "p2p_all_gather_kernel"),
    ("colsum_bf16", "colsum_bf16_kernel"),
    ("f32_to_bf16_zero", "f32_to_bf16_zero_kernel"),
]
# matched on whole dot-separated fields of the opcode (so HMMA does not match UTCHMMA)
PROOF = ["UTCHMMA", "UTCHMMA.2CTA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "LDGMC", "REDG",
         "RED", "ATOMG", "MEMBAR.SC.SYS", "MEMBAR.ALL.SYS", "LDG.STRONG.SYS", "STG.STRONG.SYS", "STG.128.STRONG.SYS",
         "HMMA", "ELECT", "UCGABAR", "PREEXIT", "ACQBULK"]


def _has(op: str, proof: str) -> bool:
    """True if the opcode starts with the proof's first field and carries its other fields."""
    of, pf = op.split("."), proof.split(".")
    return of[0] == pf[0] and all(f in of[1:] for f in pf[1:])


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--all", action="store_true", help="one file per function in the library, not only the hot ones")
    args = ap.parse_args()
    if not LIB.exists():
        print(f"{LIB} not built", file=sys.stderr)
        return 1
    sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    funcs = {}
    cur, buf = None, []
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            if cur:
                funcs[cur] = buf
            cur, buf = m.group(1), [line]
        elif cur is not None:
            buf.append(line)
    if cur:
        funcs[cur] = buf
    OUT.mkdir(parents=True, exist_ok=True)
    for old in OUT.glob("*.sass"):
        old.unlink()
    index = [f"# SASS listings (cuobjdump -sass {LIB.relative_to(ROOT)}; sm_100a)", "",
             "| file | function | instructions | proof mnemonics |", "|---|---|---:|---|"]
    wanted = [(short, pat) for short, pat in HOT]
    if args.all:
        wanted = [(re.sub(r"[^A-Za-z0-9_]", "_", n)[:120], n) for n in funcs]
    for short, pat in wanted:
        hits = [n for n in funcs if pat in n]
        if not hits:
            index.append(f"| - | `{pat}` | 0 | NOT FOUND in this build |")
            continue
        name = sorted(hits, key=len)[0]
        body = funcs[name]
        (OUT / f"{short}.sass").write_text("\n".join(body) + "\n")
        ops = collections.Counter()
        n_ins = 0
        for line in body:
            m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]+)", line)
            if m:
                n_ins += 1
                op = m.group(1)
                for p in PROOF:
                    if _has(op, p):
                        ops[p] += 1
        proof = ", ".join(f"{k} x{v}" for k, v in sorted(ops.items(), key=lambda kv: -kv[1]) if v) or "-"
        index.append(f"| `sass/{short}.sass` | `{name[:90]}` | {n_ins} | {proof} |")
    index += ["", "Reading the mnemonics: `UTCHMMA(.2CTA)` = `tcgen05.mma` (`cta_group::2`), `LDTM` = `tcgen05.ld`,",
              "`UTCBAR` = `tcgen05.commit`, `UTMALDG` = TMA tensor load, `SYNCS` = mbarrier ops, `UCGABAR` = cluster barrier,",
              "`LDGMC` = `multimem.ld_reduce` (NVLS), `STG.128.STRONG.SYS` on a multicast address = `multimem.st`,",
              "`LDG/STG.STRONG.SYS` = acquire/release flag traffic over NVLink peer memory, `MEMBAR.ALL.SYS` =",
              "`fence.acq_rel.sys`, `PREEXIT` / `ACQBULK` = `griddepcontrol.launch_dependents` / `.wait`.  No `HMMA`",
              "(legacy `mma.sync`) appears in any kernel."]
    (ROOT / "profiles" / "sass_index.md").write_text("\n".join(index) + "\n")
    print("\n".join(index))
    return 0


if __name__ == "__main__":
    sys.exit(main())
