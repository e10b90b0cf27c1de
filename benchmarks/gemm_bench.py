#!/usr/bin/env python
"""Kernel-level roofline check of the tcgen05 GEMM family at the flagship shapes.

Times (CUDA events, L2 flushed between iterations) the three GEMMs of one MaxoutWindowEncoder
layer - forward (window + bias + maxout epilogue), dX (weights as stored, window + residual
epilogue), dW (MN-major operands, split-K fp32 red) - next to cuBLAS (torch.matmul) on the
materialised operands the library path needs, and prints TFLOP/s and the fraction of the
measured cuBLAS bf16 peak (MEASURED_PEAKS.json).

    python benchmarks/gemm_bench.py [--rows 25683] [--width 256] [--json out.json]
"""
from __future__ import annotations

import argparse
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=25683)
    ap.add_argument("--width", type=int, default=256)
    ap.add_argument("--iters", type=int, default=30)
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    from spacy_ray_b200.ops.b200_ops import EPI_MAXOUT3, EPI_STORE, MODE_KK, MODE_KMN, B200Ops

    dev = torch.device("cuda:0")
    ops = B200Ops(dev)
    T, w = args.rows, args.width
    N = 3 * w                                   # nO * nP
    peaks = {}
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        peaks = json.loads(p.read_text())
    peak_tf = float(peaks.get("bf16_tflops_sustained", 1386.0))
    torch.manual_seed(0)
    X = torch.randn(T, w, device=dev).bfloat16()
    W2 = (torch.randn(N, 3 * w, device=dev) * 0.05).bfloat16()
    bias = torch.randn(N, device=dev).bfloat16()
    dZ = torch.randn(T, N, device=dev).bfloat16()
    dY = torch.randn(T, w, device=dev).bfloat16()
    mask = torch.ones(T, device=dev)
    H = torch.empty(T, w, device=dev, dtype=torch.bfloat16)
    which = torch.empty(T, w, device=dev, dtype=torch.uint8)
    dX = torch.empty(T, w, device=dev, dtype=torch.bfloat16)
    dW = torch.zeros(N, 3 * w, device=dev, dtype=torch.float32)
    Xw = torch.cat([X, X, X], dim=1).contiguous()           # what cuBLAS needs: the materialised window
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    flops = 2.0 * T * N * 3 * w

    def timeit(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        tot = 0.0
        for _ in range(args.iters):
            flush.zero_()                                    # > L2 (126 MB)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        return tot / args.iters * 1e3                         # us

    rows = []

    def report(name, us):
        tf = flops / us / 1e6
        rows.append({"kernel": name, "us": round(us, 2), "tflops": round(tf, 1), "frac_of_cublas_peak": round(tf / peak_tf, 3)})
        print(f"{name:58s} {us:8.1f} us  {tf:7.1f} TFLOP/s  {100 * tf / peak_tf:5.1f}% of cuBLAS sustained peak")

    for cl in (1, 2, 3):
        report(f"fwd  window+bias+maxout  bn=192 cluster={cl}", timeit(lambda: ops.tc_gemm(
            X, W2, H, mode=MODE_KK, epi=EPI_MAXOUT3, block_n=192, M=T, N=N, K=w, a_row_shift=(-1, 0, 1),
            a_col_off=(0, 0, 0), b_row_off=(0, 0, 0), b_col_off=(0, w, 2 * w), bias=bias, which=which, cluster=cl)))
    report("fwd  cuBLAS (T,3w)@(3w,N) on a materialised window", timeit(lambda: torch.matmul(Xw, W2.t())))
    for bn in (256, 128):
        for cl in (1, 2, 3):
            report(f"dX   weights-as-stored window+residual bn={bn} cluster={cl}", timeit(lambda: ops.tc_gemm(
                dZ, W2, dX, mode=MODE_KMN, epi=EPI_STORE, block_n=bn, M=T, N=w, K=N, a_row_shift=(1, 0, -1),
                a_col_off=(0, 0, 0), b_row_off=(0, 0, 0), b_col_off=(0, w, 2 * w), add_src=dY, row_scale=mask,
                cluster=cl)))
    report("dX   cuBLAS (T,N)@(N,3w) (window still to be folded)", timeit(lambda: torch.matmul(dZ, W2)))
    for cl in (1, 2, 3):
        ops.gemm_cluster = cl
        report(f"dW   MN-major split-K fp32 red cluster={cl}", timeit(lambda: ops._dw_tc(dZ, X, 1, out=dW)))
    report("dW   cuBLAS (N,T)@(T,3w) fp32 out", timeit(lambda: torch.mm(dZ.t(), Xw, out_dtype=torch.float32)))
    out = {"rows": T, "width": w, "flops_per_gemm": flops, "cublas_peak_tflops": peak_tf, "results": rows}
    if args.json:
        Path(args.json).write_text(json.dumps(out, indent=1))
    return 0


if __name__ == "__main__":
    sys.exit(main())
