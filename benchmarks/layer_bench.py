"""One tok2vec encoder layer forward (window GEMM + maxout + LayerNorm + dropout + residual) at the
flagship shape, fused epilogue vs the two-kernel path, CUDA events, inputs rotated through > L2.

    python benchmarks/layer_bench.py [--rows 25683] [--width 256] [--iters 200]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=25683)
    ap.add_argument("--width", type=int, default=256)
    ap.add_argument("--iters", type=int, default=200)
    ap.add_argument("--sets", type=int, default=8)
    args = ap.parse_args()
    from spacy_ray_b200.ops.b200_ops import B200Ops

    dev = "cuda:0"
    T, w = args.rows, args.width
    g = torch.Generator(device=dev).manual_seed(0)
    mask = (torch.rand(T, 1, device=dev, generator=g) > 0.04).float()
    Xs = [(torch.randn(T, w, device=dev, generator=g) * mask).bfloat16() for _ in range(args.sets)]
    W = (torch.randn(w, 3, 3 * w, device=dev, generator=g) * 0.05).bfloat16()
    b = (torch.randn(w, 3, device=dev, generator=g) * 0.1).bfloat16()
    G = torch.ones(w, device=dev).bfloat16()
    beta = torch.zeros(w, device=dev).bfloat16()
    out = {"rows": T, "width": w}
    for name, fused in (("fused", True), ("two_kernel", False)):
        ops = B200Ops(dev)
        ops.fused_ln = fused

        def fn(i):
            return ops.maxout_block(Xs[i % len(Xs)], W, b, G, beta, mask, window=1, residual=True, dropout=0.1,
                                    is_train=True, seed=3)

        for i in range(10):
            fn(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.iters):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        out[name + "_us"] = round(e0.elapsed_time(e1) * 1000.0 / args.iters, 2)
        if fused:
            ops.check_fused_ln()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
