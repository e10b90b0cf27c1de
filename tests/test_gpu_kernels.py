"""Numerics of every sm_100a kernel against the plain-PyTorch fp32 reference
(``TorchOps``) of the same op.  Run on a B200: ``pytest -m gpu``."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from spacy_ray_b200.ops.b200_ops import B200Ops

    return B200Ops("cuda:0")


@pytest.fixture(scope="module")
def ref():
    from spacy_ray_b200.ops.torch_ops import TorchOps

    return TorchOps("cuda:0", dtype=torch.float32)


def _padded_batch(lens, w, dev="cuda:0", seed=0):
    g = torch.Generator().manual_seed(seed)
    rows = sum(lens) + len(lens) + 1
    mask = torch.zeros(rows, 1)
    r = 1
    for n in lens:
        mask[r:r + n] = 1
        r += n + 1
    X = torch.randn(rows, w, generator=g) * mask
    return X.to(dev), mask.to(dev)


def _close(a, b, rtol, atol, name=""):
    a, b = a.float(), b.float()
    err = (a - b).abs()
    tol = atol + rtol * b.abs()
    bad = int((err > tol).sum().item())
    assert bad == 0, f"{name}: {bad} of {err.numel()} elements out of tolerance, max err {err.max().item():.4g}"


# ---------------------------------------------------------------------------- K1
@pytest.mark.parametrize("w", [64, 96, 256, 16])
def test_hash_embed_fwd_bwd(ops, ref, w):
    torch.manual_seed(0)
    Tp = 301
    attrs = torch.randint(-(1 << 62), 1 << 62, (Tp, 4), dtype=torch.int64, device="cuda")
    attrs[:, 0] = attrs[torch.randint(0, 20, (Tp,), device="cuda"), 0]      # heavy collisions
    mask = (torch.rand(Tp, 1, device="cuda") > 0.1).float()
    rows = [5000, 1000, 2500, 2500]
    tables = [(torch.randn(n, w, device="cuda") * 0.1).bfloat16() for n in rows]
    seeds, cols = [8, 9, 10, 11], [0, 1, 2, 3]
    Y = ops.multi_hash_embed(attrs, mask, tables, seeds, cols)
    Yr = ref.multi_hash_embed(attrs, mask, [t.float() for t in tables], seeds, cols)
    _close(Y, Yr, 1e-2, 1e-2, "hash_embed_fwd")
    assert float((Y.float() * (1 - mask)).abs().sum()) == 0.0
    dY = torch.randn(Tp, 4 * w, device="cuda").bfloat16()
    g = ops.multi_hash_embed_backward(dY, attrs, mask, rows, seeds, cols)
    gr = ref.multi_hash_embed_backward(dY.float(), attrs, mask, rows, seeds, cols)
    for a, b in zip(g, gr):
        _close(a, b, 1e-3, 1e-3, "hash_embed_bwd")
    # rows pre-grouped on the host (int32, what engine.Trainer ships with each batch); masked rows
    # may appear anywhere (the kernel skips them) and the tail repeats row 0
    host = attrs.cpu().numpy()
    perm = np.stack([np.argsort(host[:, c], kind="stable") for c in range(4)]).astype(np.int32)
    g2 = ops.multi_hash_embed_backward(dY, attrs, mask, rows, seeds, cols, perm=torch.from_numpy(perm).cuda())
    for a, b in zip(g2, gr):
        _close(a, b, 1e-3, 1e-3, "hash_embed_bwd(host perm)")
    # a permuted table<->column assignment must pick the matching perm rows
    cols2 = [3, 2, 1, 0]
    g3 = ops.multi_hash_embed_backward(dY, attrs, mask, rows, seeds, cols2, perm=torch.from_numpy(perm).cuda())
    gr3 = ref.multi_hash_embed_backward(dY.float(), attrs, mask, rows, seeds, cols2)
    for a, b in zip(g3, gr3):
        _close(a, b, 1e-3, 1e-3, "hash_embed_bwd(host perm, permuted columns)")


# ---------------------------------------------------------------------------- tcgen05 GEMMs
@pytest.mark.parametrize("M,N,K,bn", [(128, 128, 64, 128), (300, 256, 256, 256), (1000, 768, 192, 192),
                                      (257, 64, 128, 64), (4096, 384, 64, 128)])
@pytest.mark.parametrize("cluster", [1, 2, 3])
def test_tc_gemm_plain_nt(ops, M, N, K, bn, cluster):
    torch.manual_seed(1)
    A = torch.randn(M, K, device="cuda").bfloat16()
    B = torch.randn(N, K, device="cuda").bfloat16()
    bias = torch.randn(N, device="cuda").bfloat16()
    out = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
    ops.tc_gemm(A, B, out, mode=0, epi=0, block_n=bn, M=M, N=N, K=K, bias=bias, cluster=cluster)
    torch.cuda.synchronize()
    want = A.float() @ B.float().t() + bias.float()
    _close(out, want, 2e-2, 2e-2 * math.sqrt(K), f"tc_gemm {M}x{N}x{K}")


@pytest.mark.parametrize("w,lens", [(64, (5, 1, 9, 30)), (256, tuple(range(3, 40)))])
def test_tc_window_maxout_epilogue(ops, ref, w, lens):
    torch.manual_seed(2)
    X, mask = _padded_batch(lens, w)
    Xb = X.bfloat16()
    nO, nP = w, 3
    W = (torch.randn(nO, nP, 3 * w, device="cuda") * 0.1).bfloat16()
    b = (torch.randn(nO, nP, device="cuda") * 0.1).bfloat16()
    Tp = X.shape[0]
    H = torch.empty(Tp, nO, device="cuda", dtype=torch.bfloat16)
    which = torch.empty(Tp, nO, device="cuda", dtype=torch.uint8)
    ops.tc_gemm(Xb, W.reshape(nO * nP, 3 * w), H, mode=0, epi=1, block_n=192, M=Tp, N=nO * nP, K=w,
                a_row_shift=(-1, 0, 1), a_col_off=(0, 0, 0), b_row_off=(0, 0, 0), b_col_off=(0, w, 2 * w),
                bias=b.reshape(-1), which=which)
    torch.cuda.synchronize()
    Xw = ref.seq2col(Xb.float(), 1)
    Z = (Xw @ W.float().reshape(nO * nP, 3 * w).t() + b.float().reshape(-1)).view(Tp, nO, nP)
    Href, wref = Z.max(dim=2)
    _close(H, Href, 2e-2, 5e-2, "window maxout H")
    # argmax may legitimately differ where two pieces are within rounding of each other
    top2 = Z.topk(2, dim=2).values
    clear = (top2[..., 0] - top2[..., 1]) > 0.05
    assert (which.long()[clear] == wref[clear]).float().mean().item() > 0.999


def test_tc_window_dx_with_residual(ops, ref):
    torch.manual_seed(3)
    w, N = 128, 384
    dZ, mask = _padded_batch((7, 2, 33, 12), N)
    dZb = dZ.bfloat16()
    Tp = dZ.shape[0]
    W2 = (torch.randn(N, 3 * w, device="cuda") * 0.1).bfloat16()
    dY = (torch.randn(Tp, w, device="cuda")).bfloat16()
    WT = W2.t().contiguous()
    dX = torch.empty(Tp, w, device="cuda", dtype=torch.bfloat16)
    ops.tc_gemm(dZb, WT, dX, mode=0, epi=0, block_n=128, M=Tp, N=w, K=N, a_row_shift=(1, 0, -1),
                a_col_off=(0, 0, 0), b_row_off=(0, w, 2 * w), b_col_off=(0, 0, 0), add_src=dY,
                row_scale=mask.reshape(-1))
    torch.cuda.synchronize()
    dXw = dZb.float() @ W2.float()
    want = ref.backprop_seq2col(dXw, 1) + dY.float() * mask
    _close(dX, want, 2e-2, 0.1, "window dX")


@pytest.mark.parametrize("M,N,K,bn", [(128, 128, 64, 128), (300, 256, 768, 256), (1000, 192, 192, 192),
                                      (257, 64, 128, 64), (4096, 512, 64, 256)])
@pytest.mark.parametrize("cluster", [1, 2, 3])
def test_tc_gemm_nn_weights_as_stored(ops, M, N, K, bn, cluster):
    """MODE_KMN: A (M,K) K-major x B (K,N) row-major (MN-major UMMA operand) - dX = dY @ W without W^T."""
    torch.manual_seed(11)
    A = torch.randn(M, K, device="cuda").bfloat16()
    B = torch.randn(K, N, device="cuda").bfloat16()
    out = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
    ops.tc_gemm(A, B, out, mode=2, epi=0, block_n=bn, M=M, N=N, K=K, cluster=cluster)
    torch.cuda.synchronize()
    _close(out, A.float() @ B.float(), 2e-2, 2e-2 * math.sqrt(K), f"tc_gemm NN {M}x{N}x{K}")


@pytest.mark.parametrize("cluster", [1, 2, 3])
def test_tc_window_dx_nn_with_residual(ops, ref, cluster):
    torch.manual_seed(3)
    w, N = 256, 768
    dZ, mask = _padded_batch((7, 2, 33, 12, 40, 1), N)
    dZb = dZ.bfloat16()
    Tp = dZ.shape[0]
    W2 = (torch.randn(N, 3 * w, device="cuda") * 0.1).bfloat16()
    dY = (torch.randn(Tp, w, device="cuda")).bfloat16()
    dX = torch.empty(Tp, w, device="cuda", dtype=torch.bfloat16)
    ops.tc_gemm(dZb, W2, dX, mode=2, epi=0, block_n=256, M=Tp, N=w, K=N, a_row_shift=(1, 0, -1),
                a_col_off=(0, 0, 0), b_row_off=(0, 0, 0), b_col_off=(0, w, 2 * w), add_src=dY,
                row_scale=mask.reshape(-1), cluster=cluster)
    torch.cuda.synchronize()
    dXw = dZb.float() @ W2.float()
    want = ref.backprop_seq2col(dXw, 1) + dY.float() * mask
    _close(dX, want, 2e-2, 0.15, "window dX (NN)")


@pytest.mark.parametrize("T,C", [(1, 64), (777, 96), (25000, 64), (5000, 768), (333, 2048)])
def test_colsum_kernel(ops, T, C):
    torch.manual_seed(5)
    X = torch.randn(T, C, device="cuda").bfloat16()
    got = ops.colsum(X)
    launched = ops.launches
    _close(got, X.float().sum(0), 1e-3, 1e-3 * math.sqrt(T), "colsum")
    acc = torch.ones(C, device="cuda")
    ops.colsum(X[:, : C // 2 // 8 * 8 or 8], out=acc)        # strided view, accumulates into `out`
    assert ops.launches == launched + 1
    Ch = C // 2 // 8 * 8 or 8
    _close(acc[:Ch], 1 + X[:, :Ch].float().sum(0), 1e-3, 1e-3 * math.sqrt(T), "colsum acc")
    assert float((acc[Ch:] - 1).abs().max()) == 0.0 if Ch < C else True


@pytest.mark.parametrize("window", [0, 1])
@pytest.mark.parametrize("cluster", [1, 2, 3])
@pytest.mark.parametrize("w,N", [(128, 384), (64, 64), (256, 128), (64, 192)])
def test_tc_dw_mn_major_split_k(ops, ref, window, cluster, w, N):
    ops = type(ops)("cuda:0")
    ops.gemm_cluster = cluster
    torch.manual_seed(4)
    X, mask = _padded_batch(tuple(range(2, 60)), w)
    Xb = X.bfloat16()
    Tp = X.shape[0]
    dZ = (torch.randn(Tp, N, device="cuda") * mask).bfloat16()
    Kt = w * (3 if window else 1)
    out = torch.zeros(N, Kt, device="cuda", dtype=torch.float32)
    got = ops._dw_tc(dZ, Xb, window, out)
    torch.cuda.synchronize()
    assert got is not None
    Xw = ref.seq2col(Xb.float(), 1) if window else Xb.float()
    want = dZ.float().t() @ Xw
    _close(out, want, 2e-2, 2e-2 * math.sqrt(Tp), f"dW window={window}")


# ---------------------------------------------------------------------------- fused blocks
@pytest.mark.parametrize("use_tc", [False, True])
@pytest.mark.parametrize("window,residual,w", [(0, False, 64), (1, True, 64), (1, True, 256), (1, True, 96), (0, False, 96),
                                                (1, True, 160)])
def test_maxout_block_fwd_bwd(ref, use_tc, window, residual, w):
    from spacy_ray_b200.ops.b200_ops import B200Ops

    ops = B200Ops("cuda:0", use_tc=use_tc)
    torch.manual_seed(5)
    nI = w if window else 4 * w
    X, mask = _padded_batch((4, 9, 1, 17, 30), nI if not window else w)
    nO, nP = w, 3
    W = (torch.randn(nO, nP, nI * (3 if window else 1), device="cuda") * 0.08).bfloat16()
    b = (torch.randn(nO, nP, device="cuda") * 0.1).bfloat16()
    G = (torch.rand(nO, device="cuda") + 0.5).bfloat16()
    beta = (torch.randn(nO, device="cuda") * 0.1).bfloat16()
    if residual and X.shape[1] != nO:
        pytest.skip("residual needs nI == nO")
    Xb = X.bfloat16()
    Y, ctx = ops.maxout_block(Xb, W, b, G, beta, mask, window=window, residual=residual, dropout=0.2,
                              is_train=True, seed=11)
    Yr, cr = ref.maxout_block(Xb.float(), W.float(), b.float(), G.float(), beta.float(), mask, window=window,
                              residual=residual, dropout=0.2, is_train=True, seed=11)
    _close(Y, Yr, 3e-2, 6e-2, "maxout_block fwd")
    assert float((Y.float() * (1 - mask)).abs().sum()) == 0.0
    dY = torch.randn_like(Yr).bfloat16()
    dX, dW, db, dG, dbeta = ops.maxout_block_backward(dY, ctx)
    # the reference backward must route through the same winners to be comparable
    cr["which"] = ctx["which"]
    dXr, dWr, dbr, dGr, dbetar = ref.maxout_block_backward(dY.float(), cr)
    T = X.shape[0]
    _close(dX, dXr, 5e-2, 0.15, "dX")
    _close(dW, dWr, 5e-2, 0.05 * math.sqrt(T), "dW")
    _close(db, dbr.reshape(-1).view_as(db), 5e-2, 0.05 * math.sqrt(T), "db")
    _close(dG, dGr, 5e-2, 0.05 * math.sqrt(T), "dG")
    _close(dbeta, dbetar, 5e-2, 0.05 * math.sqrt(T), "dbeta")


@pytest.mark.parametrize("w,window,rows", [(256, 1, 25683), (256, 0, 3000), (128, 1, 20000), (512, 1, 9000), (64, 1, 700)])
def test_fused_layernorm_epilogue_matches_the_two_kernel_path(w, window, rows):
    """EPI_MAXOUT3_LN (LayerNorm statistics exchanged between the N tiles of a row, activations kept in
    the epilogue's registers) against GEMM+maxout followed by the LayerNorm kernel: same winners, same
    dropout mask, outputs within one bf16 ulp of rounding; run twice (the counters re-arm themselves)."""
    from spacy_ray_b200.ops.b200_ops import B200Ops

    fused, plain = B200Ops("cuda:0"), B200Ops("cuda:0")
    fused.fused_ln, plain.fused_ln = True, False
    assert fused.gemm_cluster == 3
    g = torch.Generator(device="cuda").manual_seed(3)
    mask = (torch.rand(rows, 1, device="cuda", generator=g) > 0.05).float()
    nI = w if window else 2 * w
    X = (torch.randn(rows, nI, device="cuda", generator=g) * mask).bfloat16()
    W = (torch.randn(w, 3, nI * (3 if window else 1), device="cuda", generator=g) * 0.05).bfloat16()
    b = (torch.randn(w, 3, device="cuda", generator=g) * 0.1).bfloat16()
    G = (torch.rand(w, device="cuda", generator=g) + 0.5).bfloat16()
    beta = (torch.randn(w, device="cuda", generator=g) * 0.1).bfloat16()
    for rep in range(2):
        kw = dict(window=window, residual=bool(window), dropout=0.1, is_train=True, seed=5 + rep)
        Yf, cf = fused.maxout_block(X, W, b, G, beta, mask, **kw)
        Yp, cp = plain.maxout_block(X, W, b, G, beta, mask, **kw)
        torch.cuda.synchronize()
        fused.check_fused_ln()
        assert fused.launches == rep + 1 and plain.launches == 2 * (rep + 1), "one kernel per layer on the fused path"
        live = mask[:, 0] != 0            # pad rows: the fused epilogue writes 0, the plain GEMM epilogue the (unused) winner
        assert torch.equal(cf["which"][live], cp["which"][live]), "different maxout winners"
        assert int(cf["which"][~live].sum().item()) == 0
        _close(cf["rstd"], cp["rstd"], 1e-4, 1e-6, "rstd")
        _close(cf["xhat"], cp["xhat"], 1e-2, 1e-3, "xhat")
        _close(Yf, Yp, 1e-2, 2e-2, "Y")
        assert float((Yf.float() * (1 - mask)).abs().sum()) == 0.0
        assert float(cf["xhat"].float()[mask[:, 0] == 0].abs().sum()) == 0.0
    scratch = [v for k, v in fused._ws.items() if k[0] == "ln_scratch"]
    assert scratch and all(v[1].tolist() == [3, 0, 0] for v in scratch), "launch tag / CTA counter not maintained"


@pytest.mark.parametrize("tc_head", [True, False])
@pytest.mark.parametrize("T,w,nC", [(500, 64, 17), (3001, 96, 50), (2000, 512, 100), (700, 256, 64)])
def test_softmax_xent(ref, tc_head, T, w, nC):
    """Tagger head: tcgen05 logits GEMM + bias/softmax/CE kernel (default) and the one-kernel CUDA-core
    version, both against the fp32 reference; called twice (the logits scratch must come back zeroed)."""
    from spacy_ray_b200.ops.b200_ops import B200Ops

    ops = B200Ops("cuda:0")
    ops.tag_head_tc = tc_head
    torch.manual_seed(6)
    W = (torch.randn(nC, w, device="cuda") * 0.2).bfloat16()
    b = (torch.randn(nC, device="cuda") * 0.1).bfloat16()
    for rep in range(2):
        X = torch.randn(T, w, device="cuda").bfloat16()
        labels = torch.randint(-1, nC, (T,), device="cuda")
        loss, d, guesses, dX, dW, db = ops.softmax_xent(X, W, b, labels)
        lr, dr, gr, dXr, dWr, dbr = ref.softmax_xent(X.float(), W.float(), b.float(), labels)
        _close(d, dr, 2e-2, 1e-2, "d_logits")
        assert abs(float(loss) - float(lr)) / float(lr) < 2e-2
        assert (guesses == gr).float().mean().item() > 0.98
        _close(dW, dWr, 3e-2, 0.02 * math.sqrt(T), "dW")
        _close(dX, dXr, 3e-2, 3e-2, "dX")
    if tc_head:
        ws = [v for k, v in ops._ws.items() if str(k[0]).startswith("tag_logits")]
        assert ws and float(ws[0].abs().sum()) == 0.0, "logits scratch not left zeroed"


def test_adam_shard_matches_reference(ops, ref):
    torch.manual_seed(7)
    lens = [128 * 3, 4096 + 128, 128, 128 * 70]
    offs = [0]
    for n in lens[:-1]:
        offs.append(offs[-1] + n)
    total = sum(lens)
    g = torch.randn(total, device="cuda") * 3
    w = torch.randn(total, device="cuda")
    m1 = torch.zeros(total, device="cuda")
    m2 = torch.zeros(total, device="cuda")
    w_bf = torch.zeros(total, device="cuda", dtype=torch.bfloat16)
    blk_key, blk_off = [], []
    for k, n in enumerate(lens):
        for c in range((n + 4095) // 4096):
            blk_key.append(k)
            blk_off.append(c)
    dev = "cuda"
    hyper = torch.tensor([0.01, 0.9, 0.999, 1e-8, 1.0, 0.01, 1.0, 1.0], device=dev)
    step = torch.zeros(1, dtype=torch.int32, device=dev)
    wr, gr_, m1r, m2r = w.clone(), g.clone(), m1.clone(), m2.clone()
    torch.ops.srb.adam_shard(g, w, m1, m2, w_bf, torch.tensor(blk_key, dtype=torch.int32, device=dev),
                             torch.tensor(blk_off, dtype=torch.int32, device=dev),
                             torch.tensor(offs, dtype=torch.int64, device=dev),
                             torch.tensor(lens, dtype=torch.int64, device=dev),
                             torch.zeros(len(lens), device=dev), hyper, step)
    for o, n in zip(offs, lens):
        ref.adam_step(wr[o:o + n], gr_[o:o + n], m1r[o:o + n], m2r[o:o + n], lr=0.01, beta1=0.9, beta2=0.999,
                      eps=1e-8, nr_update=1, grad_clip=1.0, l2=0.01, l2_is_weight_decay=True)
    _close(w, wr, 1e-4, 1e-5, "adam w")
    _close(m2, m2r, 1e-4, 1e-7, "adam m2")
    _close(w_bf, wr, 1e-2, 1e-2, "adam bf16 out")
    assert float(g.abs().sum()) == 0.0


def test_width_96_runs_on_the_tcgen05_kernels_not_on_the_library_path(ops):
    """BASELINE config 1's model (tok2vec width 96): K = 96 / 288 is not a multiple of the 64-element
    k-block and N = 96 / 288 not of any tile width.  Round 1 sent it to cuBLAS + seq2col; now the TMA
    zero-fills the K tail and the epilogue masks the partial last N tile - count the launches."""
    torch.manual_seed(11)
    w, nO, nP = 96, 96, 3
    X, mask = _padded_batch((5, 17, 40, 2), w)
    Xb = X.bfloat16()
    W = (torch.randn(nO, nP, 3 * w, device="cuda") * 0.1).bfloat16()
    b = (torch.randn(nO, nP, device="cuda") * 0.1).bfloat16()
    G = torch.ones(nO, device="cuda").bfloat16()
    beta = torch.zeros(nO, device="cuda").bfloat16()
    l0 = ops.launches
    Y, ctx = ops.maxout_block(Xb, W, b, G, beta, mask, window=1, residual=True, dropout=0.0, is_train=True, seed=3)
    assert ops.launches - l0 == 2, "forward = one tcgen05 GEMM (window + bias + maxout) + one LN kernel"
    dY = (torch.randn_like(Y.float()) * mask).bfloat16()
    l0 = ops.launches
    dX, dW, db, dG, dbeta = ops.maxout_block_backward(dY, ctx)
    torch.cuda.synchronize()
    # LN-bwd + seq2col (window materialised for the split-K dW at this width) + dW GEMM + dX GEMM
    assert ops.launches - l0 == 4, ops.launches - l0
    # linear layers at K = 96 (tagger / transition heads on a width-96 tok2vec)
    Wl = (torch.randn(64, 96, device="cuda") * 0.1).bfloat16()
    bl = torch.zeros(64, device="cuda").bfloat16()
    l0 = ops.launches
    out = ops.linear(Xb, Wl, bl)
    assert ops.launches - l0 == 1
    want = Xb.float() @ Wl.float().t()
    _close(out, want, 2e-2, 2e-2 * math.sqrt(96), "linear K=96")


# ---------------------------------------------------------------------------- K7
def test_biluo_kernel_matches_reference_loop(ops, ref):
    from spacy_ray_b200.models.transition_model import (
        TransitionGold, _biluo_steps_reference, transition_backward,
    )
    from spacy_ray_b200.models.transitions import BiluoSystem, spans_to_biluo_actions
    from spacy_ray_b200.nn.batch import make_token_batch
    import numpy as np
    import random

    rng = random.Random(0)
    torch.manual_seed(8)
    L = 5
    system = BiluoSystem([f"L{i}" for i in range(L)])
    lens = [rng.randint(1, 30) for _ in range(67)]
    batch = make_token_batch([np.ones((n, 4), dtype=np.uint64) for n in lens], "cuda:0")
    golds = []
    for n in lens:
        spans, t = [], 0
        while t < n:
            if rng.random() < 0.3:
                ln = min(n - t, rng.randint(1, 3))
                spans.append((t, t + ln, rng.randint(0, L - 1)))
                t += ln
            t += 1
        golds.append(spans_to_biluo_actions(n, spans))
    flat = torch.tensor([a for g in golds for a in g], device="cuda")
    offs = torch.tensor([sum(lens[:i]) for i in range(len(lens))], device="cuda")
    gold = TransitionGold(actions=flat, offsets=offs)
    nF, nO, nP = 3, 64, 2
    Tp = batch.n_rows
    Yf = (torch.randn(Tp, nF * nO * nP, device="cuda") * batch.mask).bfloat16()
    params = {
        "pad": (torch.randn(nF, nO * nP, device="cuda") * 0.3).bfloat16(),
        "b": (torch.randn(nO * nP, device="cuda") * 0.3).bfloat16(),
        "Wu": (torch.randn(system.n_actions, nO, device="cuda") * 0.3).bfloat16(),
        "bu": (torch.randn(system.n_actions, device="cuda") * 0.1).bfloat16(),
        "nF": nF, "nO": nO, "nP": nP,
    }
    rec = ops.transition_steps(system, Yf, params, batch, gold, True)
    pf = {k: (v.float() if torch.is_tensor(v) else v) for k, v in params.items()}
    rref = _biluo_steps_reference(system, Yf.float(), pf, batch, gold, True)
    torch.cuda.synchronize()
    agree = (rec["actions_flat"] == rref["actions_flat"]).float().mean().item()
    assert agree > 0.97, agree          # bf16 hidden vs fp32 can flip near-ties, which then diverge
    # teacher-forced comparison of loss/grad is not possible once trajectories diverge; compare on
    # docs whose whole action sequence agrees
    tok_doc = torch.repeat_interleave(torch.arange(len(lens), device="cuda"), torch.tensor(lens, device="cuda"))
    same_tok = rec["actions_flat"] == rref["actions_flat"]
    doc_ok = torch.ones(len(lens), dtype=torch.int32, device="cuda")
    doc_ok.scatter_reduce_(0, tok_doc, same_tok.to(torch.int32), reduce="amin")
    assert doc_ok.float().mean().item() > 0.8
    # reference records are ordered step-major; rebuild a per-token view of d_scores from the kernel's
    # records and check that it is a valid gradient: rows sum to ~0 and are zero for docs w/o gold
    d = rec["d_scores"].float()[:, : system.n_actions]
    assert d.shape[0] == sum(lens)
    assert float(d.sum(dim=1).abs().max()) < 2e-2
    # loss of the agreeing docs must match the reference contribution closely overall
    assert abs(float(rec["loss"]) - float(rref["loss"])) / max(float(rref["loss"]), 1e-6) < 0.25
    # backward scatter vs reference backward on the kernel's own records
    g = ops.transition_backward(rec, params, Tp)
    rec_f = {"d_scores": d, "hid": rec["hid"].float(), "which": rec["which"], "feats": rec["feats"].long()}
    gr = transition_backward(ref, rec_f, pf, Tp)
    _close(g["dYf"], gr["dYf"], 3e-2, 2e-2, "dYf")
    _close(g["dpad"], gr["dpad"], 3e-2, 5e-2, "dpad")
    _close(g["db"], gr["db"], 3e-2, 5e-2, "db")
    _close(g["dWu"], gr["dWu"], 3e-2, 5e-2, "dWu")


def test_biluo_kernel_teacher_forced_matches_reference_tightly(ops):
    """Teacher forcing makes the kernel and the fp32 reference loop walk the same (gold) trajectory,
    so loss and every row of d_scores are compared at bf16 resolution - not the 25 % / 97 % of the
    free-running comparison above."""
    from spacy_ray_b200.models.transition_model import TransitionGold, _biluo_steps_reference
    from spacy_ray_b200.models.transitions import BiluoSystem, spans_to_biluo_actions
    from spacy_ray_b200.nn.batch import make_token_batch
    import random

    rng = random.Random(3)
    torch.manual_seed(9)
    L = 6
    system = BiluoSystem([f"L{i}" for i in range(L)])
    lens = [rng.randint(1, 70) for _ in range(45)]          # beyond 64 tokens: gold comes from memory, not registers
    batch = make_token_batch([np.ones((n, 4), dtype=np.uint64) for n in lens], "cuda:0")
    golds = []
    for n in lens:
        spans, t = [], 0
        while t < n:
            if rng.random() < 0.3:
                ln = min(n - t, rng.randint(1, 4))
                spans.append((t, t + ln, rng.randint(0, L - 1)))
                t += ln
            t += 1
        golds.append(spans_to_biluo_actions(n, spans))
    flat = torch.tensor([a for g in golds for a in g], device="cuda")
    offs = torch.tensor([sum(lens[:i]) for i in range(len(lens))], device="cuda")
    gold = TransitionGold(actions=flat, offsets=offs, teacher_forced=True)
    nF, nO, nP = 3, 64, 2
    Tp = batch.n_rows
    Yf = (torch.randn(Tp, nF * nO * nP, device="cuda") * batch.mask).bfloat16()
    params = {
        "pad": (torch.randn(nF, nO * nP, device="cuda") * 0.3).bfloat16(),
        "b": (torch.randn(nO * nP, device="cuda") * 0.3).bfloat16(),
        "Wu": (torch.randn(system.n_actions, nO, device="cuda") * 0.3).bfloat16(),
        "bu": (torch.randn(system.n_actions, device="cuda") * 0.1).bfloat16(),
        "nF": nF, "nO": nO, "nP": nP,
    }
    rec = ops.transition_steps(system, Yf, params, batch, gold, True)
    pf = {k: (v.float() if torch.is_tensor(v) else v) for k, v in params.items()}
    rref = _biluo_steps_reference(system, Yf.float(), pf, batch, gold, True)
    torch.cuda.synchronize()
    assert abs(float(rec["loss"]) - float(rref["loss"])) <= 1e-2 * max(float(rref["loss"]), 1e-9)
    # one record per token in both; key both by the padded row of the token (feature slot 0)
    A = system.n_actions
    key_k = rec["feats"][:, 0].long()
    key_r = rref["feats"][:, 0].long()
    assert torch.equal(torch.sort(key_k).values, torch.sort(key_r).values)
    dk = rec["d_scores"].float()[:, :A][torch.argsort(key_k)]
    dr = rref["d_scores"][torch.argsort(key_r)]
    err = (dk - dr).abs()
    assert float(err.max()) < 1e-2 * float(dr.abs().max()) + 1e-3, float(err.max())
    assert torch.equal(rec["feats"][torch.argsort(key_k)].long(), rref["feats"][torch.argsort(key_r)].long())
    hk = rec["hid"].float()[torch.argsort(key_k)]
    hr = rref["hid"][torch.argsort(key_r)]
    assert float((hk - hr).abs().max()) < 2e-2 * float(hr.abs().max())


# ---------------------------------------------------------------------------- end to end
def test_gpu_training_loss_goes_down_and_uses_native_kernels():
    from conftest import multi_cfg
    from spacy_ray_b200.config import Config
    from spacy_ray_b200.worker import Worker

    cfg = Config().from_str(multi_cfg(["ner"], width=64, depth=2, n_docs=400, max_len=16, hidden=64), interpolate=False)
    w = Worker(cfg, rank=0, num_workers=1, use_gpu=0, mode="sync", comm="auto")
    w.set_proxy(None)
    assert w.proxy.comm.name == "fused"
    nlp = w.nlp
    ops = nlp.get_pipe("ner").model.ops
    assert ops.name == "b200"
    exs = list(w.train_corpus(nlp))
    hist = []
    for step in range(100):
        losses = {}
        lo = (step * 64) % (len(exs) - 64)
        nlp.update(exs[lo:lo + 64], drop=0.1, sgd=False, losses=losses)
        w.proxy.step()
        hist.append(float(losses["ner"]))
    w.proxy.comm.check()
    assert ops.launches > 0 and w.proxy.comm.launches == 100 * sum(w.proxy.comm.kernels_for(b) for b in range(w.proxy.comm.plan.n))
    assert sum(hist[-5:]) < 0.5 * sum(hist[:5]), hist[::6]
    scores = nlp.evaluate(exs[:100])
    assert scores["ents_f"] > 0.25, scores


def test_worker_training_loop_uses_the_device_resident_engine(tmp_path):
    """CLI-level path on the GPU: Worker.train() -> train_while_improving -> nlp.update is served by
    engine.Trainer (native collate, one H2D copy, CUDA-graph replay incl. the fused comm kernel)."""
    from conftest import multi_cfg
    from spacy_ray_b200.config import Config
    from spacy_ray_b200.worker import Worker

    text = multi_cfg(["ner"], width=64, depth=2, n_docs=600, max_len=16, hidden=64)
    text = text.replace("max_steps = 10", "max_steps = 40").replace("eval_frequency = 5", "eval_frequency = 20")
    text = text.replace("dropout = 0.0", "dropout = 0.1")
    text += """
[training.batcher]
@batchers = "spacy.batch_by_sequence.v1"
size = 64
"""
    w = Worker(Config().from_str(text, interpolate=False), rank=0, num_workers=1, use_gpu=0, output_path=tmp_path)
    w.set_proxy(None)
    w.train(None, None)
    w.join(timeout=300)
    assert w.get_error() is None
    trainer = getattr(w.nlp, "_trainer", None)
    assert trainer is not None and trainer.steps >= 39, "fast path was not used"
    assert len(trainer._graphs) >= 1                      # at least one bucket was captured and replayed
    assert (tmp_path / "model-last" / "ner" / "model").exists()
    stats = w.get_stats()
    assert stats["docs"] > 0 and stats["docs_per_sec"] > 0


def test_engine_serves_accumulate_gradient_and_streamed_corpora(tmp_path):
    """Round 1 dropped to the per-op path for accumulate_gradient > 1 and needed the whole corpus in
    memory.  Now micro-batches replay an accumulate-only graph (the exchange runs once per full batch)
    and batches of docs that are not in the pre-built store are collated from a per-batch store."""
    from conftest import multi_cfg
    from spacy_ray_b200.config import Config
    from spacy_ray_b200.worker import Worker

    text = multi_cfg(["ner"], width=64, depth=2, n_docs=600, max_len=16, hidden=64)
    text = text.replace("max_steps = 10", "max_steps = 24").replace("eval_frequency = 5", "eval_frequency = 12")
    text = text.replace("[training]\n", "[training]\naccumulate_gradient = 2\nmax_epochs = -1\n")
    text += """
[training.batcher]
@batchers = "spacy.batch_by_sequence.v1"
size = 64
"""
    import os

    os.environ["SRB_FAST_PATH_SAMPLE"] = "128"          # the store only knows the first 128 docs of the stream
    try:
        w = Worker(Config().from_str(text, interpolate=False), rank=0, num_workers=1, use_gpu=0)
        w.set_proxy(None)
        w.train(None, None)
        w.join(timeout=300)
    finally:
        os.environ.pop("SRB_FAST_PATH_SAMPLE", None)
    assert w.get_error() is None
    trainer = getattr(w.nlp, "_trainer", None)
    assert trainer is not None and trainer.exchange is False
    assert trainer.steps >= 2 * 23, "micro-batches did not go through the engine"
    assert trainer.adhoc_batches > 0, "streamed docs outside the store must use the per-batch store"
    comm = w.proxy.comm
    assert int(comm.step_t.item()) == 24, "one optimizer step per FULL batch"
    assert float(w.proxy.grad_flat.abs().sum()) == 0.0
