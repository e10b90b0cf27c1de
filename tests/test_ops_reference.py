import math

import pytest
import torch

from spacy_ray_b200.ops.torch_ops import TorchOps, fmix64_int, hash_rows_int

ops = TorchOps("cpu")


def test_hash_rows_matches_python_bigint_reference():
    ids = [0, 1, 2, 12345678901234567, (1 << 63) + 5, (1 << 64) - 1]
    as_i64 = torch.tensor([i - (1 << 64) if i >= (1 << 63) else i for i in ids], dtype=torch.int64)
    for seed, n_rows in [(8, 5000), (11, 1000), (9, 2500)]:
        got = ops.hash_rows(as_i64, seed, n_rows).tolist()
        want = [list(hash_rows_int(i, seed, n_rows)) for i in ids]
        assert got == want


def test_dropout_mask_is_deterministic_and_has_the_right_rate():
    m1 = ops.dropout_mask(7, 200, 64, 0.25)
    m2 = ops.dropout_mask(7, 200, 64, 0.25)
    assert torch.equal(m1, m2)
    keep = (m1 > 0).float().mean().item()
    assert abs(keep - 0.75) < 0.02
    assert torch.allclose(m1[m1 > 0], torch.tensor(1 / 0.75))
    assert not torch.equal(m1, ops.dropout_mask(8, 200, 64, 0.25))


def _padded(T_docs=(3, 5, 2), w=8):
    rows = sum(T_docs) + len(T_docs) + 1
    mask = torch.zeros(rows, 1)
    r = 1
    for n in T_docs:
        mask[r:r + n] = 1
        r += n + 1
    X = torch.randn(rows, w) * mask
    return X, mask


def test_seq2col_respects_doc_boundaries_through_pad_rows():
    X, mask = _padded()
    Xw = ops.seq2col(X, 1)
    w = X.shape[1]
    assert torch.equal(Xw[:, w:2 * w], X)
    # first token of doc 0 is row 1: its left neighbour is the zero pad row 0
    assert torch.equal(Xw[1, :w], torch.zeros(w))
    # last token of doc 0 (row 3) has the pad row 4 on its right
    assert torch.equal(Xw[3, 2 * w:], torch.zeros(w))
    # adjoint test: <seq2col(X), D> == <X, backprop(D)>
    D = torch.randn_like(Xw)
    assert torch.allclose((Xw * D).sum(), (X * ops.backprop_seq2col(D, 1)).sum(), atol=1e-4)


@pytest.mark.parametrize("window,residual,ln", [(0, False, True), (1, True, True), (1, False, False)])
def test_maxout_block_backward_matches_autograd(window, residual, ln):
    torch.manual_seed(0)
    X, mask = _padded(w=8)
    nO, nP = 8, 3
    nI = 8 * (3 if window else 1)
    W = torch.randn(nO, nP, nI) * 0.3
    b = torch.randn(nO, nP) * 0.1
    G = torch.rand(nO) + 0.5 if ln else None
    beta = torch.randn(nO) * 0.1 if ln else None
    Y, ctx = ops.maxout_block(X, W, b, G, beta, mask, window=window, residual=residual, dropout=0.2,
                              is_train=True, seed=3)
    dY = torch.randn_like(Y)
    dX, dW, db, dG, dbeta = ops.maxout_block_backward(dY, ctx)

    Xa, Wa, ba = X.clone().requires_grad_(), W.clone().requires_grad_(), b.clone().requires_grad_()
    Xw = ops.seq2col(Xa, window)
    Z = Xw @ Wa.reshape(nO * nP, nI).t() + ba.reshape(-1)
    H = Z.view(-1, nO, nP).max(dim=2).values
    if ln:
        Ga, Ba = G.clone().requires_grad_(), beta.clone().requires_grad_()
        N = (H - H.mean(1, keepdim=True)) / torch.sqrt(H.var(1, unbiased=False, keepdim=True) + 1e-8) * Ga + Ba
    else:
        N = H
    N = N * ops.dropout_mask(3, N.shape[0], N.shape[1], 0.2)
    Yr = ((Xa + N) if residual else N) * mask
    assert torch.allclose(Y, Yr.detach(), atol=1e-5)
    Yr.backward(dY)
    assert torch.allclose(dX, Xa.grad, atol=1e-4)
    assert torch.allclose(dW, Wa.grad, atol=1e-4)
    assert torch.allclose(db, ba.grad, atol=1e-4)
    if ln:
        assert torch.allclose(dG, Ga.grad, atol=1e-4) and torch.allclose(dbeta, Ba.grad, atol=1e-4)
    # pad rows stay exactly zero
    assert float((Y * (1 - mask)).abs().sum()) == 0.0


def test_multi_hash_embed_forward_backward():
    torch.manual_seed(0)
    attrs = torch.randint(-(1 << 62), 1 << 62, (9, 4), dtype=torch.int64)
    mask = torch.ones(9, 1)
    mask[0] = 0
    mask[5] = 0
    tables = [torch.randn(n, 4) for n in (50, 20, 30, 30)]
    seeds, cols = [8, 9, 10, 11], [0, 1, 2, 3]
    Y = ops.multi_hash_embed(attrs, mask, tables, seeds, cols)
    assert Y.shape == (9, 16) and float(Y[0].abs().sum()) == 0.0
    t = 3
    rows = hash_rows_int(int(attrs[t, 1]) & ((1 << 64) - 1), 9, 20)
    assert torch.allclose(Y[t, 4:8], sum(tables[1][r] for r in rows), atol=1e-5)
    dY = torch.randn(9, 16)
    grads = ops.multi_hash_embed_backward(dY, attrs, mask, [50, 20, 30, 30], seeds, cols)
    tabs = [x.clone().requires_grad_() for x in tables]
    (ops.multi_hash_embed(attrs, mask, tabs, seeds, cols) * dY).sum().backward()
    for g, tt in zip(grads, tabs):
        assert torch.allclose(g, tt.grad, atol=1e-5)


def test_softmax_xent_grads():
    torch.manual_seed(0)
    X = torch.randn(7, 5)
    W = torch.randn(4, 5, requires_grad=True)
    b = torch.randn(4, requires_grad=True)
    labels = torch.tensor([0, 3, -1, 2, 1, -1, 0])
    loss, d, guesses, dX, dW, db = ops.softmax_xent(X, W.detach(), b.detach(), labels)
    P = torch.softmax(X @ W.t() + b, -1)
    assert torch.equal(guesses, P.argmax(1))
    have = labels >= 0
    onehot = torch.zeros_like(P)
    onehot[have, labels[have]] = 1
    dref = (P - onehot) * have.unsqueeze(1)
    assert torch.allclose(d, dref.detach(), atol=1e-6)
    assert torch.allclose(loss, (dref ** 2).sum().detach(), atol=1e-6)
    assert torch.allclose(dW, dref.detach().t() @ X, atol=1e-5) and torch.allclose(dX, dref.detach() @ W.detach(), atol=1e-5)


def test_adam_step_matches_formula_with_clip_and_weight_decay():
    torch.manual_seed(0)
    w = torch.randn(10)
    g = torch.randn(10) * 5
    m1, m2 = torch.zeros(10), torch.zeros(10)
    w0, g0 = w.clone(), g.clone()
    ops.adam_step(w, g, m1, m2, lr=0.01, beta1=0.9, beta2=0.999, eps=1e-8, nr_update=1, grad_clip=1.0,
                  l2=0.01, l2_is_weight_decay=True)
    gc = g0 * (1.0 / g0.norm())
    e1, e2 = 0.1 * gc, 0.001 * gc * gc
    lr_t = 0.01 * math.sqrt(1 - 0.999) / (1 - 0.9)
    expect = (w0 - lr_t * e1 / (e2.sqrt() + 1e-8)) * (1 - 0.01 * 0.01)
    assert torch.allclose(w, expect, atol=1e-6) and float(g.abs().sum()) == 0.0
    assert torch.allclose(m1, e1) and torch.allclose(m2, e2)
