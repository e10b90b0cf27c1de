"""``spacy.TextCatCNN`` / ``spacy.TextCatReduce``: tok2vec -> pool over the tokens of each doc -> affine ->
softmax (exclusive classes) or logistic (multi-label).  The reference trains whatever spaCy pipeline the
config names (``/root/reference/spacy_ray/worker.py:88-96``); this is the document-level head of that zoo.

The pooled matrix is (n_docs, width) - tiny next to the token matrices - so it is built with plain
tensor ops from the padded-ragged layout (``nn/batch.py``); the affine layer is the shared ``Linear``.
``backprop`` takes the gradient w.r.t. the LOGITS (``probs - truth``, as ``Softmax`` does)."""
from __future__ import annotations

from typing import Optional

import torch

from ..nn.batch import TokenBatch
from ..nn.layers import Linear
from ..nn.model import Model


def _segments(batch: TokenBatch):
    lens = torch.as_tensor(batch.lengths, dtype=torch.int64, device=batch.device)
    seg = torch.repeat_interleave(torch.arange(len(batch.lengths), device=batch.device), lens)
    return lens, seg


def build_textcat_model(tok2vec: Model, exclusive_classes: bool = True, nO: Optional[int] = None,
                        use_reduce_mean: bool = True, use_reduce_max: bool = False, use_reduce_first: bool = False,
                        use_reduce_last: bool = False, ngram_size: Optional[int] = None,
                        no_output_layer: bool = False) -> Model:
    if use_reduce_first or use_reduce_last:
        raise ValueError("TextCat: only mean / max pooling are implemented (use_reduce_first / use_reduce_last are not)")
    if not (use_reduce_mean or use_reduce_max):
        raise ValueError("TextCat: enable use_reduce_mean and/or use_reduce_max")
    width = tok2vec.get_dim("nO")
    n_pool = int(use_reduce_mean) + int(use_reduce_max)
    output = Linear(nO, width * n_pool, init_zero=True, name="textcat_output")

    def init(model: Model, X=None, Y=None):
        tok2vec.initialize()
        if output.has_dim("nO") is None:
            if model.has_dim("nO") is None:
                raise ValueError("TextCat model: number of labels (nO) not set before initialize")
            output.set_dim("nO", model.get_dim("nO"))
        output.initialize()

    def forward(model: Model, batch: TokenBatch, is_train: bool):
        if not batch.lengths:
            raise ValueError("TextCat needs host-side doc lengths (generic path); it is not served by engine.Trainer")
        X, bp_t2v = tok2vec(batch, is_train)                       # (Tp, width), pad rows zero
        Xt = batch.unpad(X).float()                                # (T, width)
        lens, seg = _segments(batch)
        B = len(batch.lengths)
        inv = 1.0 / lens.clamp(min=1).to(torch.float32).unsqueeze(1)
        parts, ctx = [], {}
        if use_reduce_mean:
            sums = torch.zeros((B, width), dtype=torch.float32, device=X.device).index_add_(0, seg, Xt)
            parts.append(sums * inv)
        if use_reduce_max:
            mx = torch.full((B, width), float("-inf"), dtype=torch.float32, device=X.device)
            mx = mx.scatter_reduce(0, seg.unsqueeze(1).expand(-1, width), Xt, reduce="amax", include_self=True)
            mx = torch.where(torch.isinf(mx), torch.zeros_like(mx), mx)      # empty docs
            ctx["argmax"] = (Xt == mx[seg])
            parts.append(mx)
        pooled = torch.cat(parts, dim=1).to(X.dtype)
        logits, bp_out = output(pooled, is_train)
        logits = logits.float()
        scores = torch.softmax(logits, dim=1) if exclusive_classes else torch.sigmoid(logits)

        def backprop(d_logits: torch.Tensor):
            d_pooled = bp_out(d_logits.to(pooled.dtype)).float()
            dXt = torch.zeros_like(Xt)
            off = 0
            if use_reduce_mean:
                dXt += (d_pooled[:, off:off + width] * inv)[seg]
                off += width
            if use_reduce_max:
                hit = ctx["argmax"].float()
                hit = hit / hit.new_zeros((B, width)).index_add_(0, seg, hit).clamp(min=1.0)[seg]   # ties share
                dXt += d_pooled[:, off:off + width][seg] * hit
            bp_t2v(batch.pad(dXt.to(X.dtype)))
            return None

        return scores, backprop

    model = Model("textcat", forward, init=init, dims={"nO": nO},
                  layers=[tok2vec, output], refs={"tok2vec": tok2vec, "output": output})
    model.attrs["exclusive_classes"] = bool(exclusive_classes)
    return model
