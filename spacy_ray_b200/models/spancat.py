"""``spacy.SpanCategorizer.v1``: score candidate token spans against a label set (overlapping, multi-label).

tok2vec -> for every candidate span (from a suggester: all n-grams of the configured sizes) the mean and the max
of its token vectors plus its first and last token vector -> one Maxout hidden layer -> affine -> logistic.
Part of the spaCy zoo the reference can train (``/root/reference/spacy_ray/worker.py:88-96``); generic path,
plain tensor ops on the padded-ragged layout (``nn/batch.py``).  ``backprop`` takes the gradient w.r.t. the
LOGITS (``probs - truth``)."""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np
import torch

from ..config import registry
from ..nn.batch import TokenBatch
from ..nn.layers import Linear, Maxout
from ..nn.model import Model


@registry.misc("spacy.ngram_suggester.v1")
def ngram_suggester(sizes: Sequence[int] = (1, 2, 3)) -> Callable[[Sequence[int]], List[Tuple[int, int, int]]]:
    """``suggest(doc_lengths) -> [(doc, start, end_exclusive), ...]``: every n-gram of the given sizes."""
    sizes = sorted(int(s) for s in sizes)

    def suggest(lengths: Sequence[int]) -> List[Tuple[int, int, int]]:
        out = []
        for d, n in enumerate(lengths):
            for size in sizes:
                for a in range(0, n - size + 1):
                    out.append((d, a, a + size))
        return out

    suggest.sizes = sizes                         # type: ignore[attr-defined]
    return suggest


@registry.misc("spacy.ngram_range_suggester.v1")
def ngram_range_suggester(min_size: int = 1, max_size: int = 3):
    return ngram_suggester(list(range(int(min_size), int(max_size) + 1)))


@registry.layers("spacy.mean_max_reducer.v1")
def mean_max_reducer(hidden_size: int = 128) -> dict:
    """Config-compatibility marker for ``[components.spancat.model.reducer]``: the model reads ``hidden_size``."""
    return {"hidden_size": int(hidden_size)}


@registry.layers("spacy.LinearLogistic.v1")
def linear_logistic(nO: Optional[int] = None, nI: Optional[int] = None) -> dict:
    """Config-compatibility marker for ``[components.spancat.model.scorer]`` (affine + logistic is built in)."""
    return {"nO": nO, "nI": nI}


def build_spancat_model(tok2vec: Model, reducer: Optional[Model] = None, scorer: Optional[Model] = None,
                        nO: Optional[int] = None, hidden_size: int = 128, maxout_pieces: int = 3) -> Model:
    """``reducer`` / ``scorer`` sub-blocks of upstream configs are accepted and only mined for ``hidden_size`` /
    ``nO``: the pooling (mean, max, first, last) and the Maxout -> affine -> logistic head are fixed here."""
    width = tok2vec.get_dim("nO")
    if isinstance(reducer, dict):
        hidden_size = int(reducer.get("hidden_size", hidden_size))
    hidden = Maxout(int(hidden_size), 4 * width, nP=int(maxout_pieces))
    output = Linear(nO, int(hidden_size), init_zero=True, name="spancat_output")

    def init(model: Model, X=None, Y=None):
        tok2vec.initialize()
        hidden.initialize()
        if output.has_dim("nO") is None:
            if model.has_dim("nO") is None:
                raise ValueError("SpanCategorizer model: number of labels (nO) not set before initialize")
            output.set_dim("nO", model.get_dim("nO"))
        output.initialize()

    def forward(model: Model, inputs, is_train: bool):
        batch, spans = inputs                                        # spans: [(doc, start, end_exclusive), ...]
        if not batch.lengths:
            raise ValueError("SpanCategorizer needs host-side doc lengths (generic path)")
        X, bp_t2v = tok2vec(batch, is_train)
        dev = X.device
        S = len(spans)
        if S == 0:
            return torch.zeros((0, model.get_dim("nO")), device=dev), (lambda d: bp_t2v(torch.zeros_like(X)))
        sp = np.asarray(spans, dtype=np.int64)
        starts = torch.as_tensor(np.asarray(batch.starts, dtype=np.int64)[sp[:, 0]] + sp[:, 1], device=dev)   # padded rows
        lens = torch.as_tensor(sp[:, 2] - sp[:, 1], device=dev)
        L = int(lens.max())
        offs = torch.arange(L, device=dev).unsqueeze(0)                                    # (1, L)
        valid = offs < lens.unsqueeze(1)                                                   # (S, L)
        rows = (starts.unsqueeze(1) + offs).clamp(max=X.shape[0] - 1)                      # (S, L)
        G = X.float()[rows]                                                                # (S, L, w)
        Gm = G * valid.unsqueeze(2)
        mean = Gm.sum(1) / lens.unsqueeze(1).float()
        neg = torch.full_like(G, float("-inf"))
        mx, arg = torch.where(valid.unsqueeze(2), G, neg).max(dim=1)                       # (S, w)
        first = G[:, 0]
        last = G[torch.arange(S, device=dev), (lens - 1)]
        feats = torch.cat([mean, mx, first, last], dim=1).to(X.dtype)
        H, bp_h = hidden(feats, is_train)
        logits, bp_o = output(H, is_train)
        scores = torch.sigmoid(logits.float())

        def backprop(d_logits: torch.Tensor):
            dF = bp_h(bp_o(d_logits.to(H.dtype))).float()
            w = X.shape[1]
            d_mean, d_max, d_first, d_last = dF[:, :w], dF[:, w:2 * w], dF[:, 2 * w:3 * w], dF[:, 3 * w:]
            dG = (d_mean / lens.unsqueeze(1).float()).unsqueeze(1) * valid.unsqueeze(2)    # (S, L, w)
            dG = dG.clone()
            dG.scatter_add_(1, arg.unsqueeze(1), d_max.unsqueeze(1))
            dG[:, 0] += d_first
            dG[torch.arange(S, device=dev), (lens - 1)] += d_last
            dX = torch.zeros(X.shape, dtype=torch.float32, device=dev)
            dX.index_add_(0, rows[valid], dG[valid])
            bp_t2v(dX.to(X.dtype))
            return None

        return scores, backprop

    return Model("spancat", forward, init=init, dims={"nO": nO},
                 layers=[tok2vec, hidden, output], refs={"tok2vec": tok2vec, "hidden": hidden, "output": output})
