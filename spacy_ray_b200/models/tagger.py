"""``spacy.Tagger``: tok2vec -> zero-initialised Softmax over tags."""
from __future__ import annotations

from typing import Optional

import torch

from ..nn.batch import TokenBatch
from ..nn.layers import Softmax
from ..nn.model import Model


def build_tagger_model(tok2vec: Model, nO: Optional[int] = None, normalize: bool = False) -> Model:
    width = tok2vec.get_dim("nO")
    output = Softmax(nO, width)

    def init(model: Model, X=None, Y=None):
        tok2vec.initialize()
        if output.has_dim("nO") is None:
            if model.has_dim("nO") is None:
                raise ValueError("Tagger model: number of labels (nO) not set before initialize")
            output.set_dim("nO", model.get_dim("nO"))
        output.initialize()

    def forward(model: Model, batch: TokenBatch, is_train: bool):
        X, bp_t2v = tok2vec(batch, is_train)
        P, bp_out = output(X, is_train)

        def backprop(d_logits):
            bp_t2v(bp_out(d_logits))
            return None

        return P, backprop

    def update_with_labels(batch: TokenBatch, labels: torch.Tensor):
        """Forward + loss + backward in one call (fused head: logits, softmax,
        ``P - onehot`` and the three gradients come out of one ``softmax_xent``
        op - K6 in SURVEY.md 2.7).  ``labels``: (Tp,) int64, -1 = no gold.
        Returns ``(loss, guesses)``."""
        X, bp_t2v = tok2vec(batch, True)
        W, b = output.get_param("W"), output.get_param("b")
        if getattr(model.ops, "fused", False):      # gradients accumulate straight into the flat bucket
            go = {"W": output.grad_buffer("W"), "b": output.grad_buffer("b")}
            loss, _d, guesses, dX, dW, db = model.ops.softmax_xent(X, W, b, labels, grad_out=go)
        else:
            loss, _d, guesses, dX, dW, db = model.ops.softmax_xent(X, W, b, labels)
        output.inc_grad("W", dW)
        output.inc_grad("b", db)
        bp_t2v(dX)
        return loss, guesses

    model = Model(
        "tagger", forward, init=init, dims={"nO": nO},
        layers=[tok2vec, output], refs={"tok2vec": tok2vec, "output": output},
    )
    model.attrs["update_with_labels"] = update_with_labels
    return model
