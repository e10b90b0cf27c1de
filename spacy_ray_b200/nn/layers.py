"""Layers (thinc/spaCy-ml equivalents) built on ``Model`` + an ``Ops`` backend.

Node structure mirrors upstream so the ``(node.id, param_name)`` key space and
the per-node grouping that ``divide_params`` relies on
(``/root/reference/spacy_ray/util.py:57-75``) are the same: every ``Maxout``
owns ``W, b``; every ``LayerNorm`` owns ``G, b``; every ``HashEmbed`` owns
``E``.  Execution, however, is *block-wise*: the parent node of a
(Maxout, LayerNorm) pair runs them - plus expand_window, dropout and the
residual add - as one ``ops.maxout_block`` call, which on the B200 backend is a
single fused tcgen05 GEMM kernel instead of thinc's ~8 launches.
"""
from __future__ import annotations

import itertools
from typing import Any, Callable, List, Optional, Sequence, Tuple

import torch

from .batch import TokenBatch
from .model import Model

# ---- seeded init + dropout seeds ------------------------------------------
_init_gen = torch.Generator(device="cpu")
_init_gen.manual_seed(0)
_dropout_seed = itertools.count(1)


def fix_random_seed(seed: int = 0) -> None:
    """Seed parameter init and the dropout counter.  Every rank calls this with
    the config seed before building the pipeline, which is the *only* thing
    that makes initial weights agree across ranks in the reference
    (``worker.py:91`` - nothing broadcasts them); our launcher additionally
    broadcasts rank 0's weights (``parallel/launcher.py``)."""
    global _dropout_seed
    _init_gen.manual_seed(int(seed))
    torch.manual_seed(int(seed))
    _dropout_seed = itertools.count(int(seed) * 1_000_003 + 1)


def next_dropout_seed() -> int:
    return next(_dropout_seed)


def offset_dropout_stream(rank: int) -> None:
    """Give each data-parallel rank a different dropout stream."""
    global _dropout_seed
    _dropout_seed = itertools.count(next(_dropout_seed) + rank * 7_919_000_003)


def set_dropout_rate(model: Model, rate: float) -> None:
    for node in model.walk():
        if "dropout_rate" in node.attrs:
            node.attrs["dropout_rate"] = float(rate)


# ---- parameter-holding leaf nodes -----------------------------------------
def _noop_forward(model: Model, X, is_train):
    raise RuntimeError(
        f"'{model.name}' is a parameter-holder node; it is executed by its parent block"
    )


def Maxout(nO: Optional[int] = None, nI: Optional[int] = None, nP: int = 3) -> Model:
    """Holds ``W (nO, nP, nI)`` and ``b (nO, nP)``.  Standalone forward works too."""

    def init(model: Model, X=None, Y=None):
        if X is not None and model.has_dim("nI") is None:
            model.set_dim("nI", int(X.shape[1]))
        if Y is not None and model.has_dim("nO") is None:
            model.set_dim("nO", int(Y.shape[1]))
        nO_, nP_, nI_ = model.get_dim("nO"), model.get_dim("nP"), model.get_dim("nI")
        if model.has_param("W") is not True:
            model.set_param("W", model.ops.glorot_uniform((nO_, nP_, nI_), nI_, nO_, _init_gen))
            model.set_param("b", model.ops.alloc((nO_, nP_)))

    def forward(model: Model, X, is_train):
        mask = torch.ones((X.shape[0], 1), dtype=torch.float32, device=X.device)
        Y, ctx = model.ops.maxout_block(
            X, model.get_param("W"), model.get_param("b"), None, None, mask,
            is_train=is_train,
        )

        def backprop(dY):
            dX, dW, db, _, _ = model.ops.maxout_block_backward(dY, ctx)
            model.inc_grad("W", dW)
            model.inc_grad("b", db)
            return dX

        return Y, backprop

    return Model(
        "maxout", forward, init=init,
        dims={"nO": nO, "nI": nI, "nP": nP}, params={"W": None, "b": None},
    )


def LayerNorm(nI: Optional[int] = None) -> Model:
    def init(model: Model, X=None, Y=None):
        if X is not None and model.has_dim("nI") is None:
            model.set_dim("nI", int(X.shape[1]))
        n = model.get_dim("nI")
        if model.has_param("G") is not True:
            model.set_param("G", model.ops.alloc((n,)) + 1.0)
            model.set_param("b", model.ops.alloc((n,)))

    return Model("layernorm", _noop_forward, init=init, dims={"nI": nI, "nO": nI}, params={"G": None, "b": None})


def HashEmbed(nO: int, nV: int, *, seed: int, column: int) -> Model:
    """Holds ``E (nV, nO)``.  ``seed``/``column`` are read by the parent."""

    def init(model: Model, X=None, Y=None):
        if model.has_param("E") is not True:
            model.set_param(
                "E", model.ops.uniform((model.get_dim("nV"), model.get_dim("nO")), -0.1, 0.1, _init_gen)
            )

    return Model(
        "hashembed", _noop_forward, init=init,
        dims={"nO": nO, "nV": nV}, params={"E": None},
        attrs={"seed": seed, "column": column},
    )


# ---- block executor shared by MultiHashEmbed-mix and the encoder ------------
def _run_block(
    parent: Model, maxout: Model, norm: Optional[Model], X: torch.Tensor, mask: torch.Tensor,
    *, window: int, residual: bool, is_train: bool,
) -> Tuple[torch.Tensor, Callable]:
    ops = parent.ops
    rate = float(parent.attrs.get("dropout_rate", 0.0))
    seed = next_dropout_seed() if (is_train and rate > 0.0) else 0
    Y, ctx = ops.maxout_block(
        X, maxout.get_param("W"), maxout.get_param("b"),
        norm.get_param("G") if norm is not None else None,
        norm.get_param("b") if norm is not None else None,
        mask, window=window, residual=residual, dropout=rate, is_train=is_train, seed=seed,
    )

    def backprop(dY: torch.Tensor) -> torch.Tensor:
        grad_out = None
        if getattr(ops, "fused", False):
            grad_out = {"W": maxout.grad_buffer("W"), "b": maxout.grad_buffer("b")}
            if norm is not None:
                grad_out["G"], grad_out["beta"] = norm.grad_buffer("G"), norm.grad_buffer("b")
        dX, dW, db, dG, dbeta = ops.maxout_block_backward(dY, ctx, grad_out=grad_out)
        maxout.inc_grad("W", dW)
        maxout.inc_grad("b", db)
        if norm is not None:
            norm.inc_grad("G", dG)
            norm.inc_grad("b", dbeta)
        return dX

    return Y, backprop


# ---- MultiHashEmbed --------------------------------------------------------
DEFAULT_ATTRS = ("NORM", "PREFIX", "SUFFIX", "SHAPE")
# columns of Doc.to_array() / the native featuriser; LOWER aliases NORM (both are the lower-cased form here)
ATTR_COLUMNS = {"NORM": 0, "PREFIX": 1, "SUFFIX": 2, "SHAPE": 3, "LOWER": 0}


def MultiHashEmbed(
    width: int,
    attrs: Sequence[str] = DEFAULT_ATTRS,
    rows: Sequence[int] = (5000, 1000, 2500, 2500),
    include_static_vectors: bool = False,
    maxout_pieces: int = 3,
) -> Model:
    """TokenBatch -> (Tp, width): 4-way hashed embeddings per attribute,
    concatenated, mixed by Maxout + LayerNorm (+ dropout)."""
    if len(attrs) != len(rows):
        raise ValueError("MultiHashEmbed: attrs and rows must have the same length")
    unknown = [a for a in attrs if a not in ATTR_COLUMNS]
    if unknown:
        raise ValueError(f"MultiHashEmbed: unsupported attrs {unknown}; this build featurises {sorted(ATTR_COLUMNS)}")
    seed = 7
    embeds = []
    for attr, nV in zip(attrs, rows):
        seed += 1
        embeds.append(HashEmbed(width, int(nV), seed=seed, column=ATTR_COLUMNS[attr]))
    static = None
    if include_static_vectors:
        from .staticvectors import StaticVectors

        static = StaticVectors(width)           # vectors[rows] @ W^T, joins the concat as one more block
    mix = Maxout(nO=width, nI=width * (len(embeds) + (1 if static is not None else 0)), nP=maxout_pieces)
    norm = LayerNorm(width)

    def init(model: Model, X=None, Y=None):
        for layer in model.layers:
            layer.initialize()

    def forward(model: Model, batch: TokenBatch, is_train: bool):
        ops = model.ops
        tables = [e.get_param("E") for e in embeds]
        seeds = [e.attrs["seed"] for e in embeds]
        cols = [e.attrs["column"] for e in embeds]
        concat = ops.multi_hash_embed(batch.attrs, batch.mask, tables, seeds, cols)
        bp_static = None
        if static is not None:
            sv, bp_static = static(batch, is_train)
            concat = torch.cat([concat, sv.to(concat.dtype) * batch.mask.to(concat.dtype)], dim=1)
        Y, bp_mix = _run_block(model, mix, norm, concat, batch.mask, window=0, residual=False, is_train=is_train)

        def backprop(dY):
            d_concat = bp_mix(dY)
            if bp_static is not None:
                w_hash = width * len(embeds)
                bp_static(d_concat[:, w_hash:].contiguous())
                d_concat = d_concat[:, :w_hash].contiguous()
            bufs = [e.grad_buffer("E") for e in embeds] if getattr(ops, "fused", False) else None
            if bufs is not None and any(b is None for b in bufs):
                bufs = None
            grads = ops.multi_hash_embed_backward(
                d_concat, batch.attrs, batch.mask, [int(t.shape[0]) for t in tables], seeds, cols, out=bufs,
                perm=batch.extra.get("embed_perm"),
            )
            for e, g in zip(embeds, grads):
                e.inc_grad("E", g)
            return None

        return Y, backprop

    return Model(
        "multihashembed", forward, init=init, dims={"nO": width},
        layers=[*embeds, *([static] if static is not None else []), mix, norm],
        attrs={"dropout_rate": 0.0, "attrs": list(attrs)},
        refs={"mix": mix, "norm": norm},
    )


# ---- MaxoutWindowEncoder ---------------------------------------------------
def _residual_block(width: int, window_size: int, maxout_pieces: int) -> Model:
    maxout = Maxout(nO=width, nI=width * (2 * window_size + 1), nP=maxout_pieces)
    norm = LayerNorm(width)

    def init(model: Model, X=None, Y=None):
        maxout.initialize()
        norm.initialize()

    def forward(model: Model, Xm: Tuple[torch.Tensor, torch.Tensor], is_train: bool):
        X, mask = Xm
        return _run_block(model, maxout, norm, X, mask, window=window_size, residual=True, is_train=is_train)

    return Model(
        "residual_maxout_window", forward, init=init, dims={"nO": width, "nI": width},
        layers=[maxout, norm], attrs={"dropout_rate": 0.0, "window_size": window_size},
    )


def MaxoutWindowEncoder(width: int, window_size: int = 1, maxout_pieces: int = 3, depth: int = 4) -> Model:
    blocks = [_residual_block(width, window_size, maxout_pieces) for _ in range(depth)]

    def init(model: Model, X=None, Y=None):
        for blk in blocks:
            blk.initialize()

    def forward(model: Model, Xm, is_train: bool):
        X, mask = Xm
        callbacks = []
        for blk in blocks:
            X, bp = blk((X, mask), is_train)
            callbacks.append(bp)

        def backprop(dY):
            for bp in reversed(callbacks):
                dY = bp(dY)
            return dY

        return X, backprop

    return Model(
        "maxout_window_encoder", forward, init=init, dims={"nO": width, "nI": width},
        layers=blocks, attrs={"receptive_field": window_size * depth, "depth": depth},
    )


# ---- Tok2Vec ----------------------------------------------------------------
def Tok2Vec(embed: Model, encode: Model) -> Model:
    def init(model: Model, X=None, Y=None):
        embed.initialize()
        encode.initialize()

    def forward(model: Model, batch: TokenBatch, is_train: bool):
        E, bp_embed = embed(batch, is_train)
        Y, bp_encode = encode((E, batch.mask), is_train)

        def backprop(dY):
            bp_embed(bp_encode(dY))
            return None

        return Y, backprop

    return Model(
        "tok2vec", forward, init=init, dims={"nO": embed.get_dim("nO")},
        layers=[embed, encode], refs={"embed": embed, "encode": encode},
    )


def HashEmbedCNN(
    width: int, depth: int, embed_size: int = 2000, window_size: int = 1, maxout_pieces: int = 3,
    subword_features: bool = True, pretrained_vectors: Any = None,
) -> Model:
    if subword_features:
        attrs = ["NORM", "PREFIX", "SUFFIX", "SHAPE"]
        rows = [embed_size, embed_size // 2, embed_size // 2, embed_size // 2]
    else:
        attrs, rows = ["NORM"], [embed_size]
    return Tok2Vec(
        MultiHashEmbed(width, attrs, rows, include_static_vectors=bool(pretrained_vectors)),
        MaxoutWindowEncoder(width, window_size, maxout_pieces, depth),
    )


def Tok2VecListener(width: int, upstream: str = "*") -> Model:
    """Stand-in for a shared upstream ``tok2vec`` component.  The component
    pushes ``(outputs, backprop)`` for the current batch via ``receive``; our
    backprop hands the gradient back (the upstream component sums gradients of
    all its listeners and runs its own backward once)."""

    def forward(model: Model, batch: TokenBatch, is_train: bool):
        store = model.attrs["_store"]
        if store.get("batch_id") != id(batch) or "outputs" not in store:
            raise RuntimeError(
                "Tok2VecListener: no upstream output for this batch - is the 'tok2vec' "
                "component earlier in the pipeline (or listed in annotating_components)?"
            )
        outputs = store["outputs"]
        bp = store.get("backprop")

        def backprop(dY):
            if bp is not None:
                bp(dY)
            return None

        return outputs, backprop

    m = Model(
        "tok2vec_listener", forward, dims={"nO": width},
        attrs={"upstream": upstream, "_store": {}},
    )

    def receive(batch, outputs, backprop) -> None:
        m.attrs["_store"] = {"batch_id": id(batch), "outputs": outputs, "backprop": backprop}

    m.attrs["receive"] = receive
    return m


# ---- Linear / Softmax -------------------------------------------------------
def Linear(nO: Optional[int] = None, nI: Optional[int] = None, *, init_zero: bool = False, name: str = "linear") -> Model:
    def init(model: Model, X=None, Y=None):
        if X is not None and model.has_dim("nI") is None:
            model.set_dim("nI", int(X.shape[1]))
        if Y is not None and model.has_dim("nO") is None:
            model.set_dim("nO", int(Y.shape[1]))
        nO_, nI_ = model.get_dim("nO"), model.get_dim("nI")
        if model.has_param("W") is not True:
            if init_zero:
                model.set_param("W", model.ops.alloc((nO_, nI_)))
            else:
                model.set_param("W", model.ops.glorot_uniform((nO_, nI_), nI_, nO_, _init_gen))
            model.set_param("b", model.ops.alloc((nO_,)))

    def forward(model: Model, X, is_train):
        W, b = model.get_param("W"), model.get_param("b")
        Y = model.ops.linear(X, W, b)

        def backprop(dY):
            if getattr(model.ops, "fused", False):
                # kernels accumulate straight into the flat gradient bucket (inc_grad recognises the buffer)
                go = {"W": model.grad_buffer("W"), "b": model.grad_buffer("b")}
                dX, dW, db = model.ops.linear_backward(dY, X, W, grad_out=go)
            else:
                dX, dW, db = model.ops.linear_backward(dY, X, W)
            model.inc_grad("W", dW)
            model.inc_grad("b", db)
            return dX

        return Y, backprop

    return Model(name, forward, init=init, dims={"nO": nO, "nI": nI}, params={"W": None, "b": None})


def Softmax(nO: Optional[int] = None, nI: Optional[int] = None) -> Model:
    """Zero-initialised affine + softmax.  As in thinc, ``backprop`` expects the
    gradient w.r.t. the *logits* (``probs - target``)."""

    def init(model: Model, X=None, Y=None):
        if X is not None and model.has_dim("nI") is None:
            model.set_dim("nI", int(X.shape[1]))
        if Y is not None and model.has_dim("nO") is None:
            model.set_dim("nO", int(Y.shape[1]))
        nO_, nI_ = model.get_dim("nO"), model.get_dim("nI")
        if model.has_param("W") is not True:
            model.set_param("W", model.ops.alloc((nO_, nI_)))
            model.set_param("b", model.ops.alloc((nO_,)))

    def forward(model: Model, X, is_train):
        W, b = model.get_param("W"), model.get_param("b")
        logits = model.ops.linear(X, W, b)
        P = model.ops.softmax(logits)

        def backprop(d_logits):
            dX, dW, db = model.ops.linear_backward(d_logits, X, W)
            model.inc_grad("W", dW)
            model.inc_grad("b", db)
            return dX

        return P, backprop

    return Model("softmax", forward, init=init, dims={"nO": nO, "nI": nI}, params={"W": None, "b": None})
