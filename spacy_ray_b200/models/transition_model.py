"""``spacy.TransitionBasedParser``: the shared model behind NER and the parser.

Structure:

    tok2vec (Tp, w) -> Linear (Tp, h) -> lower "PrecomputableAffine":
        Yf = X @ W_lower^T                      (Tp, nF, nO*nP)   once per batch
    per transition step, per live state with feature token rows ``ids[nF]``:
        hidden = maxout_nP( b + sum_f (Yf[ids[f], f] or pad[f] if ids[f] < 0) )
        scores = hidden @ W_upper^T + b_upper    (n_actions,)
        invalid actions masked; training loss pushes mass onto zero-cost actions;
        states advance by the best-scoring valid action.

The per-step loop is sequential only through the tiny state; everything in the
backward pass is batched over all recorded steps (two GEMMs + one scatter).
``ops.transition_steps`` runs the whole forward loop - on the B200 backend as
ONE persistent kernel with a warp per doc (no host round trip per step,
SURVEY.md K7); the PyTorch reference below is the spec it is tested against.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Dict, List, Optional

import torch

from ..nn.batch import TokenBatch
from ..nn.layers import Linear, _init_gen, _noop_forward
from ..nn.model import Model
from .transitions import ArcEagerSystem, BiluoSystem


@dataclass
class TransitionGold:
    """Gold annotations in the form each system's oracle wants."""
    # BILUO: flat per-token gold action ids in doc order (unpadded), -1 = missing
    actions: Optional[torch.Tensor] = None
    offsets: Optional[torch.Tensor] = None      # (B,) int64 offset of each doc in ``actions``
    # arc-eager
    heads: Optional[List[List[int]]] = None
    labels: Optional[List[List[int]]] = None
    heads_flat: Optional[torch.Tensor] = None   # (T,) int32 doc-relative gold heads (self = root, -1 = missing)
    labels_flat: Optional[torch.Tensor] = None  # (T,) int32 gold labels (-1 = unknown)
    # training only: advance each state by the oracle's action (BILUO: the single zero-cost action when
    # there is one; arc-eager: the first minimum-cost action) instead of the model's arg-max.  Makes
    # two implementations follow identical trajectories, so their losses / gradients can be compared
    # tightly (tests); the default (False) is spaCy's behaviour.
    teacher_forced: bool = False


@dataclass
class TransitionModelOutput:
    loss: Any = 0.0                       # float or 0-d tensor
    histories: Optional[List[List[int]]] = None     # per doc action sequence (host)
    actions_flat: Optional[torch.Tensor] = None     # BILUO: (T,) predicted action per token, doc order
    states: Optional[list] = None                   # arc-eager final states (host loop)
    heads_flat: Optional[torch.Tensor] = None       # arc-eager on device: (T,) predicted heads (doc-relative)
    labels_flat: Optional[torch.Tensor] = None      # ... and labels
    n_steps: int = 0


def PrecomputableAffine(nO: int, nI: int, nF: int, nP: int) -> Model:
    """Parameter holder: ``W (nF, nO, nP, nI)``, ``b (nO, nP)``, ``pad (1, nF, nO, nP)``."""

    def init(model: Model, X=None, Y=None):
        if model.has_param("W") is not True:
            nF_, nO_, nP_, nI_ = (model.get_dim(d) for d in ("nF", "nO", "nP", "nI"))
            model.set_param("W", model.ops.glorot_uniform((nF_, nO_, nP_, nI_), nI_ * nF_, nO_ * nP_, _init_gen))
            model.set_param("b", model.ops.alloc((nO_, nP_)))
            model.set_param("pad", model.ops.uniform((1, nF_, nO_, nP_), -0.05, 0.05, _init_gen))

    return Model(
        "precomputable_affine", _noop_forward, init=init,
        dims={"nO": nO, "nI": nI, "nF": nF, "nP": nP}, params={"W": None, "b": None, "pad": None},
    )


def build_transition_model(
    tok2vec: Model,
    state_type: str = "ner",
    extra_state_tokens: bool = False,
    hidden_width: int = 64,
    maxout_pieces: int = 2,
    use_upper: bool = True,
    nO: Optional[int] = None,
) -> Model:
    if state_type == "ner":
        nF = 3
    elif state_type == "parser":
        nF = 8
    else:
        raise ValueError(f"Unknown state_type {state_type!r} (expected 'ner' or 'parser')")
    if extra_state_tokens:
        raise NotImplementedError("extra_state_tokens=true is not supported")
    if not use_upper:
        raise NotImplementedError("use_upper=false is not supported")
    width = tok2vec.get_dim("nO")
    t2v_linear = Linear(hidden_width, width, name="tok2vec_to_hidden")
    lower = PrecomputableAffine(hidden_width, hidden_width, nF, maxout_pieces)
    upper = Linear(nO, hidden_width, init_zero=True, name="upper")

    def init(model: Model, X=None, Y=None):
        tok2vec.initialize()
        t2v_linear.initialize()
        lower.initialize()
        if upper.has_dim("nO") is None:
            if model.has_dim("nO") is None:
                raise ValueError("Transition model: number of actions (nO) not set before initialize")
            upper.set_dim("nO", model.get_dim("nO"))
        upper.initialize()

    def run(batch: TokenBatch, system, gold: Optional[TransitionGold], is_train: bool) -> TransitionModelOutput:
        ops = model.ops
        tokvecs, bp_t2v = tok2vec(batch, is_train)
        H, bp_lin = t2v_linear(tokvecs, is_train)
        Wl = lower.get_param("W")
        nF_, nO_, nP_, nI_ = Wl.shape
        Wl2 = Wl.reshape(nF_ * nO_ * nP_, nI_)
        Yf = ops.linear(H, Wl2, None)                                   # (Tp, nF*nO*nP)
        params = {
            "pad": lower.get_param("pad").reshape(nF_, nO_ * nP_),
            "b": lower.get_param("b").reshape(nO_ * nP_),
            "Wu": upper.get_param("W"), "bu": upper.get_param("b"),
            "nF": nF_, "nO": nO_, "nP": nP_,
        }
        rec = transition_steps(ops, system, Yf, params, batch, gold, is_train)
        out = TransitionModelOutput(
            loss=rec.get("loss", 0.0), histories=rec.get("histories"),
            actions_flat=rec.get("actions_flat"), states=rec.get("states"), n_steps=rec["n_steps"],
            heads_flat=rec.get("arc_heads"), labels_flat=rec.get("arc_labels"),
        )
        if not is_train or rec["n_steps"] == 0:
            return out
        # ---- batched backward over all recorded steps --------------------
        fused = getattr(ops, "fused", False)
        go = None
        if fused:         # kernels accumulate straight into the flat gradient bucket
            go = {"Wu": upper.grad_buffer("W"), "bu": upper.grad_buffer("b"), "b": lower.grad_buffer("b"),
                  "pad": lower.grad_buffer("pad")}
        g = transition_backward(ops, rec, params, batch.n_rows, grad_out=go)
        upper.inc_grad("W", g["dWu"])
        upper.inc_grad("b", g["dbu"])
        lower.inc_grad("b", g["db"].reshape(nO_, nP_))
        lower.inc_grad("pad", g["dpad"].reshape(1, nF_, nO_, nP_))
        if fused:
            dH, dWl2, _ = ops.linear_backward(g["dYf"], H, Wl2, need_db=False,
                                              grad_out={"W": lower.grad_buffer("W")})
        else:
            dH, dWl2, _ = ops.linear_backward(g["dYf"], H, Wl2, need_db=False)  # the precompute layer has no bias
        lower.inc_grad("W", dWl2.reshape(nF_, nO_, nP_, nI_))
        bp_t2v(bp_lin(dH))
        return out

    def forward(model_: Model, batch: TokenBatch, is_train: bool):
        raise RuntimeError("Transition models are driven through model.attrs['run'](batch, system, gold, is_train)")

    model = Model(
        "transition_model", forward, init=init, dims={"nO": nO},
        layers=[tok2vec, t2v_linear, lower, upper],
        refs={"tok2vec": tok2vec, "lower": lower, "upper": upper, "t2v_linear": t2v_linear},
        attrs={"state_type": state_type, "nF": nF},
    )
    model.attrs["run"] = run
    return model


# ----------------------------------------------------------------------------
# reference step loop + batched backward (used by TorchOps; B200Ops overrides
# ``transition_steps`` for BILUO with the persistent kernel)
# ----------------------------------------------------------------------------
def transition_steps(ops, system, Yf, params, batch: TokenBatch, gold, is_train) -> Dict[str, Any]:
    fused = getattr(ops, "transition_steps", None)
    if fused is not None:
        rec = fused(system, Yf, params, batch, gold, is_train)
        if rec is not None:
            return rec
    if isinstance(system, BiluoSystem):
        return _biluo_steps_reference(system, Yf, params, batch, gold, is_train)
    if isinstance(system, ArcEagerSystem):
        return _arc_steps_reference(system, Yf, params, batch, gold, is_train)
    raise TypeError(f"Unknown transition system {type(system)}")


def _state_scores(Yf3, params, feats):
    """feats (S, nF) rows (-1 missing) -> hidden (S, nO), which (S, nO), scores (S, A)."""
    nF, nO, nP = params["nF"], params["nO"], params["nP"]
    pre = params["b"].to(torch.float32).unsqueeze(0).expand(feats.shape[0], -1).clone()
    pad = params["pad"].to(torch.float32)
    for f in range(nF):
        rows = feats[:, f]
        ok = rows >= 0
        vals = Yf3[rows.clamp(min=0), f].to(torch.float32)
        pre += torch.where(ok.unsqueeze(1), vals, pad[f].unsqueeze(0))
    hid, which = pre.view(-1, nO, nP).max(dim=2)
    scores = hid @ params["Wu"].to(torch.float32).t() + params["bu"].to(torch.float32)
    return hid, which, scores


def _loss_grad(scores, valid, gold_mask):
    """d = softmax_valid(scores) - softmax_{valid & gold}(scores); rows with an
    empty gold set get d = 0.  Returns (d, loss=sum d^2)."""
    neg = torch.finfo(torch.float32).min
    sv = torch.where(valid, scores, torch.full_like(scores, neg))
    pv = torch.softmax(sv, dim=1) * valid.to(torch.float32)
    gm = valid & gold_mask
    has_gold = gm.any(dim=1, keepdim=True)
    sg = torch.where(gm, scores, torch.full_like(scores, neg))
    pg = torch.softmax(sg, dim=1) * gm.to(torch.float32)
    d = (pv - pg) * has_gold.to(torch.float32)
    return d


def _biluo_steps_reference(system: BiluoSystem, Yf, params, batch, gold, is_train) -> Dict[str, Any]:
    dev = Yf.device
    nF, nO, nP = params["nF"], params["nO"], params["nP"]
    Yf3 = Yf.view(Yf.shape[0], nF, nO * nP)
    starts = batch.doc_starts.to(torch.int64)
    st = system.batch_init(batch.doc_lens)
    B = batch.n_docs
    T = batch.n_tokens
    max_len = max(batch.lengths) if batch.lengths else 0
    tok_off = torch.zeros(B, dtype=torch.int64, device=dev)
    if B:
        tok_off[1:] = torch.cumsum(batch.doc_lens.to(torch.int64), 0)[:-1]
    actions_flat = torch.zeros(T, dtype=torch.int64, device=dev)
    feats_l, which_l, hid_l, d_l = [], [], [], []
    loss = torch.zeros((), dtype=torch.float32, device=dev)
    have_gold = gold is not None and gold.actions is not None
    for _k in range(max_len):
        active = system.batch_active(st)
        idx = torch.nonzero(active, as_tuple=False).squeeze(1)
        if idx.numel() == 0:
            break
        feats = system.batch_features(st, starts)[idx]
        valid = system.batch_valid(st)[idx]
        hid, which, scores = _state_scores(Yf3, params, feats)
        neg = torch.finfo(torch.float32).min
        guess = torch.where(valid, scores, torch.full_like(scores, neg)).argmax(dim=1)
        if is_train and have_gold:
            ga = system.batch_gold_action(st, gold.actions, gold.offsets)[idx]
            A = system.n_actions
            onehot = torch.arange(A, device=dev).unsqueeze(0) == ga.unsqueeze(1)
            gold_mask = torch.where((ga < 0).unsqueeze(1), valid, onehot)
            # a single gold action that is not valid (inconsistent gold) -> no constraint
            gold_mask = torch.where((gold_mask & valid).any(dim=1, keepdim=True), gold_mask, valid)
            d = _loss_grad(scores, valid, gold_mask) / float(idx.numel())
            loss = loss + (d * d).sum()
            feats_l.append(feats)
            which_l.append(which.to(torch.uint8))
            hid_l.append(hid)
            d_l.append(d)
        actions_flat[(tok_off + st["i"])[idx]] = guess              # the record is always the model's prediction
        if is_train and have_gold and getattr(gold, "teacher_forced", False):
            ga_ok = (ga >= 0) & valid.gather(1, ga.clamp(min=0).unsqueeze(1)).squeeze(1)
            guess = torch.where(ga_ok, ga, guess)
        full_guess = torch.zeros(B, dtype=torch.int64, device=dev)
        full_guess[idx] = guess
        st = system.batch_apply(st, full_guess, gold.actions if have_gold else None, gold.offsets if have_gold else None)
    rec: Dict[str, Any] = {"actions_flat": actions_flat, "n_steps": 0, "loss": loss}
    if feats_l:
        rec.update({
            "feats": torch.cat(feats_l), "which": torch.cat(which_l), "hid": torch.cat(hid_l),
            "d_scores": torch.cat(d_l),
        })
        rec["n_steps"] = int(rec["feats"].shape[0])
    return rec


def _arc_steps_reference(system: ArcEagerSystem, Yf, params, batch, gold, is_train) -> Dict[str, Any]:
    dev = Yf.device
    nF, nO, nP = params["nF"], params["nO"], params["nP"]
    Yf3 = Yf.view(Yf.shape[0], nF, nO * nP)
    states = [system.init_state(n) for n in batch.lengths]
    starts = batch.starts
    A = system.n_actions
    feats_l, which_l, hid_l, d_l = [], [], [], []
    loss = torch.zeros((), dtype=torch.float32, device=dev)
    have_gold = gold is not None and gold.heads is not None
    guard = 0
    max_steps = 4 * (max(batch.lengths) if batch.lengths else 0) + 8
    while guard < max_steps:
        guard += 1
        live = [d for d, s in enumerate(states) if not s.is_final]
        if not live:
            break
        feats_host = []
        valid_host = []
        for d in live:
            s = states[d]
            feats_host.append([(starts[d] + t) if t >= 0 else -1 for t in system.features(s)])
            valid_host.append(system.valid(s))
        feats = torch.tensor(feats_host, dtype=torch.int64, device=dev)
        valid = torch.tensor(valid_host, dtype=torch.bool, device=dev)
        hid, which, scores = _state_scores(Yf3, params, feats)
        neg = torch.finfo(torch.float32).min
        guess = torch.where(valid, scores, torch.full_like(scores, neg)).argmax(dim=1).tolist()
        if is_train and have_gold:
            cost_host = [system.costs(states[d], gold.heads[d], gold.labels[d]) for d in live]
            costs = torch.tensor(cost_host, dtype=torch.int64, device=dev)
            masked = torch.where(valid, costs, torch.full_like(costs, 1 << 20))
            gold_mask = valid & (masked == masked.min(dim=1, keepdim=True).values)
            # per-step gradient scale 1 / #docs (constant per batch; the device kernel cannot know
            # how many docs are still live at a given step without a grid-wide sync)
            d = _loss_grad(scores, valid, gold_mask) / float(len(states))
            loss = loss + (d * d).sum()
            feats_l.append(feats)
            which_l.append(which.to(torch.uint8))
            hid_l.append(hid)
            d_l.append(d)
            if getattr(gold, "teacher_forced", False):
                guess = gold_mask.to(torch.int8).argmax(dim=1).tolist()      # first minimum-cost action
        for d, a in zip(live, guess):
            system.apply(states[d], int(a))
    rec: Dict[str, Any] = {"states": states, "n_steps": 0, "loss": loss,
                           "histories": [s.history for s in states]}
    if feats_l:
        rec.update({
            "feats": torch.cat(feats_l), "which": torch.cat(which_l), "hid": torch.cat(hid_l),
            "d_scores": torch.cat(d_l),
        })
        rec["n_steps"] = int(rec["feats"].shape[0])
    return rec


def transition_backward(ops, rec, params, n_rows: int, grad_out=None) -> Dict[str, torch.Tensor]:
    """Gradients of everything downstream of ``Yf`` from the step records.

    dWu = d^T hid; dbu = sum d; d_hid = d Wu; dPre = route d_hid to the winning
    piece; db = sum dPre; dYf[ids[f], f] += dPre (or dpad[f] if ids[f] < 0)."""
    fused = getattr(ops, "transition_backward", None)
    if fused is not None:
        out = fused(rec, params, n_rows, grad_out=grad_out) if grad_out is not None else fused(rec, params, n_rows)
        if out is not None:
            return out
    nF, nO, nP = params["nF"], params["nO"], params["nP"]
    d = rec["d_scores"].to(torch.float32)
    hid = rec["hid"].to(torch.float32)
    feats = rec["feats"]
    dWu = d.t() @ hid
    dbu = d.sum(dim=0)
    d_hid = d @ params["Wu"].to(torch.float32)
    S = d.shape[0]
    dPre = torch.zeros((S, nO, nP), dtype=torch.float32, device=d.device)
    dPre.scatter_(2, rec["which"].to(torch.int64).unsqueeze(2), d_hid.unsqueeze(2))
    dPre = dPre.view(S, nO * nP)
    db = dPre.sum(dim=0)
    dYf = torch.zeros((n_rows, nF, nO * nP), dtype=torch.float32, device=d.device)
    dpad = torch.zeros((nF, nO * nP), dtype=torch.float32, device=d.device)
    for f in range(nF):
        rows = feats[:, f]
        ok = rows >= 0
        okf = ok.to(torch.float32).unsqueeze(1)
        dYf[:, f].index_add_(0, rows.clamp(min=0), dPre * okf)
        dpad[f] = (dPre * (1.0 - okf)).sum(dim=0)
    return {"dWu": dWu, "dbu": dbu, "db": db, "dpad": dpad, "dYf": dYf.view(n_rows, nF * nO * nP)}
