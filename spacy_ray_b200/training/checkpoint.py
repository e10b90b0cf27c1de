"""Checkpoint / resume.

The reference defines ``Worker.save_checkpoint`` but never calls it and drops
``--output``.  Here: rank 0 writes the pipeline directory
(``model-best`` / ``model-last``: ``config.cfg``, ``meta.json``, ``<pipe>/model``,
``<pipe>/cfg``); every rank additionally writes the optimizer state of the
parameter keys it *owns* to ``optim/rank{r}-of{n}.pt`` keyed by *structural*
key (component name, node index in ``walk()`` order, param name), so a resumed
run re-derives ownership with the same ``divide_params`` and refuses a
world-size mismatch unless told to re-shard."""
from __future__ import annotations

import json
from pathlib import Path
from typing import Any, Dict, Iterable, Optional, Tuple

import torch

KeyT = Tuple[int, str]


def structural_keys(nlp) -> Dict[KeyT, Tuple[str, int, str]]:
    """``(node.id, name) -> (component, walk_index, name)``; stable across
    processes regardless of the absolute id values."""
    out: Dict[KeyT, Tuple[str, int, str]] = {}
    for cname, comp in nlp.components:
        model = getattr(comp, "model", None)
        if model is None:
            continue
        for i, node in enumerate(model.walk()):
            for pname in node.param_names:
                out.setdefault((node.id, pname), (cname, i, pname))
    return out


def save_optimizer_shard(path: Path, nlp, optimizer, owned_keys: Iterable[KeyT], *, rank: int, world_size: int,
                         extra: Optional[Dict[str, Any]] = None, master: Optional[Dict[KeyT, torch.Tensor]] = None) -> Path:
    """``master``: fp32 master weights of the owned keys when the model weights are kept in lower
    precision (bf16) - without them a resumed run would restart from bf16-rounded weights."""
    path = Path(path) / "optim"
    path.mkdir(parents=True, exist_ok=True)
    skeys = structural_keys(nlp)
    state = optimizer.state_dict(keys=list(owned_keys))

    def conv(d):
        return None if d is None else {skeys[k]: v for k, v in d.items() if k in skeys}

    own = set(owned_keys)
    mstate = dict(state.get("master") or {})
    if master:
        mstate.update({k: v.detach().to("cpu", torch.float32) for k, v in master.items() if k in own})
    blob = {
        "rank": rank, "world_size": world_size,
        "mom1": conv(state["mom1"]), "mom2": conv(state["mom2"]), "nr_update": conv(state["nr_update"]),
        "averages": conv(state["averages"]), "master": conv(mstate), "step": state["step"], "hyper": state["hyper"],
        "extra": extra or {},
    }
    out = path / f"rank{rank}-of{world_size}.pt"
    tmp = path / f".rank{rank}-of{world_size}.pt.tmp"
    torch.save(blob, tmp)
    tmp.replace(out)                      # atomic: a crash mid-save never leaves a truncated shard
    if rank == 0:                         # shards of an earlier run with a different world size are stale
        for f in path.glob("rank*-of*.pt"):
            try:
                if int(f.stem.split("-of")[1]) != world_size:
                    f.unlink()
            except (ValueError, OSError):
                pass
    return out


def load_optimizer_shards(path: Path, nlp, optimizer, owned_keys: Iterable[KeyT], *, rank: int, world_size: int,
                          allow_reshard: bool = True) -> Dict[str, Any]:
    """Load the optimizer state for ``owned_keys``.  This rank's own file is read first; if it does
    not cover every owned key (different world size, or the same world size with a different
    ``shard_balance``) and ``allow_reshard``, the remaining shards of the newest complete set are
    scanned too.  Returns ``{"extra": ..., "master": {key: fp32 tensor}, "nr_update": int}``."""
    path = Path(path) / "optim"
    files = sorted(path.glob("rank*-of*.pt"))
    if not files:
        raise FileNotFoundError(f"No optimizer shards under {path}")
    by_ws: Dict[int, list] = {}
    for f in files:
        try:
            by_ws.setdefault(int(f.stem.split("-of")[1]), []).append(f)
        except ValueError:
            continue
    complete = {ws: fs for ws, fs in by_ws.items() if len(fs) == ws}
    if world_size in complete:
        saved_ws = world_size
    elif complete:
        saved_ws = max(complete, key=lambda ws: max(f.stat().st_mtime for f in complete[ws]))
    else:
        saved_ws = max(by_ws, key=lambda ws: len(by_ws[ws]))
    if saved_ws != world_size and not allow_reshard:
        raise ValueError(f"Checkpoint was written with world_size={saved_ws}, now {world_size}")
    group = by_ws[saved_ws]
    mine = path / f"rank{rank}-of{world_size}.pt"
    wanted = ([mine] if (saved_ws == world_size and mine in group) else []) + [f for f in group if f != mine]
    by_struct = {v: k for k, v in structural_keys(nlp).items()}
    owned = set(owned_keys)
    state: Dict[str, Any] = {"mom1": {}, "mom2": {}, "nr_update": {}, "averages": None, "master": {}, "step": 0}
    extra: Dict[str, Any] = {}
    for i, f in enumerate(wanted):
        if i > 0 and (not allow_reshard or owned <= set(state["nr_update"])):
            break                          # own file covered everything (the common case)
        blob = torch.load(f, map_location="cpu", weights_only=False)
        state["step"] = max(state["step"], int(blob.get("step", 0)))
        extra = blob.get("extra", {}) or extra
        for fld in ("mom1", "mom2", "nr_update", "master"):
            for skey, v in (blob.get(fld) or {}).items():
                key = by_struct.get(tuple(skey))
                if key is not None and key in owned and key not in state[fld]:
                    state[fld][key] = v
        if blob.get("averages"):
            state["averages"] = state["averages"] or {}
            for skey, v in blob["averages"].items():
                key = by_struct.get(tuple(skey))
                if key is not None and key in owned:
                    state["averages"].setdefault(key, v)
    master = state.pop("master")
    optimizer.load_state_dict(state)
    nr = max(state["nr_update"].values()) if state["nr_update"] else 0
    return {"extra": extra, "master": master, "nr_update": int(nr)}


def save_pipeline(nlp, path: Path, *, training_cfg: Optional[Dict[str, Any]] = None, info: Optional[Dict[str, Any]] = None,
                  before_to_disk=None) -> None:
    from .loop import update_meta

    path = Path(path)
    if training_cfg is not None and info is not None:
        frozen = training_cfg.get("frozen_components", []) or []
        with nlp.select_pipes(disable=frozen):
            update_meta(training_cfg, nlp, info)
    target = before_to_disk(nlp) if before_to_disk else nlp
    target.to_disk(path)
    if info is not None:
        slim = {k: v for k, v in info.items() if k in ("epoch", "step", "score", "words", "seconds")}
        (path / "training_state.json").write_text(json.dumps(slim, default=str))
