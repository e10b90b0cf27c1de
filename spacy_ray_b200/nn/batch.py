"""``TokenBatch``: the padded-ragged device representation of a batch of docs.

Layout (see ``ops/torch_ops.py`` module docstring): row 0 is a zero pad row,
doc ``d`` occupies rows ``[starts[d], starts[d] + lens[d])`` and is followed
by one pad row; rows past the last pad row (capacity padding, used so CUDA
graphs see a fixed shape) are pad rows too.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np
import torch


def padded_rows(lengths: Sequence[int]) -> int:
    return int(sum(lengths)) + len(lengths) + 1


@dataclass
class TokenBatch:
    attrs: torch.Tensor            # (Tp, n_attr) int64 - hashed lexical attribute ids
    mask: torch.Tensor             # (Tp, 1) float - 1.0 on real tokens
    doc_starts: torch.Tensor       # (B,) int32 row of each doc's first token
    doc_lens: torch.Tensor         # (B,) int32
    lengths: List[int]             # host copy of doc_lens
    starts: List[int]              # host copy of doc_starts
    n_tokens: int
    n_rows: int                    # Tp (>= padded_rows(lengths))
    token_rows: Optional[torch.Tensor] = None   # (T,) int64 rows of the real tokens, doc order
    extra: dict = field(default_factory=dict)

    @property
    def n_docs(self) -> int:
        # device-resident batches (engine.Trainer) carry no host-side lengths: their doc count is
        # the static doc capacity of the staging buffer
        return len(self.lengths) if self.lengths else int(self.doc_lens.shape[0])

    @property
    def device(self) -> torch.device:
        return self.attrs.device

    def rows_of_tokens(self) -> torch.Tensor:
        if self.token_rows is None:
            idx = np.concatenate(
                [np.arange(s, s + n, dtype=np.int64) for s, n in zip(self.starts, self.lengths)]
            ) if self.lengths else np.zeros((0,), dtype=np.int64)
            self.token_rows = torch.from_numpy(idx).to(self.attrs.device)
        return self.token_rows

    def unpad(self, X: torch.Tensor) -> torch.Tensor:
        """(Tp, ...) -> (T, ...) keeping only real-token rows, doc order."""
        return X.index_select(0, self.rows_of_tokens())

    def pad(self, X: torch.Tensor) -> torch.Tensor:
        """(T, ...) -> (Tp, ...) with zeros in pad rows."""
        out = torch.zeros((self.n_rows,) + tuple(X.shape[1:]), dtype=X.dtype, device=X.device)
        out.index_copy_(0, self.rows_of_tokens(), X)
        return out

    def split(self, X: torch.Tensor) -> List[torch.Tensor]:
        """Per-doc views of a padded (Tp, ...) array."""
        return [X[s:s + n] for s, n in zip(self.starts, self.lengths)]


def collate_attrs(
    doc_attrs: Sequence[np.ndarray],
    *,
    capacity_rows: Optional[int] = None,
    out: Optional[np.ndarray] = None,
):
    """Host-side collation: list of per-doc ``(n_i, n_attr)`` uint64 arrays ->
    padded ``(Tp, n_attr)`` int64 array + starts/lens.  ``out`` may be a view of
    a pinned staging buffer (the engine double-buffers these)."""
    lengths = [int(a.shape[0]) for a in doc_attrs]
    n_attr = int(doc_attrs[0].shape[1]) if doc_attrs else 4
    need = padded_rows(lengths)
    Tp = max(need, capacity_rows or 0)
    if out is None:
        out = np.zeros((Tp, n_attr), dtype=np.int64)
    else:
        if out.shape[0] < need:
            raise ValueError(f"collate_attrs: staging buffer has {out.shape[0]} rows, need {need}")
        Tp = out.shape[0]
        out[:] = 0
    starts = []
    row = 1
    for a, n in zip(doc_attrs, lengths):
        starts.append(row)
        if n:
            out[row:row + n] = a.view(np.int64) if a.dtype == np.uint64 else a
        row += n + 1
    return out, starts, lengths, Tp


H2D_BYTES = 0          # running count of bytes staged host -> device (bench.py reads this)
_PIN_RING: dict = {}
_PIN_SLOTS = 4


def to_device(arr: np.ndarray, dev: torch.device, dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    """numpy -> device tensor.  On CUDA the copy goes through a small ring of reusable
    *pinned* staging buffers and is asynchronous on the current stream."""
    global H2D_BYTES
    src = torch.from_numpy(np.ascontiguousarray(arr))
    if dev.type != "cuda":
        out = src.to(dev)
        return out.to(dtype) if dtype is not None else out
    n = src.numel()
    key = (src.dtype, max(1024, 1 << (max(n, 1) - 1).bit_length()))
    ring = _PIN_RING.get(key)
    if ring is None:
        ring = _PIN_RING[key] = {"bufs": [torch.empty(key[1], dtype=src.dtype, pin_memory=True) for _ in range(_PIN_SLOTS)],
                                 "events": [None] * _PIN_SLOTS, "next": 0}
    i = ring["next"]
    ring["next"] = (i + 1) % _PIN_SLOTS
    if ring["events"][i] is not None:
        ring["events"][i].synchronize()          # the previous copy out of this slot has finished
    stage = ring["bufs"][i][:n].view(src.shape)
    stage.copy_(src)
    out = stage.to(dev, non_blocking=True)
    ev = torch.cuda.Event()
    ev.record()
    ring["events"][i] = ev
    H2D_BYTES += n * src.element_size()
    return out.to(dtype) if dtype is not None else out


def make_token_batch(
    doc_attrs: Sequence[np.ndarray],
    device: torch.device | str = "cpu",
    *,
    capacity_rows: Optional[int] = None,
    mask_dtype: torch.dtype = torch.float32,
) -> TokenBatch:
    arr, starts, lengths, Tp = collate_attrs(doc_attrs, capacity_rows=capacity_rows)
    mask = np.zeros((Tp, 1), dtype=np.float32)
    for s, n in zip(starts, lengths):
        mask[s:s + n] = 1.0
    dev = torch.device(device)
    return TokenBatch(
        attrs=to_device(arr, dev),
        mask=to_device(mask, dev, mask_dtype),
        doc_starts=to_device(np.asarray(starts, dtype=np.int32), dev),
        doc_lens=to_device(np.asarray(lengths, dtype=np.int32), dev),
        lengths=lengths,
        starts=starts,
        n_tokens=int(sum(lengths)),
        n_rows=Tp,
    )
