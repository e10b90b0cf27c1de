"""Static (pretrained) word vectors: the ``include_static_vectors = true`` branch of
``spacy.MultiHashEmbed`` and spaCy's ``StaticVectors`` layer.

Upstream the table lives in ``nlp.vocab.vectors`` (keys = ORTH hashes, ``key2row``) and the
layer computes ``vectors[rows] @ W^T`` with a learned ``W (nO, nM)``; the table itself is not
trained.  Here the table is a process-level ``Vectors`` object (set by ``[initialize] vectors =
path`` / ``Language.load_vectors``), keyed by spaCy-compatible ORTH hashes
(``training.docbin.hash_string`` = MurmurHash64A seed 1, so a table exported from spaCy with its
own keys works as is).  ``Language.make_batch`` resolves each token's row on the host and ships
``vec_rows`` with the batch; the projection and its gradient run on the backend's GEMM path.
"""
from __future__ import annotations

from pathlib import Path
from typing import Dict, Iterable, Optional, Sequence, Union

import numpy as np
import torch

from .model import Model

_CURRENT: Optional["Vectors"] = None


class Vectors:
    def __init__(self, data: np.ndarray, keys: Optional[Sequence[int]] = None, words: Optional[Sequence[str]] = None):
        from ..training.docbin import hash_string

        self.data = np.ascontiguousarray(data, dtype=np.float32)
        if keys is None:
            if words is None:
                raise ValueError("Vectors needs `keys` (ORTH hashes) or `words`")
            keys = [hash_string(w) for w in words]
        if len(keys) != self.data.shape[0]:
            raise ValueError(f"Vectors: {len(keys)} keys for {self.data.shape[0]} rows")
        self.keys = np.asarray(keys, dtype=np.uint64)
        self.key2row: Dict[int, int] = {int(k): i for i, k in enumerate(self.keys.tolist())}
        self._device_cache: Dict[tuple, torch.Tensor] = {}

    @property
    def shape(self):
        return self.data.shape

    def rows_for(self, words: Iterable[str]) -> np.ndarray:
        from ..training.docbin import hash_string

        get = self.key2row.get
        return np.asarray([get(hash_string(w), -1) for w in words], dtype=np.int64)

    def table(self, device: torch.device, dtype: torch.dtype) -> torch.Tensor:
        key = (str(device), dtype)
        t = self._device_cache.get(key)
        if t is None:
            t = self._device_cache[key] = torch.from_numpy(self.data).to(device=device, dtype=dtype)
        return t

    # ---- (de)serialisation: npz with `keys` + `data` ------------------------------------------
    def to_disk(self, path: Union[str, Path]) -> None:
        path = Path(path)
        path.parent.mkdir(parents=True, exist_ok=True)
        np.savez(path, keys=self.keys, data=self.data)

    @classmethod
    def from_disk(cls, path: Union[str, Path]) -> "Vectors":
        path = Path(path)
        if path.is_dir():
            for cand in ("vectors.npz", "vocab/vectors.npz"):
                if (path / cand).exists():
                    path = path / cand
                    break
            else:
                raise FileNotFoundError(f"no vectors.npz under {path}")
        if path.suffix == ".npz":
            z = np.load(path, allow_pickle=False)
            if "keys" in z:
                return cls(z["data"], keys=z["keys"].tolist())
            return cls(z["data"], words=[str(w) for w in z["words"].tolist()])
        # word2vec / GloVe / fastText text format: optional "n dim" header, then "word v1 ... vd"
        words, rows = [], []
        with path.open("r", encoding="utf8") as fh:
            first = fh.readline().rstrip("\n").split(" ")
            if not (len(first) == 2 and all(x.isdigit() for x in first)):
                words.append(first[0])
                rows.append(np.asarray(first[1:], dtype=np.float32))
            for line in fh:
                parts = line.rstrip("\n").split(" ")
                if len(parts) > 2:
                    words.append(parts[0])
                    rows.append(np.asarray(parts[1:], dtype=np.float32))
        if not rows:
            raise ValueError(f"no vectors found in {path}")
        return cls(np.stack(rows), words=words)


def set_vectors(v: Optional[Vectors]) -> None:
    global _CURRENT
    _CURRENT = v


def get_vectors() -> Optional[Vectors]:
    return _CURRENT


def StaticVectors(nO: int, nM: Optional[int] = None) -> Model:
    """TokenBatch -> (Tp, nO): ``vectors[vec_rows] @ W^T`` (rows of OOV / pad tokens are zero)."""
    from .layers import _init_gen

    def init(model: Model, X=None, Y=None):
        if model.has_dim("nM") is None:
            v = get_vectors()
            if v is None:
                raise ValueError("StaticVectors: no vectors table loaded - set [initialize] vectors = <path> "
                                 "(.npz with keys/words + data, or word2vec text format)")
            model.set_dim("nM", int(v.shape[1]))
        if model.has_param("W") is not True:
            nO_, nM_ = model.get_dim("nO"), model.get_dim("nM")
            model.set_param("W", model.ops.glorot_uniform((nO_, nM_), nM_, nO_, _init_gen))

    def forward(model: Model, batch, is_train: bool):
        ops = model.ops
        v = get_vectors()
        rows = batch.extra.get("vec_rows")
        if v is None or rows is None:
            raise RuntimeError("StaticVectors: the batch carries no `vec_rows` (vectors not loaded before make_batch?)")
        W = model.get_param("W")
        table = v.table(ops.device, W.dtype)
        ok = (rows >= 0).to(W.dtype).unsqueeze(1)
        V = (table.index_select(0, rows.clamp(min=0)) * ok).contiguous()
        Y = ops.linear(V, W, None)

        def backprop(dY):
            _dx, dW, _db = ops.linear_backward(dY, V, W, need_dX=False, need_db=False)
            model.inc_grad("W", dW)
            return None

        return Y, backprop

    return Model("staticvectors", forward, init=init, dims={"nO": nO, "nM": nM}, params={"W": None})
