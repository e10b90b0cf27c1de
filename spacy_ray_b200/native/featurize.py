"""ctypes binding for ``csrc/host_runtime.cpp``."""
from __future__ import annotations

import ctypes
import os
from pathlib import Path
from typing import Optional, Sequence

import numpy as np

_LIB_PATH = Path(__file__).parent / "_host_runtime.so"
_lib: Optional[ctypes.CDLL] = None
_tried = False


def _load() -> Optional[ctypes.CDLL]:
    global _lib, _tried
    if _tried:
        return _lib
    _tried = True
    if os.environ.get("SRB_DISABLE_NATIVE") == "1" or not _LIB_PATH.exists():
        return None
    try:
        lib = ctypes.CDLL(str(_LIB_PATH))
        lib.srb_featurize.argtypes = [ctypes.c_char_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p]
        lib.srb_featurize.restype = None
        lib.srb_collate.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64,
                                    ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64]
        lib.srb_collate.restype = ctypes.c_int64
        lib.srb_collate_gold.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
                                         ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64]
        lib.srb_collate_gold.restype = ctypes.c_int64
        lib.srb_abi_version.restype = ctypes.c_int
        if lib.srb_abi_version() != 2:
            return None
        lib.srb_group_rows.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
                                       ctypes.c_int32, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p,
                                       ctypes.c_void_p]
        lib.srb_group_rows.restype = ctypes.c_int64
        _lib = lib
    except OSError:
        _lib = None
    return _lib


def available() -> bool:
    return _load() is not None


def featurize_words(words: Sequence[str]) -> np.ndarray:
    lib = _load()
    assert lib is not None
    n = len(words)
    enc = [w.encode("utf8") for w in words]
    offsets = np.zeros(n + 1, dtype=np.int64)
    if n:
        np.cumsum([len(b) for b in enc], out=offsets[1:])
    buf = b"".join(enc)
    out = np.empty((n, 4), dtype=np.uint64)
    flags = np.zeros(n, dtype=np.uint8)
    lib.srb_featurize(buf, offsets.ctypes.data, n, out.ctypes.data, flags.ctypes.data)
    if flags.any():
        from ..pipeline.doc import lex_attrs

        for i in np.nonzero(flags)[0]:
            a = lex_attrs(words[int(i)])
            out[i, 0], out[i, 1], out[i, 2], out[i, 3] = a[0], a[1], a[2], a[3]
    return out


def collate(store: np.ndarray, doc_off: np.ndarray, ids: np.ndarray, out_attrs: np.ndarray,
            out_mask: np.ndarray, out_starts: np.ndarray, out_lens: np.ndarray) -> int:
    """Gather docs ``ids`` from the corpus-wide ``store`` into padded staging
    buffers (may be pinned-tensor-backed numpy views).  Returns rows used."""
    lib = _load()
    cap = out_attrs.shape[0]
    n_attr = store.shape[1]
    if lib is not None:
        used = lib.srb_collate(store.ctypes.data, doc_off.ctypes.data, n_attr, ids.ctypes.data, len(ids),
                               out_attrs.ctypes.data, out_mask.ctypes.data, out_starts.ctypes.data,
                               out_lens.ctypes.data, cap)
        if used < 0:
            raise ValueError("collate: staging buffer too small")
        return int(used)
    out_attrs[:] = 0
    out_mask[:] = 0
    row = 1
    for d, i in enumerate(ids):
        a, b = int(doc_off[i]), int(doc_off[i + 1])
        n = b - a
        if row + n + 1 > cap:
            raise ValueError("collate: staging buffer too small")
        out_starts[d] = row
        out_lens[d] = n
        out_attrs[row:row + n] = store[a:b]
        out_mask[row:row + n] = 1.0
        row += n + 1
    return row


def collate_gold(store: np.ndarray, doc_off: np.ndarray, ids: np.ndarray, out: np.ndarray, out_off: np.ndarray) -> int:
    lib = _load()
    if lib is not None:
        used = lib.srb_collate_gold(store.ctypes.data, doc_off.ctypes.data, ids.ctypes.data, len(ids),
                                    out.ctypes.data, out_off.ctypes.data, out.shape[0])
        if used < 0:
            raise ValueError("collate_gold: staging buffer too small")
        return int(used)
    pos = 0
    for d, i in enumerate(ids):
        a, b = int(doc_off[i]), int(doc_off[i + 1])
        n = b - a
        out_off[d] = pos
        out[pos:pos + n] = store[a:b]
        pos += n
    return pos


def group_rows(gid: np.ndarray, n_groups: np.ndarray, doc_off: np.ndarray, ids: np.ndarray, rb: int,
               out: np.ndarray, scratch: np.ndarray) -> int:
    """HashEmbed-backward grouping for one batch (see ``srb_group_rows``): per attribute column,
    ``out[c*rb : (c+1)*rb]`` lists the padded-layout rows of the batch's tokens with equal ids
    adjacent; the tail is row 0 (a pad row).  ``gid`` holds dense per-column vocabulary indices."""
    n_cols = gid.shape[1]
    assert gid.dtype == np.int32 and gid.flags.c_contiguous and out.dtype == np.int32 and out.size >= n_cols * rb
    lib = _load()
    if lib is not None:
        n = lib.srb_group_rows(gid.ctypes.data, doc_off.ctypes.data, ids.ctypes.data, len(ids), n_cols,
                               n_groups.ctypes.data, rb, out.ctypes.data, scratch.ctypes.data)
        if n < 0:
            raise ValueError("group_rows: batch does not fit the row capacity")
        return int(n)
    rows, toks, row = [], [], 1
    for i in ids:
        a, b = int(doc_off[i]), int(doc_off[i + 1])
        rows.append(np.arange(row, row + (b - a), dtype=np.int32))
        toks.append(np.arange(a, b))
        row += (b - a) + 1
    rows_a = np.concatenate(rows) if rows else np.zeros((0,), dtype=np.int32)
    toks_a = np.concatenate(toks) if toks else np.zeros((0,), dtype=np.int64)
    n = len(rows_a)
    for c in range(n_cols):
        order = np.argsort(gid[toks_a, c], kind="stable")
        out[c * rb: c * rb + n] = rows_a[order]
        out[c * rb + n: (c + 1) * rb] = 0
    return n
