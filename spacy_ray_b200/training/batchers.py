"""Batchers (spaCy names).  The reference resolves ``T["batcher"]`` from the
config (``/root/reference/spacy_ray/worker.py:170-175``); the default is
``spacy.batch_by_words.v1`` with a compounding size 100 -> 1000."""
from __future__ import annotations

import itertools
from typing import Any, Callable, Iterable, Iterator, List, Optional, Union

from ..config import registry

Sizing = Union[int, Iterator[float], Iterable[float]]


def _size_iter(size: Sizing) -> Iterator[float]:
    if isinstance(size, (int, float)):
        return itertools.repeat(size)
    return iter(size)


@registry.batchers("spacy.batch_by_words.v1")
def configure_minibatch_by_words(size: Sizing, tolerance: float = 0.2, discard_oversize: bool = False,
                                 get_length: Optional[Callable[[Any], int]] = None):
    def batcher(items: Iterable[Any]) -> Iterator[List[Any]]:
        return minibatch_by_words(items, size, tolerance=tolerance, discard_oversize=discard_oversize,
                                  get_length=get_length or len)
    return batcher


@registry.batchers("spacy.batch_by_sequence.v1")
def configure_minibatch(size: Sizing, get_length: Optional[Callable[[Any], int]] = None):
    def batcher(items: Iterable[Any]) -> Iterator[List[Any]]:
        sizes = _size_iter(size)
        it = iter(items)
        while True:
            n = int(next(sizes))
            chunk = list(itertools.islice(it, n))
            if not chunk:
                return
            yield chunk
    return batcher


@registry.batchers("spacy.batch_by_padded.v1")
def configure_minibatch_by_padded_size(size: Sizing, buffer: int = 256, discard_oversize: bool = False,
                                       get_length: Optional[Callable[[Any], int]] = None):
    get_len = get_length or len

    def batcher(items: Iterable[Any]) -> Iterator[List[Any]]:
        sizes = _size_iter(size)
        it = iter(items)
        while True:
            outer = list(itertools.islice(it, buffer))
            if not outer:
                return
            target = int(next(sizes))
            outer.sort(key=get_len)
            batch: List[Any] = []
            longest = 0
            for item in outer:
                n = get_len(item)
                if discard_oversize and n > target:
                    continue
                if batch and max(longest, n) * (len(batch) + 1) > target:
                    yield batch
                    batch, longest = [], 0
                batch.append(item)
                longest = max(longest, n)
            if batch:
                yield batch
    return batcher


def minibatch_by_words(items: Iterable[Any], size: Sizing, tolerance: float = 0.2,
                       discard_oversize: bool = False, get_length: Callable[[Any], int] = len) -> Iterator[List[Any]]:
    """Batches of roughly ``size`` words; a batch may overshoot by up to
    ``tolerance * size`` to avoid a tiny trailing batch; items larger than
    ``size * (1 + tolerance)`` form their own batch (or are dropped)."""
    sizes = _size_iter(size)
    target = float(next(sizes))
    tol = target * tolerance
    batch: List[Any] = []
    overflow: List[Any] = []
    batch_words = 0
    overflow_words = 0
    for item in items:
        n = get_length(item)
        if n > target + tol:
            if not discard_oversize:
                yield [item]
            continue
        if not overflow and batch_words + n <= target:
            batch.append(item)
            batch_words += n
        elif overflow_words + n <= tol and batch_words + overflow_words + n <= target + tol:
            overflow.append(item)
            overflow_words += n
        else:
            if batch:
                yield batch
            target = float(next(sizes))
            tol = target * tolerance
            batch, batch_words = overflow, overflow_words
            overflow, overflow_words = [], 0
            if batch_words + n <= target:
                batch.append(item)
                batch_words += n
            elif n <= tol and batch_words + n <= target + tol:
                overflow.append(item)
                overflow_words += n
            else:
                if batch:
                    yield batch
                target = float(next(sizes))
                tol = target * tolerance
                batch, batch_words = [item], n
    batch.extend(overflow)
    if batch:
        yield batch
