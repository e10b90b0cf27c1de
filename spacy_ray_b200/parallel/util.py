"""Parameter-ownership partitioning, proxy installation, timers.

API-compatible with ``/root/reference/spacy_ray/util.py`` (``KeyT``,
``make_key``, ``divide_params``, ``set_params_proxy``, ``Timer``,
``ManyTimer``).  ``divide_params`` reproduces the reference partition exactly
(SURVEY.md 2.1 #6): keys grouped per node so a layer's params share an owner,
``max(1, n_groups // num_workers)`` consecutive groups per rank, all leftovers
to the last rank, trailing ranks empty if there are fewer groups than ranks.
``divide_params_balanced`` is the opt-in byte-balanced alternative.
"""
from __future__ import annotations

import time
from typing import Dict, List, Tuple

KeyT = Tuple[int, str]


def make_key(model_id: int, name: str) -> KeyT:
    return (model_id, name)


def _key_groups(model) -> List[List[KeyT]]:
    groups: Dict[int, List[KeyT]] = {}
    for node in model.walk():
        for name in node.param_names:
            groups.setdefault(node.id, []).append(make_key(node.id, name))
    return [g for g in groups.values() if g]


def divide_params(model, num_workers: int) -> List[List[KeyT]]:
    """Owner partition of a component model's parameter keys (reference layout)."""
    groups = _key_groups(model)
    per_rank = max(1, len(groups) // num_workers)
    shares: List[List[KeyT]] = [[] for _ in range(num_workers)]
    for gi, group in enumerate(groups):
        owner = min(gi // per_rank, num_workers - 1)
        shares[owner].extend(group)
    return shares


def divide_params_balanced(model, num_workers: int) -> List[List[KeyT]]:
    """Contiguous partition of node groups minimising the largest shard (by
    element count) greedily: keeps a layer's params together like
    ``divide_params`` but evens out bytes (the reference counts nodes, so the
    embedding tables make rank 0 ~2x heavier than the rest)."""
    groups = _key_groups(model)
    sizes = []
    by_key = {}
    for node in model.walk():
        for name in node.param_names:
            if node.has_param(name):
                by_key[make_key(node.id, name)] = int(node.get_param(name).numel())
    for g in groups:
        sizes.append(sum(by_key.get(k, 0) for k in g))
    # Exact "linear partition": split the group sequence into num_workers contiguous
    # runs minimising the heaviest run (DP over a few dozen groups).
    n = len(groups)
    k = min(num_workers, max(n, 1))
    prefix = [0]
    for sz in sizes:
        prefix.append(prefix[-1] + sz)
    INF = float("inf")
    best = [[INF] * (n + 1) for _ in range(k + 1)]
    cut = [[0] * (n + 1) for _ in range(k + 1)]
    best[0][0] = 0
    for j in range(1, k + 1):
        for i in range(j, n + 1):
            for s in range(j - 1, i):
                cost = max(best[j - 1][s], prefix[i] - prefix[s])
                if cost < best[j][i]:
                    best[j][i], cut[j][i] = cost, s
    bounds = [n]
    i = n
    for j in range(k, 0, -1):
        i = cut[j][i]
        bounds.append(i)
    bounds.reverse()
    shares: List[List[KeyT]] = [[] for _ in range(num_workers)]
    for r in range(k):
        for g in groups[bounds[r]:bounds[r + 1]]:
            shares[r].extend(g)
    return shares


def divide_params_lpt(model, num_workers: int) -> List[List[KeyT]]:
    """Longest-processing-time assignment of node groups to ranks: groups sorted by size,
    each handed to the currently lightest rank (ties -> lowest rank; deterministic, so every
    rank derives the same map).  Not contiguous in walk order, but the flat layout is
    owner-major anyway; the heaviest rank ends up at max(largest tensor, ~mean)."""
    groups = _key_groups(model)
    by_key = {}
    for node in model.walk():
        for name in node.param_names:
            if node.has_param(name):
                by_key[make_key(node.id, name)] = int(node.get_param(name).numel())
    sized = [(sum(by_key.get(k, 0) for k in g), i, g) for i, g in enumerate(groups)]
    sized.sort(key=lambda t: (-t[0], t[1]))
    load = [0] * num_workers
    picked: List[List[int]] = [[] for _ in range(num_workers)]
    for size, i, _g in sized:
        r = min(range(num_workers), key=lambda j: (load[j], j))
        load[r] += size
        picked[r].append(i)
    shares: List[List[KeyT]] = [[] for _ in range(num_workers)]
    for r in range(num_workers):
        for i in sorted(picked[r]):                # keep walk order inside a rank
            shares[r].extend(groups[i])
    return shares


DIVIDERS = {"nodes": divide_params, "bytes": divide_params_balanced, "lpt": divide_params_lpt}


def set_params_proxy(model, proxy) -> None:
    """Install ``proxy`` on every node of ``model``: existing parameter values
    are handed to ``proxy.set_param`` first, then the node's ``ParamServer``
    routes all further reads/grad writes through the proxy."""
    for node in model.walk():
        node._params.proxy = None
        for name in node.param_names:
            if node.has_param(name):
                proxy.set_param(node.id, name, node.get_param(name))
        node._params.proxy = proxy


def clear_params_proxy(model) -> None:
    for node in model.walk():
        node._params.proxy = None


class Timer:
    """Accumulating wall-clock timer (context manager)."""

    def __init__(self, state: str):
        self.state = state
        self.sum = 0.0
        self.n = 0
        self.start = 0.0

    def __enter__(self) -> "Timer":
        self.start = time.time()
        self.n += 1
        return self

    def __exit__(self, *exc) -> None:
        self.sum += time.time() - self.start

    @property
    def mean(self) -> float:
        return self.sum / self.n if self.n else 0.0


class ManyTimer:
    def __init__(self):
        self.timers: Dict[str, Timer] = {}

    def __call__(self, key: str) -> Timer:
        timer = self.timers.get(key)
        if timer is None:
            timer = self.timers[key] = Timer(key)
        return timer

    def report(self) -> Dict[str, Dict[str, float]]:
        return {k: {"sum": t.sum, "n": t.n, "mean": t.mean} for k, t in self.timers.items()}
