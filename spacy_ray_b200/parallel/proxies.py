"""Parameter proxies: the object a model's ``ParamServer`` delegates to.

Two implementations of the same duck-typed interface
(``get_param / set_param / inc_grad / set_grad`` + worker-facing
``check_version / send_param / receive_param``):

``PeerProxy`` (aliases ``RayPeerProxy``, ``RayOptimizer``)
    The reference's asynchronous protocol, message for message
    (``/root/reference/spacy_ray/proxies.py:9-133``): every key
    has one owner; non-owners push each gradient to the owner tagged with the
    parameter version it was computed against; the owner drops stale
    gradients, steps the optimizer once ``grads_per_update`` gradients for the
    current version have arrived (lazily, at the next read of that key) and
    pushes the new parameter to every peer, who stage it and adopt it at their
    next read.  Differences from the reference, all deliberate: state is
    guarded by a lock (the reference mutates it from two threads unguarded,
    SURVEY.md 5.2); used/discarded gradient counters are real (the reference's
    are never incremented); ``quorum`` is configurable end to end.

``ShardedSyncProxy`` (``parallel/sync_proxy.py``)
    The synchronous, collective formulation used on the B200 path: same
    ownership map, but gradients live in one flat buffer and one
    reduce-scatter -> sharded Adam -> all-gather per step replaces the
    per-key messages.
"""
from __future__ import annotations

import threading
from collections import Counter
from typing import Any, Callable, Dict, Iterable, Optional, Set, Tuple

import torch

from .util import KeyT, make_key


def _remote(method: Any, *args) -> Any:
    """Invoke an actor-handle method the Ray way (``h.method.remote(...)``) or,
    for plain objects used in tests, directly."""
    fire = getattr(method, "fire", None)          # built-in runtime: no-reply send
    if fire is not None:
        return fire(*args)
    rem = getattr(method, "remote", None)
    return rem(*args) if rem is not None else method(*args)


class PeerProxy:
    def __init__(
        self,
        peers: Dict[KeyT, Any],
        optimizer: Callable,
        keys: Iterable[KeyT],
        *,
        grads_per_update: int = 2,
        ray: Any = None,
        stage_to_host: bool = False,
        all_peers: Any = None,
        self_index: Optional[int] = None,
    ):
        self.ray = ray
        self.optimizer = optimizer
        self.grads_per_update = int(grads_per_update)
        self.peers = dict(peers)
        self._owned_keys: Set[KeyT] = set(keys)
        self.other_workers: list = []
        for key, peer in self.peers.items():
            if key not in self._owned_keys and not any(peer is w or peer == w for w in self.other_workers):
                self.other_workers.append(peer)
        if all_peers is not None and self_index is not None:
            # The reference derives `other_workers` from key ownership only, so a
            # worker that owns no key (fewer param groups than workers) is never
            # sent updated parameters.  Given the full peer list we push to everyone.
            self.other_workers = [p for i, p in enumerate(all_peers) if i != self_index]
        self._params: Dict[KeyT, torch.Tensor] = {}
        self._grads: Dict[KeyT, Optional[torch.Tensor]] = {}
        self._versions: Counter = Counter()
        self._grad_counts: Counter = Counter()
        self._next_params: Dict[KeyT, Tuple[int, torch.Tensor]] = {}
        self._lock = threading.RLock()
        self.stage_to_host = stage_to_host
        # observability (the reference's counters are dead code, worker.py:105-106)
        self.n_grads_used = 0
        self.n_grads_discarded = 0
        self.n_updates = 0
        self.n_msgs_sent = 0
        self.bytes_sent = 0

    # ---- helpers ---------------------------------------------------------
    def _payload(self, value: torch.Tensor) -> torch.Tensor:
        """What goes on the wire.  With ``stage_to_host`` a device tensor is copied
        to host memory first - what pickling a GPU array through Ray's object
        store costs."""
        v = value.detach()
        if self.stage_to_host and v.device.type != "cpu":
            v = v.to("cpu")
        self.n_msgs_sent += 1
        self.bytes_sent += v.numel() * v.element_size()
        return v

    def _to_local(self, value: torch.Tensor, like: Optional[torch.Tensor]) -> torch.Tensor:
        if like is not None and (value.device != like.device or value.dtype != like.dtype):
            return value.to(device=like.device, dtype=like.dtype)
        return value

    # ---- worker-facing ---------------------------------------------------
    def check_version(self, key: KeyT, version: int) -> Optional[bool]:
        with self._lock:
            if key not in self._versions:
                return None
            return self._versions[key] == version

    def send_param(self, key: KeyT) -> None:
        with self._lock:
            param = self._payload(self._params[key])
            version = self._versions[key]
        for peer in self.other_workers:
            _remote(peer.set_param, key, version, param)

    def receive_param(self, key: KeyT, version: int, value: torch.Tensor) -> None:
        """Stage a pushed parameter; it is adopted at the next ``get_param`` so a
        gradient computed against the old value is never labelled with the new
        version."""
        with self._lock:
            self._next_params[key] = (version, value)

    def receive_grad(self, key: KeyT, version: int, value: torch.Tensor) -> bool:
        """Owner side of a gradient push: accumulate iff the sender computed it
        against our current version."""
        with self._lock:
            if self.check_version(key, version):
                self.inc_grad(key[0], key[1], value, _remote_origin=True)
                self.n_grads_used += 1
                return True
            self.n_grads_discarded += 1
            return False

    # ---- ParamServer-facing ------------------------------------------------
    def set_param(self, id: int, name: str, value: torch.Tensor) -> None:
        key = make_key(id, name)
        with self._lock:
            if key in self._owned_keys or key not in self._params:
                self._params[key] = value
                self._versions[key] += 1
                self._grads[key] = None
                self._grad_counts[key] = 0

    def load_param(self, id: int, name: str, value: torch.Tensor, version: int) -> None:
        """Resume: adopt ``value`` for ``key`` on THIS rank whether or not it owns the key, at an
        explicit ``version``.  Every rank loads the same checkpoint with the same version, so the
        replicas agree and the owners' version gates (``check_version``) accept the peers' first
        gradients - ``set_param`` alone ignores non-owned keys that already exist (reference
        ``proxies.py:62-69``), which left resumed peers on their random initialisation."""
        key = make_key(id, name)
        with self._lock:
            self._params[key] = self._to_local(value, self._params.get(key))
            self._versions[key] = int(version)
            self._grads[key] = None
            self._grad_counts[key] = 0
            self._next_params.pop(key, None)

    def get_param(self, id: int, name: str) -> torch.Tensor:
        key = make_key(id, name)
        with self._lock:
            self._maybe_update_param(key)
            return self._params[key]

    def set_grad(self, id: int, name: str, value: torch.Tensor) -> None:
        key = make_key(id, name)
        with self._lock:
            if key in self._owned_keys:
                self._grads[key] = value
                self._grad_counts[key] = 1

    def inc_grad(self, id: int, name: str, value: torch.Tensor, _remote_origin: bool = False) -> None:
        key = make_key(id, name)
        with self._lock:
            self._grad_counts[key] += 1
            if key not in self._owned_keys:
                peer = self.peers[key]
                version = self._versions[key]
                payload = self._payload(value)
            else:
                peer = None
                cur = self._grads.get(key)
                value = self._to_local(value, self._params.get(key))
                if cur is None:
                    self._grads[key] = value.to(torch.float32).clone()
                else:
                    cur += value
                if not _remote_origin:
                    self.n_grads_used += 1
        if peer is not None:
            _remote(peer.inc_grad, key, version, payload)

    def _maybe_update_param(self, key: KeyT) -> bool:
        with self._lock:
            staged = self._next_params.pop(key, None)
            if staged is not None:
                version, value = staged
                self._params[key] = self._to_local(value, self._params.get(key))
                self._versions[key] = version
                self._grad_counts[key] = 0
                self._grads[key] = None
                return True
            if key not in self._owned_keys:
                return False
            if self._grad_counts[key] < self.grads_per_update:
                return False
            grad = self._grads.get(key)
            if grad is None:
                return False
            self._versions[key] += 1
            param, _ = self.optimizer(key, self._params[key], grad)
            self._params[key] = param
            self._grads[key] = None
            self._grad_counts[key] = 0
            self.n_updates += 1
        self.send_param(key)
        return True

    # ---- extras ----------------------------------------------------------
    def flush_updates(self) -> int:
        """Run every pending owner update now (instead of at the next read)."""
        n = 0
        for key in list(self._owned_keys):
            if key in self._params and self._maybe_update_param(key):
                n += 1
        return n

    @property
    def percent_grads_used(self) -> Optional[float]:
        total = self.n_grads_used + self.n_grads_discarded
        return (self.n_grads_used / total) if total else None


# Names the reference / north star use for this class.
RayPeerProxy = PeerProxy
RayOptimizer = PeerProxy
