"""A small actor runtime with Ray's surface (``init``, ``remote``, ``get``, ``shutdown``, handles with
``.method.remote(...)``): one node here, more nodes over TCP through ``cluster.py``.

The reference's only inter-process transport is Ray actor RPC;
``ray>=0.8,<1.0`` cannot be installed here, and on one 8xB200 box the data
plane moves to NCCL / peer-memory kernels anyway.  What remains of Ray's job is
the *control plane* (create N worker processes, call ``set_proxy`` / ``train``
/ ``is_running`` on them, share eval scores) plus - for the faithful
asynchronous mode and the "B-ray" baseline emulation - a per-key host-staged
message path.  This module provides both with stdlib ``multiprocessing``:

* one OS process per actor (``spawn``), one mailbox (``mp.Queue``) each, drawn
  from a pool created at ``init()`` so handles are plain ``(index)`` values that
  can be passed between actors;
* calls run sequentially on the actor's main thread (Ray actor semantics - that
  is why ``Worker.train`` moves the training loop to a thread,
  ``/root/reference/spacy_ray/worker.py:46-49``);
* ``handle.method.remote(*a)`` returns an ``ObjectRef``; ``get`` resolves it;
  ``handle.method.fire(*a)`` is a no-reply variant for the fire-and-forget
  gradient/parameter pushes;
* liveness: ``get`` on a dead actor raises ``ActorDiedError`` instead of
  hanging (the reference has no failure detection beyond polling).

It is injected exactly where the reference injects its mock
(``Worker(..., ray=...)``, ``RayPeerProxy(..., ray=...)``).
"""
from __future__ import annotations

import itertools
import multiprocessing as mp
import os
import pickle
import queue
import threading
import time
import traceback
from typing import Any, Dict, List, Optional

_CTX = mp.get_context("spawn")
_POOL_SIZE = 40


class ActorDiedError(RuntimeError):
    pass


class RemoteError(RuntimeError):
    pass


class ObjectRef:
    __slots__ = ("call_id", "target")

    def __init__(self, call_id: int, target: int):
        self.call_id = call_id
        self.target = target


class _Runtime:
    """Per-process runtime state (driver or actor)."""

    def __init__(self, pool: List[Any], my_index: int, n_gpus_used=None, extra_env: Optional[Dict[str, str]] = None):
        self.pool = pool
        self.my_index = my_index
        self.results: Dict[int, Any] = {}
        self.events: Dict[int, threading.Event] = {}
        self.deferred: List[tuple] = []
        self.lock = threading.Lock()
        self.call_ids = itertools.count(my_index * (1 << 40) + 1)
        self.procs: Dict[int, Any] = {}
        self.next_actor = 1            # pool slot 0 is the driver
        self.gpu_cursor = 0
        self.mailbox_thread: Optional[int] = None
        self.extra_env = dict(extra_env or {})
        self.alive = True
        self.cluster: Any = None       # cluster.Head in the driver of a multi-node run

    # -- sending ------------------------------------------------------------
    def send_call(self, target: int, method: str, args, kwargs, want_reply: bool) -> Optional[ObjectRef]:
        call_id = next(self.call_ids) if want_reply else 0
        if want_reply:
            with self.lock:
                self.events[call_id] = threading.Event()
        self.pool[target].put(("call", call_id, self.my_index, method, args, kwargs))
        return ObjectRef(call_id, target) if want_reply else None

    # -- receiving ----------------------------------------------------------
    def _handle_reply(self, call_id: int, ok: bool, value: Any) -> None:
        with self.lock:
            self.results[call_id] = (ok, value)
            ev = self.events.get(call_id)
        if ev is not None:
            ev.set()

    def pump_once(self, timeout: float) -> Optional[tuple]:
        """Read one message; replies are filed, calls are returned to the caller."""
        try:
            msg = self.pool[self.my_index].get(timeout=timeout)
        except queue.Empty:
            return None
        if msg[0] == "reply":
            self._handle_reply(msg[1], msg[2], msg[3])
            return None
        return msg

    def wait(self, ref: ObjectRef, timeout: Optional[float]) -> Any:
        deadline = None if timeout is None else time.time() + timeout
        ev = self.events.get(ref.call_id)
        on_mailbox = self.mailbox_thread is None or self.mailbox_thread == threading.get_ident()
        while True:
            with self.lock:
                if ref.call_id in self.results:
                    ok, value = self.results.pop(ref.call_id)
                    self.events.pop(ref.call_id, None)
                    if ok:
                        return value
                    raise RemoteError(value)
            if deadline is not None and time.time() > deadline:
                raise TimeoutError(f"get(): no reply from actor {ref.target} within {timeout}s")
            proc = self.procs.get(ref.target)
            if proc is not None and not proc.is_alive():
                # drain anything it managed to send before dying
                while self.pump_once(0.0) is not None:
                    pass
                with self.lock:
                    if ref.call_id in self.results:
                        continue
                raise ActorDiedError(f"actor {ref.target} died (exit code {proc.exitcode})")
            if on_mailbox:
                msg = self.pump_once(0.05)
                if msg is not None:
                    self.deferred.append(msg)     # a call that arrived while we were blocked
            else:
                if ev is not None:
                    ev.wait(0.05)
                else:
                    time.sleep(0.01)


_rt: Optional[_Runtime] = None


def _require_rt() -> _Runtime:
    if _rt is None:
        raise RuntimeError("actor runtime not initialised: call init() first")
    return _rt


# ---- public API (Ray surface) -------------------------------------------------
def is_initialized() -> bool:
    return _rt is not None


def init(address: Optional[str] = None, ignore_reinit_error: bool = True, **kwargs) -> None:
    """Start the runtime in the driver.  Without ``address`` (or ``auto`` / ``local``): single node.  With
    ``address="HOST:PORT"`` and ``nodes=N`` the driver becomes the head of an N-node cluster: it listens there
    and waits for N-1 agents (``python -m spacy_ray_b200 ray node --address HOST:PORT``), see ``cluster.py``."""
    global _rt
    if _rt is not None:
        if ignore_reinit_error:
            return
        raise RuntimeError("actor runtime already initialised")
    nodes = int(kwargs.get("nodes") or os.environ.get("SRB_NODES", "1") or 1)
    if address not in (None, "", "auto", "local") and nodes > 1:
        from . import cluster

        head = cluster.Head(address, nodes, _POOL_SIZE, _CTX, accept_timeout=float(kwargs.get("accept_timeout", 300.0)))
        _rt = _Runtime(head.pool, 0, extra_env=kwargs.get("env"))
        _rt.cluster = head
        return
    if address not in (None, "", "auto", "local") and nodes <= 1:
        # Ray semantics would be "join the cluster at ADDRESS"; with one node there is nothing to join
        pass
    pool = [_CTX.Queue() for _ in range(_POOL_SIZE)]
    _rt = _Runtime(pool, 0, extra_env=kwargs.get("env"))


def shutdown() -> None:
    global _rt
    if _rt is None:
        return
    rt = _rt
    for idx, proc in list(rt.procs.items()):
        if proc.is_alive():
            try:
                rt.pool[idx].put(("stop",))
            except Exception:
                pass
    for idx, proc in list(rt.procs.items()):
        proc.join(timeout=5)
        if proc.is_alive():
            proc.terminate()
            proc.join(timeout=2)
    if rt.cluster is not None:
        rt.cluster.shutdown()
    _rt = None


def get(refs, timeout: Optional[float] = None):
    rt = _require_rt()
    if isinstance(refs, (list, tuple)):
        return [rt.wait(r, timeout) for r in refs]
    if refs is None:
        return None
    return rt.wait(refs, timeout)


class _MethodHandle:
    __slots__ = ("target", "name")

    def __init__(self, target: int, name: str):
        self.target = target
        self.name = name

    def remote(self, *args, **kwargs) -> ObjectRef:
        return _require_rt().send_call(self.target, self.name, args, kwargs, True)

    def fire(self, *args, **kwargs) -> None:
        _require_rt().send_call(self.target, self.name, args, kwargs, False)


class ActorHandle:
    """Picklable reference to an actor; ``handle.method`` -> callable stub."""

    def __init__(self, index: int, cls_name: str = ""):
        self._index = index
        self._cls_name = cls_name

    def __getattr__(self, name: str) -> _MethodHandle:
        if name.startswith("_"):
            raise AttributeError(name)
        return _MethodHandle(self._index, name)

    def __reduce__(self):
        return (ActorHandle, (self._index, self._cls_name))

    def __eq__(self, other) -> bool:
        return isinstance(other, ActorHandle) and other._index == self._index

    def __hash__(self) -> int:
        return hash(("actor", self._index))

    def __repr__(self) -> str:
        return f"<ActorHandle {self._cls_name}#{self._index}>"


class _RemoteClass:
    def __init__(self, cls, options: Optional[Dict[str, Any]] = None):
        self._cls = cls
        self._options = dict(options or {})

    def options(self, **kwargs) -> "_RemoteClass":
        return _RemoteClass(self._cls, {**self._options, **kwargs})

    def remote(self, *args, **kwargs) -> ActorHandle:
        rt = _require_rt()
        if rt.my_index != 0:
            raise RuntimeError("actors can only be created from the driver process")
        head = rt.cluster
        want_gpu = int(self._options.get("num_gpus", 0) or 0) > 0
        if head is not None:
            node = head.place(int(want_gpu))
            index = head.new_index(node)
        else:
            node = 0
            index = rt.next_actor
            if index >= len(rt.pool):
                raise RuntimeError(f"actor pool exhausted ({len(rt.pool) - 1} actors)")
            rt.next_actor += 1
        env = dict(rt.extra_env)
        if want_gpu:
            visible = os.environ.get("CUDA_VISIBLE_DEVICES") if node == 0 else None
            devices = visible.split(",") if visible else None
            if head is not None:
                gpu = head.next_gpu(node)
            else:
                gpu = rt.gpu_cursor
                rt.gpu_cursor += 1
            if self._options.get("isolate_gpu", False):
                # Ray-style isolation: the actor sees exactly one device (index 0)
                env["CUDA_VISIBLE_DEVICES"] = devices[gpu % len(devices)] if devices else str(gpu)
            else:
                # All devices stay visible (CUDA symmetric memory / NVLink peer mappings need every
                # rank to address a *distinct* device ordinal); the actor is told which one is its own.
                env["SRB_ASSIGNED_GPU"] = str(gpu % len(devices) if devices else gpu)
        payload = pickle.dumps((self._cls, args, kwargs))
        name = f"srb-actor-{self._cls.__name__}-{index}"
        if node != 0:
            rt.procs[index] = head.spawn_remote(node, index, payload, env, name)
            return ActorHandle(index, self._cls.__name__)
        proc = _CTX.Process(target=_actor_main, args=(rt.pool, index, payload, env), daemon=True, name=name)
        proc.start()
        rt.procs[index] = proc
        return ActorHandle(index, self._cls.__name__)


def remote(cls=None, **options):
    if cls is None:
        return lambda c: _RemoteClass(c, options)
    return _RemoteClass(cls, options)


# ---- actor process ---------------------------------------------------------------
def _actor_main(pool, index: int, payload: bytes, env: Dict[str, str]) -> None:
    global _rt
    os.environ.update(env)
    _rt = _Runtime(pool, index)
    _rt.mailbox_thread = threading.get_ident()
    instance = None
    init_error = None
    try:
        cls, args, kwargs = pickle.loads(payload)
        instance = cls(*args, **kwargs)
    except BaseException:
        init_error = traceback.format_exc()
    rt = _rt
    while rt.alive:
        if rt.deferred:
            msg = rt.deferred.pop(0)
        else:
            msg = rt.pump_once(0.2)
            if msg is None:
                continue
        if msg[0] == "stop":
            break
        _kind, call_id, sender, method, args, kwargs = msg
        ok, value = True, None
        if init_error is not None:
            ok, value = False, f"actor constructor failed:\n{init_error}"
        else:
            try:
                value = getattr(instance, method)(*args, **kwargs)
            except BaseException:
                ok, value = False, traceback.format_exc()
        if call_id:
            try:
                pool[sender].put(("reply", call_id, ok, value))
            except Exception:
                pool[sender].put(("reply", call_id, False, f"unpicklable result from {method}:\n{traceback.format_exc()}"))
        elif not ok:
            print(f"[actor {index}] error in fire-and-forget call {method}:\n{value}", flush=True)
    shutdown_hook = getattr(instance, "_on_actor_shutdown", None)
    if shutdown_hook is not None:
        try:
            shutdown_hook()
        except Exception:
            pass
