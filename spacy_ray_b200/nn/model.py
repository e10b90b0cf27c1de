"""Model tree with stable node ids, named parameters and a proxy hook.

This is the substrate the reference gets from thinc and manipulates through
private attributes (``node._params.proxy = proxy`` at
``/root/reference/spacy_ray/util.py:46,50``; ``model.walk()``, ``node.id``,
``node.param_names`` at ``util.py:58-63``).  The contract kept here:

* every node gets a unique integer ``id`` from a process-global counter at
  construction, so two processes that build the same pipeline in the same
  order agree on ids - the whole ``(node.id, param_name)`` key scheme depends
  on this;
* ``walk()`` is breadth first and visits each node once;
* when a ``proxy`` is installed on the node's ``ParamServer``, ``get_param``
  reads through it, ``inc_grad``/``set_grad`` are delegated and *not* stored
  locally (so ``has_grad`` stays False and ``finish_update`` is a no-op -
  that is what makes the reference's ``FakeOptimizer`` trick work,
  ``worker.py:265-278``).

Tensors are ``torch.Tensor``; layers do explicit forward + backprop callbacks
(no autograd) so the hot path can be handed to fused sm_100a kernels.
"""
from __future__ import annotations

import itertools
import threading
from typing import Any, Callable, Dict, Iterable, Iterator, List, Optional, Sequence, Tuple

import torch

KeyT = Tuple[int, str]

_id_lock = threading.Lock()
_id_counter = itertools.count(1)


def _next_id() -> int:
    with _id_lock:
        return next(_id_counter)


def reset_model_ids(start: int = 1) -> None:
    """Restart the global id counter.  ``init_nlp`` calls this so that node ids
    depend only on the config (same ids on every rank and after a restart),
    not on what else the process constructed earlier."""
    global _id_counter
    with _id_lock:
        _id_counter = itertools.count(start)


class ParamServer:
    """Per-node parameter/gradient store with an optional proxy.

    Proxy interface (duck typed; see ``parallel/proxies.py``):
    ``get_param(id, name)``, ``set_param(id, name, value)``,
    ``inc_grad(id, name, value)``, ``set_grad(id, name, value)``."""

    __slots__ = ("_params", "_grads", "proxy")

    def __init__(self):
        self._params: Dict[KeyT, torch.Tensor] = {}
        self._grads: Dict[KeyT, torch.Tensor] = {}
        self.proxy: Any = None

    @property
    def param_keys(self) -> Tuple[KeyT, ...]:
        return tuple(self._params.keys())

    @property
    def grad_keys(self) -> Tuple[KeyT, ...]:
        return tuple(self._grads.keys())

    def has_param(self, model_id: int, name: str) -> bool:
        return (model_id, name) in self._params

    def has_grad(self, model_id: int, name: str) -> bool:
        return (model_id, name) in self._grads

    def get_param(self, model_id: int, name: str) -> torch.Tensor:
        key = (model_id, name)
        if self.proxy is not None:
            value = self.proxy.get_param(model_id, name)
            self._params[key] = value
            return value
        return self._params[key]

    def set_param(self, model_id: int, name: str, value: torch.Tensor) -> None:
        if self.proxy is not None:
            self.proxy.set_param(model_id, name, value)
        self._params[(model_id, name)] = value

    def get_grad(self, model_id: int, name: str) -> torch.Tensor:
        return self._grads[(model_id, name)]

    def set_grad(self, model_id: int, name: str, value: torch.Tensor) -> None:
        if self.proxy is not None:
            self.proxy.set_grad(model_id, name, value)
        else:
            self._grads[(model_id, name)] = value

    def inc_grad(self, model_id: int, name: str, value: torch.Tensor) -> None:
        if self.proxy is not None:
            self.proxy.inc_grad(model_id, name, value)
            return
        key = (model_id, name)
        if key not in self._grads:
            self._grads[key] = value.clone()
        else:
            self._grads[key] += value

    def clear_grads(self) -> None:
        self._grads.clear()


class Model:
    """A node in the model tree.

    ``forward(model, X, is_train) -> (Y, backprop)``;
    ``init(model, X, Y)`` infers missing dims and allocates parameters."""

    def __init__(
        self,
        name: str,
        forward: Callable[["Model", Any, bool], Tuple[Any, Callable]],
        *,
        init: Optional[Callable[["Model", Any, Any], None]] = None,
        dims: Optional[Dict[str, Optional[int]]] = None,
        params: Optional[Dict[str, Optional[torch.Tensor]]] = None,
        layers: Sequence["Model"] = (),
        attrs: Optional[Dict[str, Any]] = None,
        refs: Optional[Dict[str, Optional["Model"]]] = None,
        ops: Any = None,
    ):
        from ..ops import get_current_ops

        self.name = name
        self.id = _next_id()
        self._func = forward
        self.init = init
        self.ops = ops if ops is not None else get_current_ops()
        self._dims: Dict[str, Optional[int]] = dict(dims or {})
        self._layers: List[Model] = list(layers)
        self._attrs: Dict[str, Any] = dict(attrs or {})
        self._refs: Dict[str, Optional[Model]] = dict(refs or {})
        self._params = ParamServer()
        self._param_names: List[str] = []
        for pname, value in (params or {}).items():
            self._param_names.append(pname)
            if value is not None:
                self._params.set_param(self.id, pname, value)

    # ---- structure -------------------------------------------------------
    @property
    def layers(self) -> List["Model"]:
        return self._layers

    def walk(self, *, order: str = "bfs") -> Iterator["Model"]:
        if order == "bfs":
            queue = [self]
            seen = set()
            while queue:
                node = queue.pop(0)
                if id(node) in seen:
                    continue
                seen.add(id(node))
                yield node
                queue.extend(node._layers)
        elif order == "dfs_pre":
            seen = set()
            stack = [self]
            while stack:
                node = stack.pop()
                if id(node) in seen:
                    continue
                seen.add(id(node))
                yield node
                stack.extend(reversed(node._layers))
        else:
            raise ValueError(f"Unknown walk order {order!r}")

    def get_ref(self, name: str) -> "Model":
        if name not in self._refs or self._refs[name] is None:
            raise KeyError(f"Model '{self.name}' has no ref '{name}'")
        return self._refs[name]  # type: ignore[return-value]

    def has_ref(self, name: str) -> bool:
        return self._refs.get(name) is not None

    def set_ref(self, name: str, value: Optional["Model"]) -> None:
        self._refs[name] = value

    @property
    def ref_names(self) -> Tuple[str, ...]:
        return tuple(self._refs)

    @property
    def attrs(self) -> Dict[str, Any]:
        return self._attrs

    # ---- dims ------------------------------------------------------------
    @property
    def dim_names(self) -> Tuple[str, ...]:
        return tuple(self._dims)

    def has_dim(self, name: str) -> Optional[bool]:
        if name not in self._dims:
            return False
        return True if self._dims[name] is not None else None

    def get_dim(self, name: str) -> int:
        if name not in self._dims:
            raise KeyError(f"Model '{self.name}' has no dim '{name}'")
        value = self._dims[name]
        if value is None:
            raise ValueError(f"Dim '{name}' of model '{self.name}' is not set")
        return value

    def maybe_get_dim(self, name: str) -> Optional[int]:
        return self._dims.get(name)

    def set_dim(self, name: str, value: int, *, force: bool = False) -> None:
        if name not in self._dims:
            raise KeyError(f"Model '{self.name}' has no dim '{name}'")
        old = self._dims[name]
        if old is not None and old != value and not force:
            raise ValueError(f"Dim '{name}' of '{self.name}' already set to {old}, can't change to {value}")
        self._dims[name] = value

    # ---- params ----------------------------------------------------------
    @property
    def param_names(self) -> Tuple[str, ...]:
        return tuple(self._param_names)

    @property
    def grad_names(self) -> Tuple[str, ...]:
        return tuple(n for n in self._param_names if self.has_grad(n))

    def has_param(self, name: str) -> Optional[bool]:
        if name not in self._param_names:
            return False
        return True if self._params.has_param(self.id, name) else None

    def get_param(self, name: str) -> torch.Tensor:
        if name not in self._param_names:
            raise KeyError(f"Unknown param '{name}' for model '{self.name}'")
        if not self._params.has_param(self.id, name):
            raise KeyError(f"Param '{name}' of model '{self.name}' has not been allocated yet")
        return self._params.get_param(self.id, name)

    def maybe_get_param(self, name: str) -> Optional[torch.Tensor]:
        return self.get_param(name) if self.has_param(name) else None

    def set_param(self, name: str, value: Optional[torch.Tensor]) -> None:
        if name not in self._param_names:
            self._param_names.append(name)
        if value is not None:
            self._params.set_param(self.id, name, value)

    def has_grad(self, name: str) -> bool:
        return self._params.has_grad(self.id, name)

    def get_grad(self, name: str) -> torch.Tensor:
        return self._params.get_grad(self.id, name)

    def set_grad(self, name: str, value: torch.Tensor) -> None:
        self._params.set_grad(self.id, name, value)

    def inc_grad(self, name: str, value: torch.Tensor) -> None:
        self._params.inc_grad(self.id, name, value)

    def grad_buffer(self, name: str) -> Optional[torch.Tensor]:
        """The proxy's accumulation buffer for this parameter's gradient, if it exposes one
        (``ShardedSyncProxy.grad_buffer``): kernels may accumulate into it in place and then
        hand the same tensor to ``inc_grad`` (which recognises it and skips the add)."""
        proxy = self._params.proxy
        fn = getattr(proxy, "grad_buffer", None) if proxy is not None else None
        if fn is None or not self._params.has_param(self.id, name):
            return None
        try:
            return fn(self.id, name)
        except KeyError:
            return None

    # ---- execution -------------------------------------------------------
    def __call__(self, X: Any, is_train: bool) -> Tuple[Any, Callable]:
        return self._func(self, X, is_train)

    def begin_update(self, X: Any) -> Tuple[Any, Callable]:
        return self._func(self, X, True)

    def predict(self, X: Any) -> Any:
        return self._func(self, X, False)[0]

    def initialize(self, X: Any = None, Y: Any = None) -> "Model":
        if self.init is not None:
            self.init(self, X, Y)
        return self

    def finish_update(self, optimizer: Callable) -> None:
        """Apply ``optimizer(key, param, grad)`` to every locally stored gradient.
        Under a proxy nothing is stored locally, so this does nothing."""
        for node in self.walk():
            for name in node.param_names:
                if node.has_grad(name):
                    param, grad = optimizer(
                        (node.id, name), node.get_param(name), node.get_grad(name)
                    )
                    node.set_param(name, param)
            node._params.clear_grads()

    def use_params(self, params: Dict[KeyT, torch.Tensor]):
        """Context manager: temporarily swap in e.g. averaged parameters."""
        model = self

        class _Ctx:
            def __enter__(self_inner):
                self_inner.backup = {}
                for node in model.walk():
                    for name in node.param_names:
                        key = (node.id, name)
                        if key in params and node.has_param(name):
                            self_inner.backup[key] = (node, name, node.get_param(name))
                            node._params._params[key] = params[key]
                return model

            def __exit__(self_inner, *a):
                for key, (node, name, value) in self_inner.backup.items():
                    node._params._params[key] = value

        return _Ctx()

    # ---- (de)serialisation ----------------------------------------------
    def to_dict(self) -> Dict[str, Any]:
        """Structure-indexed dump: node order is ``walk()`` order, so it can be
        loaded into a freshly built model whose ids differ."""
        nodes = []
        for i, node in enumerate(self.walk()):
            nodes.append(
                {
                    "index": i,
                    "name": node.name,
                    "id": node.id,
                    "dims": dict(node._dims),
                    "params": {
                        n: node.get_param(n).detach().to("cpu")
                        for n in node.param_names
                        if node.has_param(n)
                    },
                    "attrs": {
                        k: v for k, v in node._attrs.items() if isinstance(v, (int, float, str, bool, list, tuple, type(None)))
                    },
                }
            )
        return {"nodes": nodes}

    def from_dict(self, data: Dict[str, Any]) -> "Model":
        nodes = list(self.walk())
        if len(nodes) != len(data["nodes"]):
            raise ValueError(
                f"Can't load model '{self.name}': {len(nodes)} nodes here, {len(data['nodes'])} in checkpoint"
            )
        for node, blob in zip(nodes, data["nodes"]):
            if node.name != blob["name"]:
                raise ValueError(f"Node mismatch loading model: '{node.name}' vs '{blob['name']}'")
            for dname, dval in blob["dims"].items():
                if dname in node._dims and dval is not None:
                    node._dims[dname] = dval
            for pname, value in blob["params"].items():
                cur = node.maybe_get_param(pname)
                if cur is not None:
                    value = value.to(device=cur.device, dtype=cur.dtype)
                else:
                    value = node.ops.asarray(value)
                node.set_param(pname, value)
        return self

    def to_bytes(self) -> bytes:
        import io

        buf = io.BytesIO()
        torch.save(self.to_dict(), buf)
        return buf.getvalue()

    def from_bytes(self, data: bytes) -> "Model":
        import io

        return self.from_dict(torch.load(io.BytesIO(data), map_location="cpu", weights_only=False))

    def __repr__(self) -> str:
        return f"<Model {self.name} id={self.id}>"


def iter_param_keys(model: Model) -> Iterable[KeyT]:
    for node in model.walk():
        for name in node.param_names:
            yield (node.id, name)
