"""Compute backends.

* ``TorchOps``  - plain PyTorch, any device, fp32 by default.  This is the CPU
  path and the numerics reference every CUDA kernel is tested against.
* ``B200Ops``   - hand-written sm_100a kernels (``ops/csrc``), CUDA only.
  Constructing it on a machine with a GPU but without the built extension
  raises: there is no silent fallback.

thinc's equivalents are ``NumpyOps``/``CupyOps`` selected by ``require_gpu``
(reference: ``/root/reference/spacy_ray/worker.py:254-262``).
"""
from __future__ import annotations

import contextlib
import threading
from typing import Optional

import torch

from .torch_ops import TorchOps

_local = threading.local()
_default_ops: Optional[TorchOps] = None


def get_current_ops():
    ops = getattr(_local, "ops", None)
    if ops is not None:
        return ops
    global _default_ops
    if _default_ops is None:
        _default_ops = TorchOps(device="cpu")
    return _default_ops


def set_current_ops(ops) -> None:
    """Set the backend for this thread *and* as process default (thinc keeps the
    ops per-thread, which is why the reference has to call ``require_gpu`` again
    inside its training thread, ``worker.py:304-307``; we also set the default
    so a fresh thread inherits it)."""
    global _default_ops
    _local.ops = ops
    _default_ops = ops


@contextlib.contextmanager
def use_ops(ops):
    prev = getattr(_local, "ops", None)
    _local.ops = ops
    try:
        yield ops
    finally:
        _local.ops = prev


def require_cpu() -> TorchOps:
    ops = TorchOps(device="cpu")
    set_current_ops(ops)
    return ops


def require_gpu(gpu_id: int = 0, *, fused: bool = True):
    """Select CUDA device ``gpu_id`` and the sm_100a backend."""
    if not torch.cuda.is_available():
        raise RuntimeError("require_gpu: CUDA is not available in this process")
    torch.cuda.set_device(gpu_id)
    if fused:
        from .b200_ops import B200Ops

        ops = B200Ops(device=f"cuda:{gpu_id}")
    else:
        ops = TorchOps(device=f"cuda:{gpu_id}")
    set_current_ops(ops)
    return ops


def get_ops(name: str, device: Optional[str] = None):
    if name in ("torch", "cpu", "reference"):
        return TorchOps(device=device or "cpu")
    if name in ("b200", "sm100", "cuda"):
        from .b200_ops import B200Ops

        return B200Ops(device=device or "cuda:0")
    raise ValueError(f"Unknown ops backend {name!r}")


__all__ = [
    "TorchOps",
    "get_current_ops",
    "set_current_ops",
    "use_ops",
    "require_cpu",
    "require_gpu",
    "get_ops",
]
