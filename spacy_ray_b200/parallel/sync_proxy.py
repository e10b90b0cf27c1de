"""Flat parameter/gradient buckets and the synchronous sharded proxy.

This is the collective re-formulation of the reference's proxy protocol
(the per-key gradient pushes + parameter pushes of
``/root/reference/spacy_ray/proxies.py:75,104`` move exactly the bytes of a
reduce-scatter + all-gather; 5.8 [DESIGN] "Flat bucket layout"):

* ``FlatLayout`` orders every parameter key of every component owner-major, so
  each rank's owned keys are one contiguous, 128-element-aligned extent (its
  reduce-scatter shard).  Ownership comes from ``divide_params`` per component,
  exactly like ``Worker.get_owned_keys`` (``worker.py:224-230``).
* ``ShardedSyncProxy`` serves ``get_param`` as views of the flat weight
  buffer, accumulates ``inc_grad`` into the flat fp32 gradient buffer and, once
  per step, runs ``comm.step(...)``: reduce-scatter -> owner-side Adam with
  per-tensor clipping on the fp32 master shard -> all-gather of the refreshed
  weights.  ``comm`` is pluggable: ``LocalComm`` (1 rank), ``TorchDistComm``
  (NCCL / gloo collectives: the "B-nccl" baseline) and ``FusedSymmComm``
  (``parallel/fused_comm.py``: one sm_100a kernel over NVLink peer memory).

``version`` (the reference's per-key counter) becomes one global epoch counter:
all keys move in lock step, so stale gradients cannot exist in this mode.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Callable, Dict, List, Optional, Sequence, Tuple

import torch

from ..ops import get_current_ops
from .util import DIVIDERS, KeyT, make_key

ALIGN = 128  # elements; keeps every key 256 B (bf16) / 512 B (fp32) aligned for TMA + vector loads


def _round_up(x: int, a: int) -> int:
    return (x + a - 1) // a * a


@dataclass
class FlatLayout:
    keys: List[KeyT]
    owner: Dict[KeyT, int]
    offset: Dict[KeyT, int]
    numel: Dict[KeyT, int]
    shape: Dict[KeyT, Tuple[int, ...]]
    shard_start: List[int]
    shard_len: List[int]           # actual extent owned by each rank (multiple of ALIGN)
    shard_cap: int                 # equal padded shard size used by library collectives
    world_size: int

    @property
    def total(self) -> int:
        return self.shard_cap * self.world_size

    def owned_keys(self, rank: int) -> List[KeyT]:
        return [k for k in self.keys if self.owner[k] == rank]

    def key_table(self, rank: Optional[int] = None) -> torch.Tensor:
        """(n_keys, 3) int64 rows ``[offset, numel, owner]`` (optionally only
        ``rank``'s keys) - the device-side table the fused Adam kernels walk."""
        rows = [[self.offset[k], self.numel[k], self.owner[k]] for k in self.keys
                if rank is None or self.owner[k] == rank]
        return torch.tensor(rows, dtype=torch.int64).reshape(-1, 3)

    @classmethod
    def build(cls, components: Sequence[Tuple[str, Any]], world_size: int, *, balance: str = "nodes") -> "FlatLayout":
        """``components``: ``[(name, model)]``.  ``balance``: ``"nodes"`` = the
        reference partition, ``"bytes"`` = byte-balanced contiguous partition, ``"lpt"`` = greedy
        largest-first assignment (what the fused exchange uses by default)."""
        if balance not in DIVIDERS:
            raise ValueError(f"Unknown shard balance {balance!r} (expected one of {sorted(DIVIDERS)})")
        divide = DIVIDERS[balance]
        per_rank: List[List[KeyT]] = [[] for _ in range(world_size)]
        shapes: Dict[KeyT, Tuple[int, ...]] = {}
        seen = set()
        for _name, model in components:
            for node in model.walk():
                for pname in node.param_names:
                    if node.has_param(pname):
                        shapes[make_key(node.id, pname)] = tuple(node.get_param(pname).shape)
            for rank, keys in enumerate(divide(model, world_size)):
                for k in keys:
                    if k in shapes and k not in seen:     # a shared node is owned once
                        seen.add(k)
                        per_rank[rank].append(k)
        keys, owner, offset, numel = [], {}, {}, {}
        shard_len = []
        for rank in range(world_size):
            pos = 0
            for k in per_rank[rank]:
                n = 1
                for d in shapes[k]:
                    n *= int(d)
                keys.append(k)
                owner[k] = rank
                numel[k] = n
                offset[k] = pos            # rank-relative for now
                pos += _round_up(n, ALIGN)
            shard_len.append(pos)
        cap = max(ALIGN, _round_up(max(shard_len) if shard_len else 0, ALIGN))
        shard_start = [r * cap for r in range(world_size)]
        for k in keys:
            offset[k] += shard_start[owner[k]]
        return cls(keys, owner, offset, numel, {k: shapes[k] for k in keys}, shard_start, shard_len, cap, world_size)


# ---- comm backends ----------------------------------------------------------------
class LocalComm:
    """world_size == 1: the 'reduce-scatter' is the identity."""
    name = "local"

    def __init__(self, rank: int = 0, world_size: int = 1):
        self.rank, self.world_size = rank, world_size

    def reduce_scatter(self, grad_flat: torch.Tensor, layout: FlatLayout) -> torch.Tensor:
        return grad_flat[layout.shard_start[0]: layout.shard_start[0] + layout.shard_cap]

    def all_gather(self, param_flat: torch.Tensor, layout: FlatLayout) -> None:
        return None

    def barrier(self) -> None:
        return None


class TorchDistComm:
    """Library collectives over ``torch.distributed`` (NCCL on GPU, gloo on CPU).
    This is the baseline path ("B-nccl" in BASELINE.md), not the product."""
    name = "dist"

    def __init__(self, rank: int, world_size: int, group=None, grad_transport: str = "fp32"):
        import torch.distributed as dist

        if grad_transport not in ("fp32", "bf16"):
            raise ValueError(f"grad_transport must be 'fp32' or 'bf16', got {grad_transport!r}")
        self.dist = dist
        self.rank, self.world_size, self.group = rank, world_size, group
        self.grad_transport = grad_transport
        self._rs_ok: Optional[bool] = None
        self._shard_buf: Optional[torch.Tensor] = None
        self._bf16_buf: Optional[torch.Tensor] = None

    def reduce_scatter(self, grad_flat: torch.Tensor, layout: FlatLayout) -> torch.Tensor:
        if self.grad_transport == "bf16" and grad_flat.dtype == torch.float32:
            # bf16 on the wire (half the bytes; SURVEY 5.8 / BASELINE "cast/scale"): every rank rounds its
            # local fp32 gradient once, the sum is formed by the collective in bf16 and handed back to the
            # fp32 optimizer.  bf16 keeps fp32's exponent range, so no loss scaling is needed.
            if self._bf16_buf is None or self._bf16_buf.shape != grad_flat.shape or self._bf16_buf.device != grad_flat.device:
                self._bf16_buf = torch.empty_like(grad_flat, dtype=torch.bfloat16)
            self._bf16_buf.copy_(grad_flat)
            return self._reduce_scatter(self._bf16_buf, layout).to(torch.float32)
        return self._reduce_scatter(grad_flat, layout)

    def _reduce_scatter(self, grad_flat: torch.Tensor, layout: FlatLayout) -> torch.Tensor:
        cap = layout.shard_cap
        mine = grad_flat[self.rank * cap:(self.rank + 1) * cap]
        if self._rs_ok is not False:
            try:
                if (self._shard_buf is None or self._shard_buf.shape != mine.shape or self._shard_buf.device != mine.device
                        or self._shard_buf.dtype != mine.dtype):
                    self._shard_buf = torch.empty_like(mine)
                self.dist.reduce_scatter_tensor(self._shard_buf, grad_flat, op=self.dist.ReduceOp.SUM, group=self.group)
                self._rs_ok = True
                return self._shard_buf
            except (RuntimeError, NotImplementedError):
                if self._rs_ok:          # it worked before: a real failure
                    raise
                self._rs_ok = False      # backend (gloo) has no reduce_scatter
        self.dist.all_reduce(grad_flat, op=self.dist.ReduceOp.SUM, group=self.group)
        return mine

    def all_gather(self, param_flat: torch.Tensor, layout: FlatLayout) -> None:
        cap = layout.shard_cap
        mine = param_flat[self.rank * cap:(self.rank + 1) * cap]
        try:
            self.dist.all_gather_into_tensor(param_flat, mine.clone(), group=self.group)
        except (RuntimeError, NotImplementedError):
            chunks = [param_flat[r * cap:(r + 1) * cap] for r in range(self.world_size)]
            tmp = [torch.empty_like(c) for c in chunks]
            self.dist.all_gather(tmp, mine.clone(), group=self.group)
            for c, t in zip(chunks, tmp):
                c.copy_(t)

    def barrier(self) -> None:
        self.dist.barrier(group=self.group)


# ---- the proxy -----------------------------------------------------------------------
class ShardedSyncProxy:
    """Synchronous data parallelism with optimizer-state sharding by parameter
    ownership (ZeRO-1-like; weights stay replicated for compute)."""

    def __init__(
        self,
        layout: FlatLayout,
        optimizer: Callable,
        *,
        rank: int,
        world_size: int,
        device: torch.device,
        comm: Any = None,
        param_dtype: torch.dtype = torch.float32,
        grad_dtype: torch.dtype = torch.float32,
        buffers: Optional[Dict[str, torch.Tensor]] = None,
    ):
        self.layout = layout
        self.optimizer = optimizer
        self.rank, self.world_size = rank, world_size
        self.device = torch.device(device)
        self.comm = comm if comm is not None else LocalComm(rank, world_size)
        self.param_dtype, self.grad_dtype = param_dtype, grad_dtype
        buffers = buffers or {}
        total = layout.total
        self.param_flat = buffers.get("param") if "param" in buffers else torch.zeros(total, dtype=param_dtype, device=self.device)
        self.grad_flat = buffers.get("grad") if "grad" in buffers else torch.zeros(total, dtype=grad_dtype, device=self.device)
        self._owned_keys = set(layout.owned_keys(rank))
        self._owned_list = layout.owned_keys(rank)
        self._views_p: Dict[KeyT, torch.Tensor] = {}
        self._views_g: Dict[KeyT, torch.Tensor] = {}
        for k in layout.keys:
            o, n = layout.offset[k], layout.numel[k]
            self._views_p[k] = self.param_flat[o:o + n].view(layout.shape[k])
            self._views_g[k] = self.grad_flat[o:o + n].view(layout.shape[k])
        # fp32 master copy of the owned shard (the weights the optimizer steps)
        s0 = layout.shard_start[rank]
        self.master = torch.zeros(layout.shard_cap, dtype=torch.float32, device=self.device)
        self._master_views: Dict[KeyT, torch.Tensor] = {
            k: self.master[layout.offset[k] - s0: layout.offset[k] - s0 + layout.numel[k]].view(layout.shape[k])
            for k in self._owned_list
        }
        self._initialised: set = set()
        self.version = 0
        self.grads_per_update = world_size
        self.n_grads_used = 0
        self.n_grads_discarded = 0
        self.other_workers: list = []
        self._grad_counts: Dict[KeyT, int] = {}
        self._grad_order: List[KeyT] = []           # gradient-completion order of the first step(s)
        self._order_taken = not hasattr(self.comm, "set_order")
        self._hook = getattr(self.comm, "key_ready", None)
        self._ensure_zeroed = getattr(self.comm, "ensure_zeroed", None)
        self._in_step = False

    # ---- ParamServer-facing ------------------------------------------------
    def set_param(self, id: int, name: str, value: torch.Tensor) -> None:
        key = make_key(id, name)
        if key not in self._views_p:
            raise KeyError(f"ShardedSyncProxy: key {key} is not part of the flat layout")
        view = self._views_p[key]
        if value.data_ptr() != view.data_ptr():
            view.copy_(value.to(device=self.device).reshape(view.shape))
        if key in self._owned_keys:
            self._master_views[key].copy_(value.to(device=self.device, dtype=torch.float32).reshape(view.shape))
        self._initialised.add(key)

    def get_param(self, id: int, name: str) -> torch.Tensor:
        return self._views_p[make_key(id, name)]

    def _before_grad_write(self) -> None:
        """A gradient is about to be written: a lazily begun step (generic path: nobody called
        ``begin_step``) starts here, and the writer is ordered after this step's accumulator clear."""
        if not self._in_step:
            self.begin_step(overlap=False)
        if self._ensure_zeroed is not None:
            self._ensure_zeroed()

    def inc_grad(self, id: int, name: str, value: torch.Tensor) -> None:
        key = make_key(id, name)
        self._before_grad_write()
        view = self._views_g[key]
        if value.data_ptr() != view.data_ptr():          # kernels may have written in place
            view.add_(value.reshape(view.shape))
        self._grad_counts[key] = self._grad_counts.get(key, 0) + 1
        if not self._order_taken:
            self._grad_order.append(key)
        if self._hook is not None:
            self._hook(key, self)                        # a completed bucket starts its exchange right away

    def set_grad(self, id: int, name: str, value: torch.Tensor) -> None:
        key = make_key(id, name)
        self._before_grad_write()
        self._views_g[key].copy_(value.reshape(self._views_g[key].shape))
        self._grad_counts[key] = 1

    def grad_buffer(self, id: int, name: str) -> torch.Tensor:
        """Destination view for kernels that accumulate a gradient in place."""
        self._before_grad_write()
        return self._views_g[make_key(id, name)]

    # ---- worker-facing (protocol compatibility) -------------------------------
    def check_version(self, key: KeyT, version: int) -> Optional[bool]:
        if key not in self._views_p:
            return None
        return version == self.version

    def send_param(self, key: KeyT) -> None:
        """Per-key pushes don't exist in this mode; the all-gather in ``step``
        publishes every owned key at once."""
        return None

    def receive_param(self, key: KeyT, version: int, value: torch.Tensor) -> None:
        self._views_p[key].copy_(value.reshape(self._views_p[key].shape))

    @property
    def percent_grads_used(self) -> Optional[float]:
        return 1.0 if self.n_grads_used else None

    # ---- the step --------------------------------------------------------------
    def begin_step(self, overlap: bool = True) -> None:
        """Optional: announce the start of a step's backward pass.  With ``overlap`` (and a comm
        backend that supports it) each gradient bucket is exchanged as soon as its last ``inc_grad``
        arrives, concurrently with the rest of the backward pass; ``step()`` then only launches
        what is left and joins.  Without the call, ``step()`` exchanges everything at the end."""
        begin = getattr(self.comm, "begin_step", None)
        if begin is not None:
            begin(self, overlap)
        self._in_step = True

    def take_grad_order(self) -> List[KeyT]:
        """The order in which this step's gradients were completed (first occurrence per key)."""
        self._order_taken = True
        seen, out = set(), []
        for k in self._grad_order:
            if k not in seen:
                seen.add(k)
                out.append(k)
        self._grad_order = []
        return out

    def quiesce(self) -> None:
        """Make the current stream wait until every peer's weights of the last exchange have landed
        (consumers normally do this themselves, kernel by kernel: ``ops/csrc/gate.cuh``)."""
        q = getattr(self.comm, "quiesce", None)
        if q is not None:
            q()

    def step(self) -> None:
        """Gradient exchange + sharded optimizer + weight publication."""
        if not self._in_step:
            self.begin_step(overlap=False)
        self._in_step = False
        join = getattr(get_current_ops(), "join_side", None)
        if join is not None:
            join()            # gradient GEMMs the backend ran on a side stream write into grad_flat
        fused = getattr(self.comm, "fused_step", None)
        if fused is not None:
            fused(self)
        else:
            layout = self.layout
            gshard = self.comm.reduce_scatter(self.grad_flat, layout)      # (cap,) summed over ranks
            s0 = layout.shard_start[self.rank]
            pshard = self.param_flat[s0:s0 + layout.shard_cap]
            multi = getattr(self.optimizer, "step_shard", None)
            if multi is not None:
                multi(self, gshard)
            else:
                for k in self._owned_list:
                    o, n = layout.offset[k] - s0, layout.numel[k]
                    g = gshard[o:o + n].view(layout.shape[k])
                    self.optimizer(k, self._master_views[k], g)
                if self.param_dtype == torch.float32:
                    pshard.copy_(self.master)
                else:
                    pshard.copy_(self.master.to(self.param_dtype))
            self.comm.all_gather(self.param_flat, layout)
            self.grad_flat.zero_()
        self.n_grads_used += self.world_size * len(self._owned_list)
        self._grad_counts.clear()
        self.version += 1

    def sync_from_owner(self) -> None:
        """Broadcast every owner's current weights (used once at start so all
        ranks begin from bit-identical parameters, and after a resume)."""
        self.comm.all_gather(self.param_flat, self.layout)

    def owned_keys(self) -> List[KeyT]:
        return list(self._owned_list)
