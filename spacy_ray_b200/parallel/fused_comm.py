"""``FusedSymmComm``: the product data plane.

One sm_100a kernel per step (``ops/csrc/comm_kernels.cu``) does what the reference
does with hundreds of Ray RPCs (gradient push ``proxies.py:104`` + parameter push
``proxies.py:75``): reduce-scatter of the flat fp32 gradient bucket over NVLink peer
memory, the sharded Adam step with per-tensor clipping on the fp32 master shard,
and the all-gather push of the refreshed bf16 weights into every rank's weight
buffer - no NCCL call, no host round trip.

Buffers are symmetric allocations (``torch.distributed._symmetric_memory`` does the
handle exchange; the kernels only see raw peer pointers / the multicast address).
With ``world_size == 1`` the same kernel runs on plain local buffers, so single-GPU
training also gets the one-launch multi-tensor Adam.
"""
from __future__ import annotations

import os
from typing import Any, Dict, List, Optional

import torch

from .sync_proxy import ALIGN, FlatLayout

_CHUNK = 4096
_MAX_WORLD = 16


def _round_up(x: int, a: int) -> int:
    return (x + a - 1) // a * a


def _symm_alloc(numel: int, dtype: torch.dtype, device: torch.device, group) -> tuple:
    """-> (local tensor, [peer pointers], multicast pointer or 0)."""
    import torch.distributed._symmetric_memory as symm_mem

    t = symm_mem.empty(numel, dtype=dtype, device=device)
    hdl = symm_mem.rendezvous(t, group)
    ptrs = [int(p) for p in hdl.buffer_ptrs]
    mc = 0
    try:
        mc = int(hdl.multicast_ptr or 0)
    except Exception:
        mc = 0
    return t, ptrs, mc, hdl


def shard_tables(layout: FlatLayout, rank: int, device: torch.device) -> Dict[str, torch.Tensor]:
    """Work-item tables for the owned shard: each key is cut into 4096-element
    chunks; ``key_len`` is padded to the 128-element alignment (padding is zero in
    every buffer, so processing it is harmless and keeps all accesses 16B-vector)."""
    keys = layout.owned_keys(rank)
    s0 = layout.shard_start[rank]
    key_off, key_len, blk_key, blk_off = [], [], [], []
    for ki, k in enumerate(keys):
        n = _round_up(layout.numel[k], ALIGN)
        key_off.append(layout.offset[k] - s0)
        key_len.append(n)
        for c in range((n + _CHUNK - 1) // _CHUNK):
            blk_key.append(ki)
            blk_off.append(c)
    mk = lambda v, dt: torch.tensor(v, dtype=dt, device=device) if v else torch.zeros(0, dtype=dt, device=device)
    return {
        "key_off": mk(key_off, torch.int64), "key_len": mk(key_len, torch.int64),
        "blk_key": mk(blk_key, torch.int32), "blk_off": mk(blk_off, torch.int32),
        "keys": keys,
    }


class FusedSymmComm:
    name = "fused"

    def __init__(self, rank: int, world_size: int, layout: FlatLayout, device, optimizer: Any = None,
                 group: Any = None, grid: Optional[int] = None, timeout_s: float = 20.0):
        from ..ops.b200_ops import load_extension

        load_extension()
        if world_size > _MAX_WORLD:
            raise ValueError(f"FusedSymmComm supports at most {_MAX_WORLD} ranks")
        self.rank, self.world_size, self.layout = rank, world_size, layout
        self.device = torch.device(device)
        self.optimizer = optimizer
        self.timeout_s = timeout_s
        self.launches = 0
        total = layout.total
        # NVLS (multimem.ld_reduce / multimem.st through the NVSwitch) whenever a multicast mapping
        # exists: 23.7 us vs 30.6 us for the flagship shard size at 8 GPUs, 96.5 % vs 94.2 % scaling
        # (profiles/r1_run8_*).  SRB_NVLS=0 forces the plain peer-pointer variant.
        self.use_nvls = os.environ.get("SRB_NVLS", "1") != "0"
        if world_size > 1:
            import torch.distributed as dist

            grp = group if group is not None else dist.group.WORLD
            self.grad, self.grad_ptrs, self.grad_mc, self._h1 = _symm_alloc(total, torch.float32, self.device, grp)
            self.param, self.param_ptrs, self.param_mc, self._h2 = _symm_alloc(total, torch.bfloat16, self.device, grp)
            self.flags, self.flag_ptrs, _, self._h3 = _symm_alloc(1024, torch.int32, self.device, grp)
            self.grad.zero_()
            self.param.zero_()
            self.flags.zero_()
            torch.cuda.synchronize(self.device)
            dist.barrier(group=grp)
            self._group = grp
        else:
            self.grad = torch.zeros(total, dtype=torch.float32, device=self.device)
            self.param = torch.zeros(total, dtype=torch.bfloat16, device=self.device)
            self.flags = torch.zeros(1024, dtype=torch.int32, device=self.device)
            self.grad_ptrs, self.param_ptrs, self.flag_ptrs = [self.grad.data_ptr()], [self.param.data_ptr()], [self.flags.data_ptr()]
            self.grad_mc = self.param_mc = 0
            self._group = None
        if not self.use_nvls:
            self.grad_mc = self.param_mc = 0
        self.buffers = {"grad": self.grad, "param": self.param}
        self.tables = shard_tables(layout, rank, self.device)
        cap = layout.shard_cap
        self.m1 = torch.zeros(cap, dtype=torch.float32, device=self.device)
        self.m2 = torch.zeros(cap, dtype=torch.float32, device=self.device)
        n_keys = max(1, len(self.tables["keys"]))
        self.norms = torch.zeros(n_keys, dtype=torch.float32, device=self.device)
        self.step_t = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.epoch = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.bar = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.error = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.hyper = torch.zeros(8, dtype=torch.float32, device=self.device)
        self._hyper_host: Optional[List[float]] = None
        n_blocks = int(self.tables["blk_key"].numel())
        sms = torch.cuda.get_device_properties(self.device).multi_processor_count
        # persistent, co-resident grid (the kernel has device-wide barriers): 2 CTAs per SM
        self.grid = int(grid or max(1, min(2 * sms, n_blocks if n_blocks else 1)))
        self.master: Optional[torch.Tensor] = None
        self._steps_since_check = 0

    # ------------------------------------------------------------------ wiring
    def bind(self, proxy) -> None:
        """Adopt the proxy's fp32 master shard and expose Adam moments per key on
        the optimizer object (so checkpoints see them)."""
        self.master = proxy.master
        layout, s0 = self.layout, self.layout.shard_start[self.rank]
        opt = self.optimizer
        if opt is not None and hasattr(opt, "mom1"):
            for k in self.tables["keys"]:
                o, n = layout.offset[k] - s0, layout.numel[k]
                opt.mom1[k] = self.m1[o:o + n].view(layout.shape[k])
                opt.mom2[k] = self.m2[o:o + n].view(layout.shape[k])
                opt.nr_update.setdefault(k, 0)

    def _sync_hyper(self) -> None:
        opt = self.optimizer
        vals = [float(opt.learn_rate), float(opt.b1), float(opt.b2), float(opt.eps), float(opt.grad_clip or 0.0),
                float(opt.L2), 1.0 if opt.L2_is_weight_decay else 0.0, 1.0]
        if vals != self._hyper_host:
            self.hyper.copy_(torch.tensor(vals, dtype=torch.float32), non_blocking=False)
            self._hyper_host = vals

    # ------------------------------------------------------------------ the step
    def fused_step(self, proxy) -> None:
        if self.master is None:
            self.bind(proxy)
        self._sync_hyper()
        L, t = self.layout, self.tables
        torch.ops.srb.fused_comm_step(
            self.grad_ptrs, self.param_ptrs, self.flag_ptrs, int(self.grad_mc), int(self.param_mc),
            self.master, self.m1, self.m2, self.norms, t["blk_key"], t["blk_off"], t["key_off"], t["key_len"],
            self.hyper, self.step_t, self.epoch, self.bar, self.error,
            int(L.shard_start[self.rank]), int(L.shard_cap), int(L.total), self.rank, self.grid,
            float(self.timeout_s), True,
        )
        self.launches += 1
        if not torch.cuda.is_current_stream_capturing():
            self.host_bookkeeping()

    def host_bookkeeping(self) -> None:
        """Host-side mirror of one executed step (update counters used by checkpoints / bias
        correction on resume, periodic error-flag check).  ``fused_step`` calls it when it runs
        eagerly; ``engine.Trainer`` calls it after every CUDA-graph replay of a captured step."""
        opt = self.optimizer
        if opt is not None and hasattr(opt, "nr_update") and self.tables is not None:
            for k in self.tables["keys"]:
                opt.nr_update[k] = opt.nr_update.get(k, 0) + 1
        self._steps_since_check += 1
        if self._steps_since_check >= 64:
            self.check()

    def check(self) -> None:
        """Raise if any spin-wait in the kernel timed out (a peer died or hung)."""
        self._steps_since_check = 0
        code = int(self.error.item())
        if code != 0:
            raise RuntimeError(
                f"fused comm kernel on rank {self.rank} timed out in phase {code} "
                f"(a peer did not arrive within {self.timeout_s}s)"
            )

    # ------------------------------------------------------------------ library-style entry points
    def all_gather(self, param_flat: torch.Tensor, layout: FlatLayout) -> None:
        """Initial weight sync (once): plain collective through torch.distributed."""
        if self.world_size == 1:
            return
        import torch.distributed as dist

        cap = layout.shard_cap
        mine = param_flat[self.rank * cap:(self.rank + 1) * cap].clone()
        out = torch.empty_like(param_flat)
        dist.all_gather_into_tensor(out, mine, group=self._group)
        param_flat.copy_(out)
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self._group)

    def reduce_scatter(self, grad_flat: torch.Tensor, layout: FlatLayout) -> torch.Tensor:
        raise RuntimeError("FusedSymmComm runs reduce-scatter inside fused_step()")

    def barrier(self) -> None:
        if self.world_size > 1:
            import torch.distributed as dist

            dist.barrier(group=self._group)
