"""``FusedSymmComm``: the product data plane.

What the reference does with hundreds of Ray RPCs per step (gradient push
``/root/reference/spacy_ray/proxies.py:102-104``, owner-side optimizer ``proxies.py:126-128``,
parameter push ``proxies.py:71-75``, lazy adoption at the next read ``proxies.py:111-118``)
becomes a handful of launches of one sm_100a kernel pair (``ops/csrc/comm_kernels.cu``,
``bucket_reduce_kernel`` + ``bucket_update_kernel``):

* the flat gradient bucket is cut into a few **buckets in the order the backward pass
  completes them** (learned from the ``inc_grad`` sequence of the first step);
* as soon as a bucket's last gradient has been produced, its exchange - reduce-scatter over
  NVLink peer memory / NVLS, sharded Adam (or RAdam / SGD, + parameter averages) with
  per-tensor clipping on the fp32 master, all-gather push of the refreshed bf16 weights into
  every rank's weight buffer - is launched on a side stream and runs UNDER the rest of the
  backward pass;
* nothing waits for the weights at the end of the step: the first consumer of a bucket in
  the next forward pass (the tcgen05 GEMM's TMA producer warp, ``hash_embed_fwd_kernel``)
  carries a gate on the owners' "published" flags (``ops/csrc/gate.cuh``).

No NCCL call, no host round trip.  Buffers are symmetric allocations
(``torch.distributed._symmetric_memory`` does the handle exchange; the kernels only see raw
peer pointers / the multicast address).  With ``world_size == 1`` the same kernel runs on
plain local buffers, so single-GPU training gets the overlapped multi-tensor optimizer too.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Sequence, Tuple

import torch

from .sync_proxy import ALIGN, FlatLayout

_CHUNK = 4096
_MAX_WORLD = 16
_MAX_BUCKETS = 32
_SIGNAL_WORDS = 2 * _MAX_BUCKETS * _MAX_WORLD
OPT_ADAM, OPT_RADAM, OPT_SGD = 0, 1, 2

KeyT = Tuple[int, str]


def _round_up(x: int, a: int) -> int:
    return (x + a - 1) // a * a


def _symm_alloc(numel: int, dtype: torch.dtype, device: torch.device, group) -> tuple:
    """-> (local tensor, [peer pointers], multicast pointer or 0, handle)."""
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


# ---------------------------------------------------------------------------------------
# bucket planning (pure python; unit-tested on CPU)
# ---------------------------------------------------------------------------------------
@dataclass
class BucketPlan:
    """``buckets[b]`` = keys (of ALL owners) exchanged by launch ``b``; launches happen in
    index order on every rank."""
    buckets: List[List[KeyT]]
    bucket_of: Dict[KeyT, int] = field(default_factory=dict)

    def __post_init__(self):
        self.bucket_of = {k: b for b, ks in enumerate(self.buckets) for k in ks}

    @property
    def n(self) -> int:
        return len(self.buckets)


def plan_buckets(order: Sequence[KeyT], layout: FlatLayout, n_target: int = 6,
                 max_buckets: int = _MAX_BUCKETS) -> BucketPlan:
    """Cut the gradient-completion sequence ``order`` into ~``n_target`` contiguous buckets.

    * a bucket closes once it holds >= total / n_target elements - but only between
      different model nodes (a layer's ``W`` and ``b`` stay together);
    * embedding tables (``E``: the last gradients of the backward pass and the first weights
      of the next forward pass) never share a bucket with other parameters, so the bucket
      on the step's critical path is as small as it can be;
    * keys of the layout that never received a gradient ride in the last bucket.
    """
    seen, seq = set(), []
    for k in order:
        if k in layout.numel and k not in seen:
            seen.add(k)
            seq.append(k)
    rest = [k for k in layout.keys if k not in seen]
    total = sum(layout.numel[k] for k in seq) or 1
    target = max(1, total // max(1, n_target))
    buckets: List[List[KeyT]] = []
    cur: List[KeyT] = []
    acc = 0
    for i, k in enumerate(seq):
        if cur:
            prev = cur[-1]
            kind_change = (k[1] == "E") != (prev[1] == "E")
            node_change = k[0] != prev[0]
            if kind_change or (node_change and acc >= target and not (k[1] == "E" and prev[1] == "E")):
                buckets.append(cur)
                cur, acc = [], 0
        cur.append(k)
        acc += layout.numel[k]
    if cur:
        buckets.append(cur)
    if not buckets:
        buckets = [[]]
    buckets[-1].extend(rest)
    while len(buckets) > max_buckets:          # merge the two smallest neighbours
        sizes = [sum(layout.numel[k] for k in b) for b in buckets]
        j = min(range(len(buckets) - 1), key=lambda i: sizes[i] + sizes[i + 1])
        buckets[j:j + 2] = [buckets[j] + buckets[j + 1]]
    return BucketPlan(buckets)


def shard_tables(layout: FlatLayout, rank: int, device: torch.device, plan: Optional[BucketPlan] = None) -> Dict[str, Any]:
    """Work-item tables for the owned shard, sorted by bucket: each key is cut into 4096-element
    chunks; ``key_len`` is padded to the 128-element alignment (padding is zero in every buffer,
    so processing it is harmless and keeps all accesses 16 B vectors).  ``ranges[b]`` =
    ``(blk_begin, blk_end, key_begin, key_end)`` of bucket ``b`` in those tables."""
    owned = layout.owned_keys(rank)
    if plan is None:
        plan = BucketPlan([list(layout.keys)])
    pos = {k: i for i, k in enumerate(owned)}
    keys = sorted(owned, key=lambda k: (plan.bucket_of.get(k, plan.n - 1), pos[k]))
    s0 = layout.shard_start[rank]
    key_off, key_len, blk_key, blk_off = [], [], [], []
    ranges = []
    ki = 0
    for b in range(plan.n):
        kb, bb = ki, len(blk_key)
        while ki < len(keys) and plan.bucket_of.get(keys[ki], plan.n - 1) == b:
            k = keys[ki]
            n = _round_up(layout.numel[k], ALIGN)
            key_off.append(layout.offset[k] - s0)
            key_len.append(n)
            for c in range((n + _CHUNK - 1) // _CHUNK):
                blk_key.append(ki)
                blk_off.append(c)
            ki += 1
        ranges.append((bb, len(blk_key), kb, ki))

    def mk(v, dt):
        return torch.tensor(v, dtype=dt, device=device) if v else torch.zeros(0, dtype=dt, device=device)

    return {
        "key_off": mk(key_off, torch.int64), "key_len": mk(key_len, torch.int64),
        "blk_key": mk(blk_key, torch.int32), "blk_off": mk(blk_off, torch.int32),
        "keys": keys, "ranges": ranges,
    }


def optimizer_mode(opt: Any) -> int:
    if opt is None or getattr(opt, "use_adam", True):
        return OPT_RADAM if getattr(opt, "use_radam", False) else OPT_ADAM
    return OPT_SGD


class FusedSymmComm:
    name = "fused"

    def __init__(self, rank: int, world_size: int, layout: FlatLayout, device, optimizer: Any = None,
                 group: Any = None, grid: Optional[int] = None, timeout_s: float = 20.0,
                 n_buckets: Optional[int] = None, ops: Any = None):
        from ..ops.b200_ops import load_extension

        load_extension()
        if world_size > _MAX_WORLD:
            raise ValueError(f"FusedSymmComm supports at most {_MAX_WORLD} ranks")
        self.rank, self.world_size, self.layout = rank, world_size, layout
        self.device = torch.device(device)
        self.optimizer = optimizer
        self.ops = ops
        self.timeout_s = timeout_s
        self.launches = 0
        self.n_buckets_target = int(n_buckets or os.environ.get("SRB_COMM_BUCKETS", 4))
        self.overlap = os.environ.get("SRB_COMM_OVERLAP", "1") != "0"
        self.terminal_wait = os.environ.get("SRB_GATE_ALWAYS", "0") == "1"    # debugging: round-1 behaviour
        self.test_delay_us = 0               # tests: make this rank a late publisher (see tests/mgpu_worker.py)
        # %globaltimer trace of the exchange kernels (8 words per bucket) + free stamp slots behind them;
        # SRB_COMM_TRACE=1 or benchmarks/exchange_trace.py switch it on (it adds a few atomics per CTA)
        self.trace: Optional[torch.Tensor] = None
        self.trace_keys: Dict[int, KeyT] = {}
        if os.environ.get("SRB_COMM_TRACE", "0") == "1":
            self.enable_trace()
        total = layout.total
        # NVLS (multimem.ld_reduce / multimem.st through the NVSwitch: the switch does the fan-in / fan-out)
        # from 4 ranks up; with 2 ranks plain peer loads/stores have the shorter round trip (measured:
        # profiles/r2_exchange_trace_2gpu.md).  SRB_NVLS=1 / 0 forces it on / off.
        env_nvls = os.environ.get("SRB_NVLS", "auto")
        self.use_nvls = (world_size >= 4) if env_nvls not in ("0", "1") else env_nvls == "1"
        if world_size > 1:
            import torch.distributed as dist

            grp = group if group is not None else dist.group.WORLD
            self.grad, self.grad_ptrs, self.grad_mc, self._h1 = _symm_alloc(total, torch.float32, self.device, grp)
            self.param, self.param_ptrs, self.param_mc, self._h2 = _symm_alloc(total, torch.bfloat16, self.device, grp)
            self.flags, self.flag_ptrs, _, self._h3 = _symm_alloc(_SIGNAL_WORDS, torch.int32, self.device, grp)
            self.grad.zero_()
            self.param.zero_()
            self.flags.zero_()
            torch.cuda.synchronize(self.device)
            dist.barrier(group=grp)
            self._group = grp
            self.red = torch.zeros(layout.shard_cap, dtype=torch.float32, device=self.device)
        else:
            self.grad = torch.zeros(total, dtype=torch.float32, device=self.device)
            self.param = torch.zeros(total, dtype=torch.bfloat16, device=self.device)
            self.flags = torch.zeros(_SIGNAL_WORDS, dtype=torch.int32, device=self.device)
            self.grad_ptrs, self.param_ptrs, self.flag_ptrs = [self.grad.data_ptr()], [self.param.data_ptr()], [self.flags.data_ptr()]
            self.grad_mc = self.param_mc = 0
            self._group = None
            s0 = layout.shard_start[rank]
            self.red = self.grad[s0:s0 + layout.shard_cap]        # one rank: the "reduced" gradient is the gradient
        if not self.use_nvls:
            self.grad_mc = self.param_mc = 0
        self.buffers = {"grad": self.grad, "param": self.param}
        cap = layout.shard_cap
        self.m1 = torch.zeros(cap, dtype=torch.float32, device=self.device)
        self.m2 = torch.zeros(cap, dtype=torch.float32, device=self.device)
        self.avg: Optional[torch.Tensor] = None
        if optimizer is not None and getattr(optimizer, "averages", None) is not None:
            self.avg = torch.zeros(cap, dtype=torch.float32, device=self.device)
        self.opt_mode = optimizer_mode(optimizer)
        self.norms = torch.zeros(max(1, len(layout.owned_keys(rank))), dtype=torch.float32, device=self.device)
        self.step_t = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.epoch = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.bar = torch.zeros(2 * _MAX_BUCKETS, dtype=torch.int32, device=self.device)
        self.error = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.hyper = torch.zeros(8, dtype=torch.float32, device=self.device)
        self._hyper_host: Optional[List[float]] = None
        self._sms = torch.cuda.get_device_properties(self.device).multi_processor_count
        self._grid_override = grid
        self.master: Optional[torch.Tensor] = None
        self._steps_since_check = 0
        self._stream: Optional[torch.cuda.Stream] = None
        self._sig_stream: Optional[torch.cuda.Stream] = None       # grad-ready signals (never behind an exchange)
        self._zero_stream: Optional[torch.cuda.Stream] = None      # gate + clear of the gradient accumulators
        self._zero_event: Optional[torch.cuda.Event] = None
        self._zero_waited: set = set()
        self.ext: Optional[Dict[str, Any]] = None
        # bucket plan: built at the first exchange from the gradient-completion order of the first step
        self.plan: Optional[BucketPlan] = None
        self.tables: Optional[Dict[str, Any]] = None
        self._ptr_bucket: Dict[int, int] = {}
        self._all_mask = 0
        self._waited = 0xFFFFFFFF            # bucket bits already waited for since the last exchange
        self._next = 0                       # next bucket to launch in this step
        self._expected: List[set] = []
        self._seen: set = set()
        self._events: Dict[int, Dict[int, Any]] = {}
        self._hooked = False

    # ------------------------------------------------------------------ wiring
    def bind(self, proxy) -> None:
        """Adopt the proxy's fp32 master shard and expose Adam moments / averages per key on
        the optimizer object (so checkpoints and evaluation see them)."""
        self.master = proxy.master
        layout, s0 = self.layout, self.layout.shard_start[self.rank]
        opt = self.optimizer
        if opt is not None and hasattr(opt, "mom1"):
            for k in layout.owned_keys(self.rank):
                o, n = layout.offset[k] - s0, layout.numel[k]
                opt.mom1[k] = self.m1[o:o + n].view(layout.shape[k])
                opt.mom2[k] = self.m2[o:o + n].view(layout.shape[k])
                if self.avg is not None and opt.averages is not None:
                    opt.averages[k] = self.avg[o:o + n].view(layout.shape[k])
                opt.nr_update.setdefault(k, 0)

    def load_optimizer_state(self, nr_update: int, master: Optional[Dict[KeyT, torch.Tensor]] = None) -> None:
        """Resume: the moments were copied INTO the ``m1`` / ``m2`` views by ``Optimizer.load_state_dict``;
        restore the device-side update counter (bias correction) and the fp32 master weights."""
        self.step_t.fill_(int(nr_update))
        if master and self.master is not None:
            layout, s0 = self.layout, self.layout.shard_start[self.rank]
            for k, v in master.items():
                if k in layout.offset and layout.owner[k] == self.rank:
                    o, n = layout.offset[k] - s0, layout.numel[k]
                    self.master[o:o + n].copy_(v.to(self.device, torch.float32).reshape(-1))

    def master_state(self) -> Dict[KeyT, torch.Tensor]:
        layout, s0 = self.layout, self.layout.shard_start[self.rank]
        out = {}
        if self.master is not None:
            for k in layout.owned_keys(self.rank):
                o, n = layout.offset[k] - s0, layout.numel[k]
                out[k] = self.master[o:o + n].view(layout.shape[k]).detach().to("cpu")
        return out

    def enable_trace(self) -> torch.Tensor:
        if self.trace is None:
            self.trace = torch.zeros(_MAX_BUCKETS * 8 + 64, dtype=torch.int64, device=self.device)
        return self.trace

    def stamp(self, slot: int) -> None:
        """Record %globaltimer on the current stream into free trace slot ``slot`` (0..63)."""
        if self.trace is not None:
            torch.ops.srb.stamp(self.trace, _MAX_BUCKETS * 8 + int(slot))

    def _sync_hyper(self) -> None:
        opt = self.optimizer
        vals = [float(opt.learn_rate), float(opt.b1), float(opt.b2), float(opt.eps), float(opt.grad_clip or 0.0),
                float(opt.L2), 1.0 if opt.L2_is_weight_decay else 0.0, 1.0]
        if vals != self._hyper_host:
            self.hyper.copy_(torch.tensor(vals, dtype=torch.float32), non_blocking=False)
            self._hyper_host = vals

    def _comm_stream(self) -> "torch.cuda.Stream":
        if self._stream is None:
            self._stream = torch.cuda.Stream(device=self.device, priority=int(os.environ.get("SRB_COMM_PRIO", "0")))
            self._sig_stream = torch.cuda.Stream(device=self.device, priority=-1)
            self._zero_stream = torch.cuda.Stream(device=self.device, priority=0)
        return self._stream

    # ------------------------------------------------------------------ plan
    def set_order(self, order: Sequence[KeyT], proxy=None) -> None:
        """Build the bucket plan from the order in which the first step's gradients completed."""
        self.plan = plan_buckets(order, self.layout, self.n_buckets_target)
        self.tables = shard_tables(self.layout, self.rank, self.device, self.plan)
        self._all_mask = (1 << self.plan.n) - 1
        had_grad = set(order)
        self._expected = [set(k for k in ks if k in had_grad) for ks in self.plan.buckets]
        # per bucket: the extents of ALL its keys (any owner) in my gradient buffer - what
        # bucket_gate_zero_kernel clears once the owners have published (adjacent extents merged)
        offs, lens, ranges = [], [], []
        for ks in self.plan.buckets:
            e0 = len(offs)
            for o, n in sorted((self.layout.offset[k], _round_up(self.layout.numel[k], ALIGN)) for k in ks):
                if len(offs) > e0 and offs[-1] + lens[-1] == o:
                    lens[-1] += n
                else:
                    offs.append(o)
                    lens.append(n)
            ranges.append((e0, len(offs)))
        self.ext = {"off": torch.tensor(offs, dtype=torch.int64, device=self.device),
                    "len": torch.tensor(lens, dtype=torch.int64, device=self.device), "ranges": ranges,
                    "elems": [sum(lens[a:b]) for a, b in ranges]}
        self._ptr_bucket = {}
        if proxy is not None:
            for k, b in self.plan.bucket_of.items():
                v = proxy._views_p.get(k)
                if v is not None:
                    self._ptr_bucket[int(v.data_ptr())] = b

    def kernels_for(self, b: int) -> int:
        """Launches bucket ``b`` costs on this rank: with peers a signal and a wait kernel (one warp
        each), then reduce + update; a rank that owns nothing of the bucket only signals and waits."""
        bb, be, _kb, _ke = self.tables["ranges"][b]
        peers = 2 if self.world_size > 1 else 0
        if be == bb:
            return max(peers, 1)
        return 2 + peers

    def _grid_for(self, b: int) -> int:
        """One CTA per 4096-element work item (the launcher derives the grid from the item range)."""
        bb, be, _kb, _ke = self.tables["ranges"][b]
        return max(1, be - bb)

    # ------------------------------------------------------------------ the step
    def begin_step(self, proxy, overlap: bool) -> None:
        """Arm the per-bucket completion tracking; with ``overlap`` every ``inc_grad`` that completes
        a bucket launches its exchange at once (``key_ready``), under the rest of the backward pass."""
        self._next = 0
        self._seen = set()
        self._events = {}
        self._signalled = set()
        self._sig_events = {}
        self._hooked = bool(overlap and self.overlap and self.plan is not None)
        if not torch.cuda.is_current_stream_capturing():
            self._sync_hyper()                    # learning-rate schedules: the kernels read the device copy
        self._zero_event = None
        self._zero_waited = set()
        if self.world_size > 1 and self.plan is not None:
            # Clear my gradient accumulators bucket by bucket as soon as every owner has published the
            # bucket (= is done reading them): side stream, under the forward pass.  The first gradient
            # write of the step waits for the event (ensure_zeroed).
            cur = torch.cuda.current_stream(self.device)
            self._comm_stream()
            zs = self._zero_stream
            zs.wait_stream(cur)
            with torch.cuda.stream(zs):
                for b in range(self.plan.n):
                    e0, e1 = self.ext["ranges"][b]
                    if e1 == e0:
                        continue
                    gate = [int(self.flags.data_ptr()), int(self.epoch.data_ptr()), int(self.error.data_ptr()),
                            1 << b, int(self.world_size), int(self.timeout_s * 1000)]
                    grid = max(1, min(2 * self._sms, self.ext["elems"][b] // 8192))
                    torch.ops.srb.gate_zero(self.grad, gate, self.ext["off"], self.ext["len"], int(e0), int(e1), int(grid))
                    self.launches += 1
                ev = torch.cuda.Event()
                ev.record(zs)
            self._zero_event = ev

    def ensure_zeroed(self) -> None:
        """The current stream is about to write gradients: order it after this step's accumulator clear."""
        ev = self._zero_event
        if ev is None:
            return
        cur = torch.cuda.current_stream(self.device)
        if cur.cuda_stream not in self._zero_waited:
            cur.wait_event(ev)
            self._zero_waited.add(cur.cuda_stream)

    def key_ready(self, key: KeyT, proxy) -> None:
        """Called by the proxy from ``inc_grad`` (overlap mode): ``key``'s gradient for this step
        is final once the work enqueued so far on the current stream has run."""
        if not self._hooked or key in self._seen:
            return
        plan = self.plan
        b = plan.bucket_of.get(key)
        if b is None:
            return
        self._seen.add(key)
        if self.trace is not None and len(self._seen) <= 48:
            self.stamp(8 + len(self._seen) - 1)       # trace: when the backward pass completed this key
            self.trace_keys[len(self._seen) - 1] = key
        cur = torch.cuda.current_stream(self.device)
        ev = torch.cuda.Event()
        ev.record(cur)
        self._events.setdefault(b, {})[cur.cuda_stream] = ev
        # buckets are launched in plan order on every rank (a rank that launched them in a different
        # order than its peers would deadlock on the grad-ready flags)
        if self._expected[b] and self._expected[b] <= self._seen and b not in self._signalled:
            self._signal(b, proxy)               # my gradients of bucket b exist: tell the owners NOW
        while self._next < plan.n and self._expected[self._next] and self._expected[self._next] <= self._seen:
            self._launch(self._next, proxy)
            self._next += 1

    def _bucket_args(self, b: int, mode: int):
        t, L = self.tables, self.layout
        bb, be, kb, ke = t["ranges"][b]
        return (self.grad_ptrs, self.param_ptrs, self.flag_ptrs, int(self.grad_mc), int(self.param_mc),
                self.red, self.master, self.m1, self.m2, self.avg, self.norms,
                t["blk_key"], t["blk_off"], t["key_off"], t["key_len"],
                self.hyper, self.step_t, self.epoch, self.bar, self.error,
                int(L.shard_start[self.rank]), int(bb), int(be), int(kb), int(ke), int(b),
                bool(b == self.plan.n - 1), self.rank, int(mode), int(self.opt_mode), float(self.timeout_s),
                int(self.test_delay_us), self.trace)

    def _wait_producers(self, stream, b: int, proxy) -> None:
        evs = self._events.get(b)
        if evs:
            for ev in evs.values():
                stream.wait_event(ev)
        else:
            stream.wait_stream(torch.cuda.current_stream(self.device))
        side_fn = getattr(self.ops, "side_stream_if_pending", None)
        side = side_fn() if side_fn is not None else None
        if side is not None:
            stream.wait_stream(side)         # weight-gradient GEMMs accumulate into the bucket on the ops' side stream

    def _signal(self, b: int, proxy) -> None:
        """Grad-ready flag of bucket ``b`` to every rank, on the signal stream, ordered only after the
        kernels that produced the gradients - never after another bucket's exchange."""
        if self.master is None:
            self.bind(proxy)
        self._comm_stream()
        ss = self._sig_stream
        self._wait_producers(ss, b, proxy)
        with torch.cuda.stream(ss):
            torch.ops.srb.fused_comm_bucket(*self._bucket_args(b, 1))
            ev = torch.cuda.Event()
            ev.record(ss)
        self._sig_events[b] = ev
        self._signalled.add(b)
        if self.world_size > 1 or self.trace is not None:
            self.launches += 1

    def _launch(self, b: int, proxy) -> None:
        """Exchange of bucket ``b`` on the comm stream: wait for the peers' flags, reduce, update, publish."""
        if b not in self._signalled:
            self._signal(b, proxy)
        cs = self._comm_stream()
        cs.wait_event(self._sig_events[b])       # includes the producers of my own gradients
        self._events.pop(b, None)
        with torch.cuda.stream(cs):
            torch.ops.srb.fused_comm_bucket(*self._bucket_args(b, 2))
        self.launches += self.kernels_for(b) - (1 if self.world_size > 1 else 0)

    def fused_step(self, proxy) -> None:
        if self.master is None:
            self.bind(proxy)
        if self.plan is None:
            self.set_order(proxy.take_grad_order(), proxy)
        if not torch.cuda.is_current_stream_capturing():
            self._sync_hyper()
        cur = torch.cuda.current_stream(self.device)
        while self._next < self.plan.n:
            self._launch(self._next, proxy)
            self._next += 1
        cur.wait_stream(self._comm_stream())
        cur.wait_stream(self._sig_stream)
        if self._zero_event is not None:
            cur.wait_event(self._zero_event)         # (already reached by every gradient writer; joins the capture)
        self._hooked = False
        self._waited = 0
        if self.terminal_wait and self.world_size > 1:
            self.quiesce()
        if not torch.cuda.is_current_stream_capturing():
            self.host_bookkeeping()

    # ------------------------------------------------------------------ consumer gates (C2)
    def gate_for(self, tensors: Sequence[torch.Tensor]) -> List[int]:
        """Gate descriptor for a kernel that is about to read ``tensors`` (parameter views): the
        buckets among them that have not been waited for since the last exchange.  ``[]`` = none."""
        if self.world_size == 1 or self.plan is None:
            return []
        mask = 0
        pb = self._ptr_bucket
        for t in tensors:
            if t is None:
                continue
            b = pb.get(int(t.data_ptr()))
            if b is not None:
                mask |= 1 << b
        need = mask & ~self._waited
        if not need:
            return []
        self._waited |= need
        return [int(self.flags.data_ptr()), int(self.epoch.data_ptr()), int(self.error.data_ptr()), int(need),
                int(self.world_size), int(self.timeout_s * 1000)]

    def reset_gates(self) -> None:
        """Forget what has been waited for (before capturing a step, so the gates are part of the graph)."""
        self._waited = 0

    def quiesce(self) -> None:
        """Stand-alone gate on every bucket (one warp): all peers' weights of the last exchange have
        landed.  For host-side readers (checkpoint, evaluation through library ops)."""
        if self.world_size == 1 or self.plan is None:
            return
        need = self._all_mask & ~self._waited
        if not need:
            return
        self._waited |= need
        torch.ops.srb.gate_wait(self.epoch, [int(self.flags.data_ptr()), int(self.epoch.data_ptr()),
                                             int(self.error.data_ptr()), int(need), int(self.world_size),
                                             int(self.timeout_s * 1000)])

    def host_bookkeeping(self) -> None:
        """Host-side mirror of one executed step (update counters used by checkpoints / bias
        correction on resume, periodic error-flag check).  ``fused_step`` calls it when it runs
        eagerly; ``engine.Trainer`` calls it after every CUDA-graph replay of a captured step."""
        opt = self.optimizer
        if opt is not None and hasattr(opt, "nr_update"):
            for k in self.layout.owned_keys(self.rank):
                opt.nr_update[k] = opt.nr_update.get(k, 0) + 1
        self._waited = 0
        self._steps_since_check += 1
        if self._steps_since_check >= 64:
            self.check()

    def check(self) -> None:
        """Raise if any spin-wait in the kernels timed out (a peer died or hung)."""
        self._steps_since_check = 0
        code = int(self.error.item())
        if code != 0:
            what = {1: "waiting for the peers' gradients", 2: "grid barrier after the reduce phase",
                    3: "grid barrier after the update phase", 7: "a consumer gate waiting for published weights"}
            raise RuntimeError(
                f"fused comm kernel on rank {self.rank} timed out ({what.get(code, 'phase ' + str(code))}; "
                f"code {code}): a peer did not arrive within {self.timeout_s}s"
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
