"""Optimizer with thinc's call contract: ``optimizer(key, weights, gradient)``.

The reference calls exactly this from inside its proxy, once per parameter key
(``/root/reference/spacy_ray/proxies.py:128``).  Semantics reproduced (SURVEY.md
appendix A): per-key update counter, optional L2 folded into the gradient,
per-tensor gradient-norm clipping, Adam with bias correction folded into the
learning rate, decoupled weight decay, optional parameter averaging,
``step_schedules``.  State (fp32 moments, counters) is serialisable so that a
run can resume - the reference never saves it.
"""
from __future__ import annotations

from typing import Any, Dict, Iterator, Optional, Tuple, Union

import torch

from ..config import registry
from ..ops import get_current_ops

KeyT = Tuple[int, str]
Schedule = Union[float, Iterator[float]]


def _first(value: Schedule) -> Tuple[float, Optional[Iterator[float]]]:
    if isinstance(value, (int, float)):
        return float(value), None
    it = iter(value)
    return float(next(it)), it


class Optimizer:
    def __init__(
        self,
        learn_rate: Schedule = 0.001,
        *,
        L2: Schedule = 0.0,
        beta1: Schedule = 0.9,
        beta2: Schedule = 0.999,
        eps: Schedule = 1e-8,
        grad_clip: Schedule = 1.0,
        use_averages: bool = False,
        use_radam: bool = False,
        L2_is_weight_decay: bool = True,
        use_adam: bool = True,
        ops: Any = None,
    ):
        self.ops = ops
        self._schedules: Dict[str, Iterator[float]] = {}
        for name, value in (("learn_rate", learn_rate), ("L2", L2), ("b1", beta1), ("b2", beta2),
                            ("eps", eps), ("grad_clip", grad_clip)):
            v, it = _first(value)
            setattr(self, name, v)
            if it is not None:
                self._schedules[name] = it
        self.use_adam = use_adam
        self.use_radam = use_radam
        self.L2_is_weight_decay = L2_is_weight_decay
        self.mom1: Dict[KeyT, torch.Tensor] = {}
        self.mom2: Dict[KeyT, torch.Tensor] = {}
        self.nr_update: Dict[KeyT, int] = {}
        self.averages: Optional[Dict[KeyT, torch.Tensor]] = {} if use_averages else None
        self.last_seen: Dict[KeyT, int] = {}
        self.master: Dict[KeyT, torch.Tensor] = {}   # fp32 master copies for low-precision (bf16) weights
        self.step = 0

    # ---- schedules -------------------------------------------------------
    def step_schedules(self) -> None:
        """Advance lr / other schedules by one step.  (The reference's
        ``FakeOptimizer.step_schedules`` is a no-op and its proxy never calls the
        real one - SURVEY.md 2.1 #7 - so schedules never advance there; our
        training loop calls this once per step.)"""
        for name, it in self._schedules.items():
            setattr(self, name, float(next(it)))
        self.step += 1

    # ---- the update ------------------------------------------------------
    def __call__(self, key: KeyT, weights: torch.Tensor, gradient: torch.Tensor, *, lr_scale: float = 1.0):
        if weights.numel() == 0:
            return weights, gradient
        ops = self.ops or get_current_ops()
        self.nr_update[key] = self.nr_update.get(key, 0) + 1
        nr = self.nr_update[key]
        if weights.dtype == torch.float32:
            w32 = weights
        else:
            w32 = self.master.get(key)
            if w32 is None or w32.shape != weights.shape:
                w32 = self.master[key] = weights.to(torch.float32).clone()
        if self.use_adam and not self.use_radam:
            if key not in self.mom1:
                self.mom1[key] = torch.zeros_like(w32)
                self.mom2[key] = torch.zeros_like(w32)
            ops.adam_step(
                w32, gradient, self.mom1[key], self.mom2[key],
                lr=self.learn_rate * lr_scale, beta1=self.b1, beta2=self.b2, eps=self.eps,
                nr_update=nr, grad_clip=self.grad_clip, l2=self.L2,
                l2_is_weight_decay=self.L2_is_weight_decay,
            )
        else:
            g = gradient.to(torch.float32)
            if self.L2 != 0.0 and not self.L2_is_weight_decay:
                g = g + self.L2 * w32
            if self.grad_clip:
                norm = torch.linalg.vector_norm(g)
                scale = torch.where(norm >= self.grad_clip, self.grad_clip / norm.clamp_min(1e-30),
                                    torch.ones_like(norm))      # no host sync
                g = g * scale
            lr = self.learn_rate * lr_scale
            if self.use_adam:                     # RAdam (thinc ``Optimizer._radam``)
                if key not in self.mom1:
                    self.mom1[key] = torch.zeros_like(w32)
                    self.mom2[key] = torch.zeros_like(w32)
                m1, m2 = self.mom1[key], self.mom2[key]
                m2.mul_(self.b2).addcmul_(g, g, value=1.0 - self.b2)
                m1.mul_(self.b1).add_(g, alpha=1.0 - self.b1)
                step_size, rect = radam_step_size(nr, self.b1, self.b2)
                if rect:
                    w32.addcdiv_(m1, m2.sqrt().add_(self.eps), value=-lr * step_size)
                else:
                    w32.add_(m1, alpha=-lr * step_size)
            else:
                w32.add_(g, alpha=-lr)
            if self.L2 != 0.0 and self.L2_is_weight_decay:
                w32.mul_(1.0 - self.learn_rate * self.L2)
            gradient.zero_()
        if w32 is not weights:
            weights.copy_(w32)
        if self.averages is not None:
            # thinc ``update_averages``: the running average starts at zero
            avg = self.averages.get(key)
            if avg is None:
                avg = self.averages[key] = torch.zeros_like(w32)
            decay = min((1.0 + nr) / (10.0 + nr), 0.9999)
            avg.sub_(avg - w32, alpha=1.0 - decay)
        return weights, gradient

    # ---- state -----------------------------------------------------------
    def state_dict(self, keys=None) -> Dict[str, Any]:
        sel = (lambda k: True) if keys is None else (lambda k, s=set(keys): k in s)
        return {
            "mom1": {k: v.detach().to("cpu") for k, v in self.mom1.items() if sel(k)},
            "mom2": {k: v.detach().to("cpu") for k, v in self.mom2.items() if sel(k)},
            "nr_update": {k: v for k, v in self.nr_update.items() if sel(k)},
            "averages": None if self.averages is None else {k: v.detach().to("cpu") for k, v in self.averages.items() if sel(k)},
            "master": {k: v.detach().to("cpu") for k, v in self.master.items() if sel(k)},
            "step": self.step,
            "hyper": {"learn_rate": self.learn_rate, "L2": self.L2, "b1": self.b1, "b2": self.b2,
                      "eps": self.eps, "grad_clip": self.grad_clip},
        }

    def load_state_dict(self, state: Dict[str, Any], device=None) -> None:
        """Entries that already exist with the same shape are COPIED INTO (they may be views of a
        fused kernel's flat moment buffers - ``FusedSymmComm.bind`` - and rebinding the dict entry
        would silently detach the kernel from the restored state)."""
        dev = device or (self.ops.device if self.ops is not None else get_current_ops().device)

        def merge(dst: Dict[KeyT, torch.Tensor], src: Dict[KeyT, torch.Tensor]) -> None:
            for k, v in src.items():
                cur = dst.get(k)
                if cur is not None and tuple(cur.shape) == tuple(v.shape):
                    cur.copy_(v.to(device=cur.device, dtype=cur.dtype))
                else:
                    dst[k] = v.to(dev)

        merge(self.mom1, state["mom1"])
        merge(self.mom2, state["mom2"])
        self.nr_update.update(state["nr_update"])
        if state.get("averages") is not None:
            if self.averages is None:
                self.averages = {}
            merge(self.averages, state["averages"])
        if state.get("master"):
            merge(self.master, state["master"])
        target = int(state.get("step", 0))
        while self.step < target:                 # fast-forward learning-rate (and other) schedules
            self.step_schedules()


def radam_step_size(t: int, beta1: float, beta2: float):
    """Rectified-Adam step multiplier at update ``t`` -> (step_size, rectified?).  When the variance
    of the adaptive rate is not tractable yet (N_sma < 5) the update degenerates to bias-corrected
    momentum SGD, exactly like thinc's ``_radam`` (``degenerated_to_sgd``)."""
    import math

    beta2_t = beta2 ** t
    sma_max = 2.0 / (1.0 - beta2) - 1.0
    sma = sma_max - 2.0 * t * beta2_t / (1.0 - beta2_t)
    if sma >= 5.0:
        return (math.sqrt((1.0 - beta2_t) * (sma - 4.0) / (sma_max - 4.0) * (sma - 2.0) / sma * sma_max / (sma_max - 2.0))
                / (1.0 - beta1 ** t)), True
    return 1.0 / (1.0 - beta1 ** t), False


@registry.optimizers("Adam.v1")
def Adam(
    learn_rate: Schedule = 0.001,
    *,
    L2: Schedule = 0.0,
    beta1: Schedule = 0.9,
    beta2: Schedule = 0.999,
    eps: Schedule = 1e-8,
    grad_clip: Schedule = 1.0,
    L2_is_weight_decay: bool = True,
    use_averages: bool = True,
) -> Optimizer:
    return Optimizer(learn_rate, L2=L2, beta1=beta1, beta2=beta2, eps=eps, grad_clip=grad_clip,
                     L2_is_weight_decay=L2_is_weight_decay, use_averages=use_averages)


@registry.optimizers("RAdam.v1")
def RAdam(learn_rate: Schedule = 0.001, *, L2: Schedule = 0.0, beta1: Schedule = 0.9, beta2: Schedule = 0.999,
          eps: Schedule = 1e-8, grad_clip: Schedule = 1.0, L2_is_weight_decay: bool = True,
          use_averages: bool = True) -> Optimizer:
    return Optimizer(learn_rate, L2=L2, beta1=beta1, beta2=beta2, eps=eps, grad_clip=grad_clip,
                     L2_is_weight_decay=L2_is_weight_decay, use_averages=use_averages, use_radam=True)


@registry.optimizers("SGD.v1")
def SGD(learn_rate: Schedule = 0.001, *, L2: Schedule = 0.0, grad_clip: Schedule = 1.0,
        L2_is_weight_decay: bool = True, use_averages: bool = True) -> Optimizer:
    return Optimizer(learn_rate, L2=L2, grad_clip=grad_clip, L2_is_weight_decay=L2_is_weight_decay,
                     use_averages=use_averages, use_adam=False)
