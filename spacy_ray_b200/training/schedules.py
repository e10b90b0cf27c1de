"""Hyper-parameter schedules (thinc names: ``compounding.v1`` etc.)."""
from __future__ import annotations

import itertools
from typing import Iterator

from ..config import registry


@registry.schedules("constant.v1")
def constant(rate: float) -> Iterator[float]:
    return itertools.repeat(float(rate))


@registry.schedules("compounding.v1")
def compounding(start: float, stop: float, compound: float, t: float = 0.0) -> Iterator[float]:
    """start * compound**k, clipped at ``stop`` (works for growth and decay)."""
    def gen():
        curr = float(start)
        while True:
            yield min(curr, stop) if start <= stop else max(curr, stop)
            curr *= compound
    return gen()


@registry.schedules("decaying.v1")
def decaying(base_rate: float, decay: float, t: int = 0) -> Iterator[float]:
    def gen():
        step = t
        while True:
            yield base_rate * (1.0 / (1.0 + decay * step))
            step += 1
    return gen()


@registry.schedules("warmup_linear.v1")
def warmup_linear(initial_rate: float, warmup_steps: int, total_steps: int) -> Iterator[float]:
    def gen():
        step = 0
        while True:
            if step < warmup_steps:
                factor = step / max(1, warmup_steps)
            else:
                factor = max(0.0, (total_steps - step) / max(1.0, total_steps - warmup_steps))
            yield factor * initial_rate
            step += 1
    return gen()


@registry.schedules("slanted_triangular.v1")
def slanted_triangular(max_rate: float, num_steps: int, cut_frac: float = 0.1, ratio: int = 32, t: float = 0.0) -> Iterator[float]:
    def gen():
        cut = int(num_steps * cut_frac)
        step = int(t)
        while True:
            step += 1
            if step < cut:
                p = step / max(1, cut)
            else:
                p = 1 - ((step - cut) / max(1.0, cut * (1 / cut_frac - 1)))
            yield max_rate * (1 + p * (ratio - 1)) * (1 / ratio)
    return gen()
