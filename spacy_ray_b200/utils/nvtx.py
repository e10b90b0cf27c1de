"""NVTX ranges (visible in Nsight Systems / ncu) - no-ops without CUDA."""
from __future__ import annotations

import contextlib

import torch


@contextlib.contextmanager
def nvtx_range(name: str):
    on = torch.cuda.is_available()
    if on:
        torch.cuda.nvtx.range_push(name)
    try:
        yield
    finally:
        if on:
            torch.cuda.nvtx.range_pop()
