"""Device-resident training engine (CUDA-graph step, native collate, pinned staging)."""
from .trainer import ExampleStore, Trainer

__all__ = ["Trainer", "ExampleStore"]
