"""Reference-compatible module path (``spacy_ray.loggers``)."""
from .training.loggers import ray_console_logger, console_logger, jsonl_logger, format_row  # noqa: F401
