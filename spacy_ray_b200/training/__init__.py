from . import schedules, batchers, corpus, loggers, optimizer  # noqa: F401  (registers functions)
from .optimizer import Optimizer, Adam
from .loop import (
    train_while_improving, create_train_batches, create_evaluation_callback,
    create_before_to_disk_callback, update_meta, subdivide_batch, ConfigSchemaTraining,
)
from .initialize import init_nlp
from .corpus import SyntheticCorpus, JsonlCorpus
from .checkpoint import save_pipeline, save_optimizer_shard, load_optimizer_shards, structural_keys

__all__ = [
    "Optimizer", "Adam", "train_while_improving", "create_train_batches", "create_evaluation_callback",
    "create_before_to_disk_callback", "update_meta", "subdivide_batch", "ConfigSchemaTraining", "init_nlp",
    "SyntheticCorpus", "JsonlCorpus", "save_pipeline", "save_optimizer_shard", "load_optimizer_shards",
    "structural_keys",
]
