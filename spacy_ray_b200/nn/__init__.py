from .model import Model, ParamServer, KeyT, reset_model_ids, iter_param_keys
from .batch import TokenBatch, make_token_batch, collate_attrs, padded_rows
from . import layers
from .layers import fix_random_seed, set_dropout_rate

__all__ = [
    "Model", "ParamServer", "KeyT", "reset_model_ids", "iter_param_keys", "TokenBatch",
    "make_token_batch", "collate_attrs", "padded_rows", "layers", "fix_random_seed", "set_dropout_rate",
]
