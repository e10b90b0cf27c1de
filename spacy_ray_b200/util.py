"""Reference-compatible module path (``spacy_ray.util``)."""
from .parallel.util import (  # noqa: F401
    KeyT, Timer, ManyTimer, make_key, divide_params, divide_params_balanced, set_params_proxy, clear_params_proxy,
)
