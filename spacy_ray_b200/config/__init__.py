"""Config system + function registry.

The reference loads a thinc/spaCy ``.cfg`` (``/root/reference/spacy_ray/train_cli.py:44-46``)
and resolves ``[training]`` through spaCy's registry
(``/root/reference/spacy_ray/worker.py:91-95``).  Neither thinc's ``Config``
nor ``catalogue``/``confection`` are available here, so this module provides
an independent implementation of the same *file format*:

* INI sections with dotted nesting (``[components.ner.model.tok2vec]``)
* JSON-typed values
* ``@registry = "name.v1"`` keys: "call the registered function with the
  sibling keys as keyword arguments"
* ``${section.key}`` interpolation (whole-value and inside strings) and
  ``${section}`` (whole-section) interpolation
* dotted CLI overrides (``--training.max_steps 100``)
"""
from .config import Config, ConfigValidationError, load_config, parse_config_overrides
from .registry import registry, Registry, resolve, resolve_dot_names, fill_defaults

__all__ = [
    "Config",
    "ConfigValidationError",
    "load_config",
    "parse_config_overrides",
    "registry",
    "Registry",
    "resolve",
    "resolve_dot_names",
    "fill_defaults",
]
