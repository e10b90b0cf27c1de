"""Function registry + config resolution.

A block such as::

    [training.optimizer]
    @optimizers = "Adam.v1"
    learn_rate = 0.001

resolves to ``registry.optimizers.get("Adam.v1")(learn_rate=0.001)``; nested
blocks resolve depth first.  Equivalent of the ``registry.resolve`` the
reference calls at ``/root/reference/spacy_ray/worker.py:93`` and the
``resolve_dot_names`` at ``worker.py:95``.
"""
from __future__ import annotations

import inspect
from typing import Any, Callable, Dict, List, Mapping, Optional, Sequence, Tuple

from .config import Config, ConfigValidationError


class _Namespace:
    def __init__(self, name: str):
        self.name = name
        self._fns: Dict[str, Callable] = {}

    def register(self, name: str, func: Optional[Callable] = None):
        if func is not None:
            self._fns[name] = func
            return func

        def deco(fn: Callable) -> Callable:
            self._fns[name] = fn
            return fn

        return deco

    __call__ = register

    def get(self, name: str) -> Callable:
        if name not in self._fns:
            avail = ", ".join(sorted(self._fns)) or "(none)"
            raise ConfigValidationError(
                f"Can't find '{name}' in registry '{self.name}'. Available: {avail}"
            )
        return self._fns[name]

    def has(self, name: str) -> bool:
        return name in self._fns

    def get_all(self) -> Dict[str, Callable]:
        return dict(self._fns)


class Registry:
    """Namespaces match the ones spaCy/thinc configs use, so an upstream config
    file's ``@architectures`` / ``@optimizers`` / ... keys resolve here."""

    NAMES = (
        "architectures",
        "optimizers",
        "schedules",
        "batchers",
        "loggers",
        "readers",
        "callbacks",
        "misc",
        "factories",
        "scorers",
        "layers",
        "initializers",
        "tokenizers",
        "augmenters",
    )

    def __init__(self):
        for n in self.NAMES:
            setattr(self, n, _Namespace(n))

    def namespace(self, name: str) -> _Namespace:
        if name not in self.NAMES:
            raise ConfigValidationError(f"Unknown registry '@{name}'")
        return getattr(self, name)

    def resolve(self, config: Mapping, *, schema: Any = None, validate: bool = True) -> Dict[str, Any]:
        return resolve(config, schema=schema, validate=validate)

    def fill(self, config: Mapping, *, schema: Any = None) -> Config:
        return fill_defaults(config, schema=schema)


registry = Registry()


def _registry_key(block: Mapping) -> Optional[str]:
    keys = [k for k in block if isinstance(k, str) and k.startswith("@")]
    if not keys:
        return None
    if len(keys) > 1:
        raise ConfigValidationError(f"Block has more than one @registry key: {keys}")
    return keys[0]


def _call(func: Callable, kwargs: Dict[str, Any], where: str) -> Any:
    args: Sequence[Any] = ()
    if "*" in kwargs:  # thinc positional-args block
        star = kwargs.pop("*")
        args = list(star.values()) if isinstance(star, Mapping) else list(star)
    try:
        sig = inspect.signature(func)
        sig.bind(*args, **kwargs)
    except TypeError as e:
        raise ConfigValidationError(
            f"Bad arguments for {getattr(func, '__name__', func)} at [{where}]", desc=str(e)
        ) from None
    return func(*args, **kwargs)


def _resolve_node(node: Any, where: str) -> Any:
    if isinstance(node, Mapping):
        reg_key = _registry_key(node)
        resolved = {
            k: _resolve_node(v, f"{where}.{k}" if where else k)
            for k, v in node.items()
            if k != reg_key
        }
        if reg_key is None:
            return resolved
        func = registry.namespace(reg_key[1:]).get(node[reg_key])
        return _call(func, resolved, where)
    if isinstance(node, list):
        return [_resolve_node(v, where) for v in node]
    return node


def resolve(config: Mapping, *, schema: Any = None, validate: bool = True) -> Dict[str, Any]:
    """Resolve every ``@registry`` block in ``config`` (must already be
    interpolated).  ``schema`` is an optional pydantic model class: defaults are
    filled in *before* resolution (so default blocks with ``@`` keys resolve too)
    and, if ``validate``, unknown top-level keys are rejected."""
    data = dict(config)
    if schema is not None:
        data = dict(fill_defaults(data, schema=schema))
        if validate:
            allowed = set(_schema_fields(schema))
            extra = [k for k in data if k not in allowed]
            if extra:
                raise ConfigValidationError(
                    "Config validation error",
                    [(k, "extra fields not permitted") for k in extra],
                )
    out = _resolve_node(data, "")
    if schema is not None and validate:
        _validate_types(out, schema)
    return out


def _schema_fields(schema: Any) -> List[str]:
    if hasattr(schema, "model_fields"):
        return list(schema.model_fields)
    return list(getattr(schema, "__fields__", {}))


def _schema_defaults(schema: Any) -> Dict[str, Any]:
    out: Dict[str, Any] = {}
    fields = getattr(schema, "model_fields", None) or {}
    for name, f in fields.items():
        if not f.is_required():
            default = f.get_default(call_default_factory=True)
            out[name] = default
    return out


def fill_defaults(config: Mapping, *, schema: Any = None) -> Config:
    data = dict(config)
    if schema is not None:
        for k, v in _schema_defaults(schema).items():
            if k not in data:
                data[k] = v
    return Config(data)


def _validate_types(resolved: Dict[str, Any], schema: Any) -> None:
    """Light validation of scalar fields only (resolved callables/objects are
    passed through untouched)."""
    errors: List[Tuple[str, str]] = []
    fields = getattr(schema, "model_fields", {})
    for name, f in fields.items():
        if name not in resolved:
            if f.is_required():
                errors.append((name, "field required"))
            continue
        ann = f.annotation
        val = resolved[name]
        if ann in (int, float, bool, str) and val is not None:
            if ann is float and isinstance(val, int) and not isinstance(val, bool):
                continue
            if ann is int and isinstance(val, bool):
                errors.append((name, "value is not a valid integer"))
            elif not isinstance(val, ann):
                errors.append((name, f"value is not a valid {ann.__name__}"))
    if errors:
        raise ConfigValidationError("Config validation error", errors)


def resolve_dot_names(config: Mapping, dot_names: Sequence[Optional[str]]) -> List[Any]:
    """``["corpora.train", "corpora.dev"]`` → the resolved objects.  Each top-level
    section is resolved once even if several names point into it."""
    cache: Dict[str, Any] = {}
    out: List[Any] = []
    for name in dot_names:
        if name is None:
            out.append(None)
            continue
        parts = name.split(".")
        section = parts[0]
        if section not in config:
            raise ConfigValidationError(f"Can't resolve '{name}': no [{section}] section")
        if section not in cache:
            cache[section] = _resolve_node(config[section], section)
        node = cache[section]
        for part in parts[1:]:
            try:
                node = node[part]
            except (KeyError, TypeError):
                raise ConfigValidationError(f"Can't resolve '{name}': '{part}' not found") from None
        out.append(node)
    return out
