"""``Config``: a nested dict that round-trips the thinc ``.cfg`` text format."""
from __future__ import annotations

import copy
import io
import json
import re
from configparser import ConfigParser, ExtendedInterpolation, MAX_INTERPOLATION_DEPTH  # noqa: F401
from pathlib import Path
from typing import Any, Dict, Iterable, List, Mapping, Optional, Union

_VAR_RE = re.compile(r"\$\{([A-Za-z0-9_.\-:]+)\}")


class ConfigValidationError(ValueError):
    """Raised for malformed config files, unknown registry functions or bad
    arguments.  Carries a list of ``(location, message)`` pairs so the CLI can
    print them the way ``show_validation_error`` does upstream."""

    def __init__(self, title: str, errors: Optional[List[tuple]] = None, desc: str = ""):
        self.title = title
        self.errors = list(errors or [])
        self.desc = desc
        lines = [title]
        if desc:
            lines.append(desc)
        for loc, msg in self.errors:
            lines.append(f"  {loc}\t{msg}")
        super().__init__("\n".join(lines))


def _parse_value(text: str) -> Any:
    """Values are JSON; anything that fails to parse is kept as a raw string
    (so ``lang = en`` and ``lang = "en"`` both work, like upstream)."""
    text = text.strip()
    if text == "":
        return ""
    try:
        return json.loads(text)
    except ValueError:
        pass
    # python-literal spellings that show up in hand-written configs
    low = text.lower()
    if low == "none":
        return None
    if low == "true":
        return True
    if low == "false":
        return False
    # single-quoted string
    if len(text) >= 2 and text[0] == text[-1] == "'":
        return text[1:-1]
    return text


def _dump_value(value: Any) -> str:
    if isinstance(value, Path):
        value = str(value)
    if isinstance(value, str) and _VAR_RE.fullmatch(value):
        return value  # keep ${a.b} unquoted so it survives a round trip as a reference
    try:
        return json.dumps(value)
    except TypeError:
        return json.dumps(str(value))


class Config(dict):
    """Nested ``dict`` with text (de)serialisation and ``${}`` interpolation.

    ``is_interpolated`` mirrors thinc: a config loaded with
    ``interpolate=False`` keeps its ``${...}`` references verbatim so it can be
    shipped to workers and interpolated there (the reference does exactly this,
    ``train_cli.py:46`` then ``worker.py:92``).
    """

    is_interpolated: bool

    def __init__(self, data: Optional[Mapping] = None, *, is_interpolated: Optional[bool] = None):
        super().__init__()
        if data is not None:
            for k, v in dict(data).items():
                self[k] = _deep_dict(v)
        if is_interpolated is not None:
            self.is_interpolated = is_interpolated
        elif isinstance(data, Config):
            self.is_interpolated = data.is_interpolated
        else:
            self.is_interpolated = not _has_vars(self)

    # ---- text format -----------------------------------------------------
    def from_str(self, text: str, *, interpolate: bool = True, overrides: Optional[Dict[str, Any]] = None) -> "Config":
        parser = ConfigParser(interpolation=None, delimiters=("=",), comment_prefixes=("#", ";"))
        parser.optionxform = str  # keys are case sensitive
        try:
            parser.read_string(text)
        except Exception as e:  # configparser.Error
            raise ConfigValidationError("Can't parse config text", desc=str(e)) from None
        self.clear()
        # Sort so parents are created before children regardless of file order
        for section in sorted(parser.sections(), key=lambda s: s.count(".")):
            node = self
            parts = section.split(".")
            for part in parts:
                if part not in node:
                    node[part] = {}
                elif not isinstance(node[part], dict):
                    raise ConfigValidationError(
                        f"Section [{section}] collides with a value of the same name"
                    )
                node = node[part]
            for key, raw in parser.items(section, raw=True):
                node[key] = _parse_value(raw)
        for dotted, value in (overrides or {}).items():
            _set_dotted(self, dotted, value)
        self.is_interpolated = not _has_vars(self)
        if interpolate and not self.is_interpolated:
            done = self.interpolate()
            self.clear()
            self.update(done)
            self.is_interpolated = True
        return self

    def from_disk(self, path: Union[str, Path], *, interpolate: bool = True, overrides: Optional[Dict[str, Any]] = None) -> "Config":
        with Path(path).open("r", encoding="utf8") as f:
            return self.from_str(f.read(), interpolate=interpolate, overrides=overrides)

    def from_bytes(self, data: bytes, **kw) -> "Config":
        return self.from_str(data.decode("utf8"), **kw)

    def to_str(self) -> str:
        out = io.StringIO()
        # top-level scalars are not representable in INI; they don't occur in practice
        for name, section in self.items():
            if isinstance(section, dict):
                _write_section(out, name, section)
        return out.getvalue().rstrip() + "\n"

    def to_bytes(self) -> bytes:
        return self.to_str().encode("utf8")

    def to_disk(self, path: Union[str, Path]) -> None:
        with Path(path).open("w", encoding="utf8") as f:
            f.write(self.to_str())

    # ---- manipulation ----------------------------------------------------
    def copy(self) -> "Config":  # type: ignore[override]
        return Config(copy.deepcopy(dict(self)), is_interpolated=self.is_interpolated)

    def merge(self, updates: Mapping, *, remove_extra: bool = False) -> "Config":
        """Deep merge ``updates`` over ``self`` (``updates`` wins).  When a block
        switches its ``@registry`` function, the old sibling arguments are dropped
        rather than leaking into the new function's kwargs."""
        merged = _deep_merge(copy.deepcopy(dict(self)), updates)
        return Config(merged)

    def interpolate(self) -> "Config":
        """Return a copy with every ``${a.b}`` replaced by the value it names."""
        data = copy.deepcopy(dict(self))
        resolved = _interpolate_tree(data, data, depth=0)
        return Config(resolved, is_interpolated=True)

    def get_dotted(self, dotted: str, default: Any = None) -> Any:
        node: Any = self
        for part in dotted.split("."):
            if not isinstance(node, dict) or part not in node:
                return default
            node = node[part]
        return node

    def set_dotted(self, dotted: str, value: Any) -> None:
        _set_dotted(self, dotted, value)


# -------------------------------------------------------------------------
def _deep_dict(v: Any) -> Any:
    if isinstance(v, Mapping):
        return {k: _deep_dict(x) for k, x in v.items()}
    return v


def _has_vars(node: Any) -> bool:
    if isinstance(node, str):
        return bool(_VAR_RE.search(node))
    if isinstance(node, Mapping):
        return any(_has_vars(v) for v in node.values())
    if isinstance(node, (list, tuple)):
        return any(_has_vars(v) for v in node)
    return False


def _set_dotted(root: dict, dotted: str, value: Any) -> None:
    parts = dotted.split(".")
    node = root
    for part in parts[:-1]:
        nxt = node.get(part)
        if not isinstance(nxt, dict):
            nxt = {}
            node[part] = nxt
        node = nxt
    node[parts[-1]] = value


def _lookup(root: Mapping, dotted: str) -> Any:
    node: Any = root
    for part in dotted.split("."):
        if isinstance(node, Mapping) and part in node:
            node = node[part]
        else:
            raise ConfigValidationError(f"Can't interpolate ${{{dotted}}}: '{part}' not found")
    return node


def _interpolate_tree(node: Any, root: Mapping, depth: int) -> Any:
    if depth > 16:
        raise ConfigValidationError("Interpolation recursion limit hit (circular ${} reference?)")
    if isinstance(node, dict):
        return {k: _interpolate_tree(v, root, depth) for k, v in node.items()}
    if isinstance(node, list):
        return [_interpolate_tree(v, root, depth) for v in node]
    if isinstance(node, str):
        whole = _VAR_RE.fullmatch(node)
        if whole:
            target = copy.deepcopy(_lookup(root, whole.group(1).replace(":", ".")))
            return _interpolate_tree(target, root, depth + 1)

        def _sub(m: "re.Match") -> str:
            v = _interpolate_tree(_lookup(root, m.group(1).replace(":", ".")), root, depth + 1)
            if isinstance(v, (dict, list)):
                raise ConfigValidationError(
                    f"Can't interpolate section/list ${{{m.group(1)}}} inside a string"
                )
            return "" if v is None else str(v)

        if _VAR_RE.search(node):
            return _sub_all(node, _sub)
    return node


def _sub_all(text: str, fn) -> str:
    return _VAR_RE.sub(fn, text)


def _deep_merge(base: dict, updates: Mapping) -> dict:
    for key, value in updates.items():
        if isinstance(value, Mapping) and isinstance(base.get(key), dict):
            old = base[key]
            reg_old = [k for k in old if k.startswith("@")]
            reg_new = [k for k in value if k.startswith("@")]
            if reg_old and reg_new and (reg_old != reg_new or old[reg_old[0]] != value[reg_new[0]]):
                base[key] = copy.deepcopy(dict(value))
            else:
                base[key] = _deep_merge(old, value)
        else:
            base[key] = copy.deepcopy(value) if isinstance(value, (dict, list)) else value
    return base


def _write_section(out: io.StringIO, name: str, section: Mapping) -> None:
    scalars = [(k, v) for k, v in section.items() if not isinstance(v, Mapping)]
    subs = [(k, v) for k, v in section.items() if isinstance(v, Mapping)]
    out.write(f"[{name}]\n")
    # registry key first, like upstream's writer
    scalars.sort(key=lambda kv: (not kv[0].startswith("@"),))
    for k, v in scalars:
        out.write(f"{k} = {_dump_value(v)}\n")
    out.write("\n")
    for k, v in subs:
        _write_section(out, f"{name}.{k}", v)


# -------------------------------------------------------------------------
def load_config(
    path: Union[str, Path],
    overrides: Optional[Dict[str, Any]] = None,
    interpolate: bool = False,
) -> Config:
    """Load a ``.cfg`` file (``-`` reads stdin).  Mirrors the call the reference
    makes at ``train_cli.py:46``: overrides applied, left un-interpolated by
    default so workers interpolate their own copy."""
    import sys

    if str(path) == "-":
        return Config().from_str(sys.stdin.read(), overrides=overrides, interpolate=interpolate)
    p = Path(path)
    if not p.exists() or not p.is_file():
        raise ConfigValidationError(f"Config file not found: {p}")
    return Config().from_disk(p, overrides=overrides, interpolate=interpolate)


def parse_config_overrides(args: Iterable[str]) -> Dict[str, Any]:
    """Turn leftover CLI args (``--training.max_steps 100 --paths.train=x.jsonl
    --training.flag``) into ``{"training.max_steps": 100, ...}``.

    Same contract as the helper the reference calls at ``train_cli.py:44``:
    keys must be dotted (contain a section), values are JSON-typed, a bare flag
    means ``true``."""
    args = list(args)
    result: Dict[str, Any] = {}
    i = 0
    while i < len(args):
        tok = args[i]
        i += 1
        if not tok.startswith("--"):
            raise ConfigValidationError(f"Invalid config override '{tok}': must start with --")
        opt = tok[2:]
        if "=" in opt:
            opt, raw = opt.split("=", 1)
            value: Any = _parse_value(raw)
        elif i < len(args) and not args[i].startswith("--"):
            value = _parse_value(args[i])
            i += 1
        else:
            value = True
        opt = opt.replace("-", "_") if "." not in opt else opt
        if "." not in opt:
            raise ConfigValidationError(
                f"Invalid config override '--{opt}': not a dotted section.key name"
            )
        result[opt] = value
    return result
