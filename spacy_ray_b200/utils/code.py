"""``--code`` support: import a user Python file so its registered functions
resolve (the reference does this in every worker, ``worker.py:87``)."""
from __future__ import annotations

import importlib.util
import sys
from pathlib import Path
from typing import Optional, Union


def import_code(code_path: Optional[Union[str, Path]]) -> None:
    if code_path is None:
        return
    path = Path(code_path)
    if not path.exists():
        raise FileNotFoundError(f"Path to Python code not found: {path}")
    name = "srb_user_code_" + path.stem
    if name in sys.modules:
        return
    spec = importlib.util.spec_from_file_location(name, str(path))
    if spec is None or spec.loader is None:
        raise ImportError(f"Couldn't load Python code: {path}")
    module = importlib.util.module_from_spec(spec)
    sys.modules[name] = module
    spec.loader.exec_module(module)
