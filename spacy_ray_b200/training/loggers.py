"""Console loggers.

``spacy-ray.ConsoleLogger.v1`` reproduces the table of
``/root/reference/spacy_ray/loggers.py:8-66``: header ``T E # W | Loss <pipe>... |
<score cols>... | SCORE`` upper-cased, widths ``[8,3,6,6] + max(len,8) per loss
+ max(len,6) per score + [6]``, right aligned; rows: ``timedelta(seconds)``,
epoch, step, words, losses ``%.2f``, ``other_scores[col]*100`` ``%.2f``, score
``%.2f``.  (wasabi is not available, so row formatting is done here.)
``spacy.ConsoleLogger.v1`` is the stock-spaCy-like variant (E/#/losses/scores).
``spacy_ray_b200.JsonlLogger.v1`` adds a machine-readable sink.
"""
from __future__ import annotations

import json
import sys
from datetime import timedelta
from typing import Any, Callable, Dict, IO, List, Optional, Tuple

from ..config import registry


def format_row(cells: List[Any], widths: List[int], aligns: Optional[List[str]] = None, spacing: int = 3) -> str:
    out = []
    for i, (cell, width) in enumerate(zip(cells, widths)):
        text = str(cell)
        align = aligns[i] if aligns else "l"
        out.append(text.rjust(width) if align == "r" else text.ljust(width))
    return (" " * spacing).join(out)


def _missing_key_error(which: str, key: str, keys: List[str]) -> KeyError:
    return KeyError(
        f"[E983] Invalid key(s) for '{which}': {key}. Available keys: {keys}"
    )


@registry.loggers("spacy-ray.ConsoleLogger.v1")
def ray_console_logger(stream: Optional[IO] = None):
    def setup_printer(nlp, stdout: Optional[IO] = None, stderr: Optional[IO] = None) -> Tuple[Callable[[Dict[str, Any]], None], Callable[[], None]]:
        out = stream or stdout or sys.stdout
        score_cols = list(nlp.config["training"].get("score_weights") or {})
        score_widths = [max(len(col), 6) for col in score_cols]
        loss_cols = [f"Loss {pipe}" for pipe in nlp.pipe_names]
        loss_widths = [max(len(col), 8) for col in loss_cols]
        header = [c.upper() for c in ["T", "E", "#", "W"] + loss_cols + score_cols + ["Score"]]
        widths = [8, 3, 6, 6] + loss_widths + score_widths + [6]
        aligns = ["r"] * len(widths)
        print(format_row(header, widths), file=out)
        print(format_row(["-" * w for w in widths], widths), file=out)
        out.flush()

        def log_step(info: Dict[str, Any]) -> None:
            try:
                losses = ["{0:.2f}".format(float(info["losses"][p])) for p in nlp.pipe_names]
            except KeyError as e:
                raise _missing_key_error("scores (losses)", str(e), list(info["losses"].keys())) from None
            other = info.get("other_scores") or {}
            scores = []
            for col in score_cols:
                v = other.get(col, 0.0)
                scores.append("{0:.2f}".format(float(v if isinstance(v, (int, float)) else 0.0) * 100))
            data = (
                [str(timedelta(seconds=info["seconds"])), info["epoch"], info["step"], info["words"]]
                + losses + scores
                + ["{0:.2f}".format(float(info["score"] if info["score"] is not None else 0.0))]
            )
            print(format_row(data, widths, aligns), file=out)
            out.flush()

        def finalize() -> None:
            pass

        return log_step, finalize

    return setup_printer


@registry.loggers("spacy.ConsoleLogger.v1")
def console_logger(progress_bar: bool = False, stream: Optional[IO] = None):
    def setup_printer(nlp, stdout: Optional[IO] = None, stderr: Optional[IO] = None):
        out = stream or stdout or sys.stdout
        score_cols = [c for c, w in (nlp.config["training"].get("score_weights") or {}).items() if w is not None]
        loss_cols = [f"Loss {p}" for p in nlp.pipe_names]
        header = [c.upper() for c in ["E", "#"] + loss_cols + score_cols + ["Score"]]
        widths = [3, 6] + [max(len(c), 8) for c in loss_cols] + [max(len(c), 6) for c in score_cols] + [6]
        print(format_row(header, widths), file=out)
        print(format_row(["-" * w for w in widths], widths), file=out)

        def log_step(info: Optional[Dict[str, Any]]) -> None:
            if info is None or info.get("score") is None:
                return
            losses = ["{0:.2f}".format(float(info["losses"].get(p, 0.0))) for p in nlp.pipe_names]
            other = info.get("other_scores") or {}
            scores = ["{0:.2f}".format(float(other.get(c) or 0.0) * 100) for c in score_cols]
            data = [info["epoch"], info["step"]] + losses + scores + ["{0:.2f}".format(float(info["score"]))]
            print(format_row(data, widths, ["r"] * len(widths)), file=out)
            out.flush()

        return log_step, (lambda: None)

    return setup_printer


@registry.loggers("spacy_ray_b200.JsonlLogger.v1")
def jsonl_logger(path: str, console: bool = True):
    """One JSON object per evaluation row (plus device-timed throughput fields when
    the worker provides them: ``docs_per_sec``, ``words_per_sec``, ``grads_used``)."""
    def setup_printer(nlp, stdout: Optional[IO] = None, stderr: Optional[IO] = None):
        fh = open(path, "a", encoding="utf8")
        inner = ray_console_logger()(nlp, stdout, stderr) if console else (lambda info: None, lambda: None)

        def log_step(info: Dict[str, Any]) -> None:
            row = {k: v for k, v in info.items() if k not in ("checkpoints",)}
            fh.write(json.dumps(row, default=str) + "\n")
            fh.flush()
            inner[0](info)

        def finalize() -> None:
            fh.close()
            inner[1]()

        return log_step, finalize

    return setup_printer
