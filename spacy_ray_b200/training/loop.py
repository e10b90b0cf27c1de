"""Training loop helpers with the names/semantics the reference imports from
spaCy (``/root/reference/spacy_ray/worker.py:8-13``): ``train_while_improving``,
``create_train_batches``, ``create_evaluation_callback``,
``create_before_to_disk_callback``, ``update_meta``.  SURVEY.md appendix A
documents the contracts (the ``info`` dict, ``is_best_checkpoint`` being
``None`` on non-eval steps, patience / max_steps stopping)."""
from __future__ import annotations

import random
import time
from typing import Any, Callable, Dict, Iterable, Iterator, List, Optional, Sequence, Tuple

from pydantic import BaseModel, ConfigDict

from ..pipeline.doc import Example
from ..pipeline.scorer import weighted_score


class ConfigSchemaTraining(BaseModel):
    model_config = ConfigDict(extra="forbid", arbitrary_types_allowed=True)

    dev_corpus: str = "corpora.dev"
    train_corpus: str = "corpora.train"
    batcher: Any = None
    dropout: float = 0.1
    patience: int = 1600
    max_epochs: int = 0
    max_steps: int = 20000
    eval_frequency: int = 200
    seed: Optional[int] = 0
    gpu_allocator: Optional[str] = None
    accumulate_gradient: int = 1
    score_weights: Dict[str, Optional[float]] = {}
    optimizer: Any = None
    logger: Any = None
    frozen_components: List[str] = []
    annotating_components: List[str] = []
    before_to_disk: Any = None
    before_update: Any = None


def subdivide_batch(batch: Sequence[Any], accumulate_gradient: int) -> Iterator[List[Any]]:
    batch = list(batch)
    if accumulate_gradient <= 1:
        yield batch
        return
    batch.sort(key=len)
    sub = len(batch) // accumulate_gradient
    start = 0
    for i in range(accumulate_gradient):
        end = start + sub + (1 if i < len(batch) % accumulate_gradient else 0)
        if end > start:
            yield batch[start:end]
        start = end


def create_train_batches(
    nlp, corpus: Callable, batcher: Callable, max_epochs: int, *,
    rank: int = 0, world_size: int = 1, shard: bool = True, seed: int = 0,
) -> Iterator[Tuple[int, List[Example]]]:
    """Yields ``(epoch, batch)``.  ``max_epochs == 0`` = forever, ``-1`` = stream
    the corpus without loading/shuffling.  With ``shard`` each data-parallel rank
    sees a disjoint 1/world_size of every epoch's (identically shuffled) batches, and all
    ranks get the same number of batches (the ragged tail of an epoch is dropped) - the
    reference gives every rank the *same* data (SURVEY.md 2.3 "Data sharding: No (gap)")."""
    epoch = 0
    if max_epochs >= 0:
        examples = list(corpus(nlp))
        if not examples:
            raise ValueError("create_train_batches: the training corpus is empty")
    rng = random.Random(seed)
    while max_epochs < 1 or epoch != max_epochs:
        if max_epochs >= 0:
            rng.shuffle(examples)
            data: Iterable[Example] = examples
        else:
            data = corpus(nlp)
        if shard and world_size > 1:
            # Batch the (identically shuffled) GLOBAL stream and deal whole batches round-robin; a
            # round is only handed out once it is complete.  Every rank therefore runs exactly the
            # same number of steps per epoch - each step contains a collective, so a rank with one
            # batch more than its peers would wait for them forever (sync mode) - and still sees a
            # disjoint 1/world_size of the data at the configured batch size.
            group: List[List[Example]] = []
            for batch in batcher(data):
                group.append(batch)
                if len(group) == world_size:
                    yield epoch, group[rank]
                    group = []
        else:
            for batch in batcher(data):
                yield epoch, batch
        epoch += 1


def create_evaluation_callback(nlp, dev_corpus: Callable, weights: Dict[str, Optional[float]]) -> Callable[[], Tuple[float, Dict[str, Any]]]:
    weights = {k: v for k, v in (weights or {}).items() if v is not None}

    def evaluate() -> Tuple[float, Dict[str, Any]]:
        examples = list(dev_corpus(nlp))
        if not examples:
            return 0.0, {}
        scores = nlp.evaluate(examples)
        return weighted_score(scores, weights), scores

    return evaluate


def create_before_to_disk_callback(callback: Optional[Callable]) -> Callable:
    def before_to_disk(nlp):
        if not callback:
            return nlp
        modified = callback(nlp)
        if modified is None or not hasattr(modified, "to_disk"):
            raise ValueError("before_to_disk callback must return the nlp object")
        return modified

    return before_to_disk


def update_meta(training: Dict[str, Any], nlp, info: Dict[str, Any]) -> None:
    nlp.meta["performance"] = {}
    for metric in (training.get("score_weights") or {}):
        if metric is not None:
            nlp.meta["performance"][metric] = (info.get("other_scores") or {}).get(metric, 0.0)
    for pipe_name in nlp.pipe_names:
        if pipe_name in info.get("losses", {}):
            nlp.meta["performance"][f"{pipe_name}_loss"] = info["losses"][pipe_name]


def train_while_improving(
    nlp,
    optimizer,
    train_data: Iterable[Tuple[int, List[Example]]],
    evaluate: Callable[[], Tuple[float, Dict[str, Any]]],
    *,
    dropout: Any,
    eval_frequency: int,
    accumulate_gradient: int,
    patience: int,
    max_steps: int,
    exclude: Sequence[str] = (),
    annotating_components: Sequence[str] = (),
    before_update: Optional[Callable] = None,
    after_step: Optional[Callable[[int], None]] = None,
):
    """Generator of ``(batch, info, is_best_checkpoint)``, one item per step.

    ``after_step(step)`` is an extension point the distributed worker uses to run
    the synchronous gradient exchange + sharded optimizer step after the
    backward pass (the reference instead hides the optimizer inside the proxy's
    ``get_param``)."""
    if isinstance(dropout, (int, float)):
        import itertools

        dropouts = itertools.repeat(float(dropout))
    else:
        dropouts = iter(dropout)
    results: List[Tuple[float, int]] = []
    losses: Dict[str, float] = {}
    words_seen = 0
    start_time = time.time()
    best_step = 0
    for step, (epoch, batch) in enumerate(train_data):
        if before_update:
            before_update(nlp, {"step": step, "epoch": epoch})
        drop = next(dropouts)
        for subbatch in subdivide_batch(batch, accumulate_gradient):
            nlp.update(subbatch, drop=drop, losses=losses, sgd=False, exclude=exclude,
                       annotates=annotating_components)
        for name, proc in nlp.pipeline:
            if name not in exclude and getattr(proc, "is_trainable", False) and hasattr(proc, "finish_update"):
                proc.finish_update(optimizer)
        if after_step is not None:
            after_step(step)
        optimizer.step_schedules()
        if not (step % eval_frequency):
            averages = getattr(optimizer, "averages", None)
            score, other_scores = evaluate()
            results.append((score, step))
            is_best_checkpoint = score == max(results)[0]
            if is_best_checkpoint:
                best_step = step
        else:
            score, other_scores = None, None
            is_best_checkpoint = None
        words_seen += sum(len(eg) for eg in batch)
        info = {
            "epoch": epoch,
            "step": step,
            "score": score,
            "other_scores": other_scores,
            "losses": losses,
            "checkpoints": results,
            "seconds": int(time.time() - start_time),
            "words": words_seen,
        }
        yield batch, info, is_best_checkpoint
        if is_best_checkpoint is not None:
            losses = {}
        if patience and (step - best_step) >= patience:
            break
        if max_steps and step + 1 >= max_steps:
            break
