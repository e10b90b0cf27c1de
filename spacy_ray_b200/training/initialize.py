"""``init_nlp``: config -> initialised pipeline (spaCy ``training.initialize``
equivalent; the reference calls it in every worker, ``worker.py:91``)."""
from __future__ import annotations


from ..config import Config, registry, resolve_dot_names
from ..nn.layers import fix_random_seed
from ..nn.model import reset_model_ids
from ..ops import get_current_ops, require_cpu, require_gpu
from ..pipeline.language import Language
from .loop import ConfigSchemaTraining


def init_nlp(config: Config, *, use_gpu: int = -1, reset_ids: bool = True) -> Language:
    """Seed, build the pipeline from ``config``, discover labels from the train
    corpus, allocate parameters.  Deterministic given the config: node ids are
    restarted and the init RNG reseeded, so every rank (and a restarted run)
    gets identical keys and identical initial weights."""
    raw = Config(config)
    interp = raw.interpolate()
    seed = (interp.get("training", {}) or {}).get("seed", (interp.get("system", {}) or {}).get("seed", 0)) or 0
    fix_random_seed(int(seed))
    if reset_ids:
        reset_model_ids()
    if use_gpu is not None and use_gpu >= 0:
        ops = get_current_ops()
        if ops.device.type != "cuda":
            require_gpu(use_gpu)
    nlp = Language.from_config(raw)
    filled = nlp.config.interpolate()
    T = registry.resolve(filled["training"], schema=ConfigSchemaTraining)
    train_corpus, _dev = resolve_dot_names(filled, [T["train_corpus"], T["dev_corpus"]])
    frozen = set(T["frozen_components"])
    with nlp.select_pipes(disable=[n for n in nlp.component_names if n in frozen]):
        nlp.initialize(lambda: train_corpus(nlp))
    return nlp
