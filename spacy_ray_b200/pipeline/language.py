"""``Language``: the pipeline container (spaCy ``nlp`` equivalent).

The reference never touches pipeline internals beyond
``nlp.pipeline`` / ``component.model`` (``worker.py:226-251``),
``nlp.config`` (``worker.py:92``), ``nlp.select_pipes`` / ``to_disk``
(``worker.py:219-222``) and what ``train_while_improving`` calls
(``nlp.update``).  Those are the contracts implemented here.
"""
from __future__ import annotations

import contextlib
import json
import time
from pathlib import Path
from typing import Any, Callable, Dict, Iterable, Iterator, List, Optional, Sequence, Tuple, Union

import torch

from .. import about
from ..config import Config, ConfigValidationError, registry
from ..config.registry import _resolve_node
from ..nn.batch import TokenBatch, make_token_batch
from ..ops import get_current_ops
from . import components as C
from .doc import Doc, Example, _tokenize_with_offsets
from .scorer import combine_score_weights

DEFAULT_CONFIG_STR = """
[paths]
train = null
dev = null
vectors = null
init_tok2vec = null

[system]
seed = 0
gpu_allocator = null

[nlp]
lang = "en"
pipeline = []
disabled = []
before_creation = null
after_creation = null
after_pipeline_creation = null
batch_size = 1000

[components]

[corpora]

[corpora.train]
@readers = "spacy.Corpus.v1"
path = ${paths.train}
gold_preproc = false
max_length = 0
limit = 0
augmenter = null

[corpora.dev]
@readers = "spacy.Corpus.v1"
path = ${paths.dev}
gold_preproc = false
max_length = 0
limit = 0
augmenter = null

[training]
train_corpus = "corpora.train"
dev_corpus = "corpora.dev"
seed = ${system.seed}
gpu_allocator = ${system.gpu_allocator}
dropout = 0.1
accumulate_gradient = 1
patience = 1600
max_epochs = 0
max_steps = 20000
eval_frequency = 200
frozen_components = []
annotating_components = []
before_to_disk = null
before_update = null

[training.batcher]
@batchers = "spacy.batch_by_words.v1"
discard_oversize = false
tolerance = 0.2
get_length = null

[training.batcher.size]
@schedules = "compounding.v1"
start = 100
stop = 1000
compound = 1.001
t = 0.0

[training.logger]
@loggers = "spacy.ConsoleLogger.v1"
progress_bar = false

[training.optimizer]
@optimizers = "Adam.v1"
beta1 = 0.9
beta2 = 0.999
L2_is_weight_decay = true
L2 = 0.01
grad_clip = 1.0
use_averages = false
eps = 0.00000001
learn_rate = 0.001

[training.score_weights]

[initialize]
vectors = ${paths.vectors}
init_tok2vec = ${paths.init_tok2vec}
vocab_data = null
lookups = null
before_init = null
after_init = null

[initialize.components]

[initialize.tokenizer]
"""


def default_config() -> Config:
    return Config().from_str(DEFAULT_CONFIG_STR, interpolate=False)


class Language:
    def __init__(self, lang: str = "en", *, config: Optional[Config] = None, meta: Optional[Dict] = None):
        self.lang = lang
        self._config: Config = config.copy() if config is not None else default_config()
        self._config.setdefault("nlp", {})["lang"] = lang
        self._components: List[Tuple[str, Any]] = []
        self._disabled: set = set()
        self._meta: Dict[str, Any] = dict(meta or {})
        self._optimizer = None

    # ---- construction ----------------------------------------------------
    @classmethod
    def from_config(cls, config: Union[Config, Dict], *, auto_fill: bool = True) -> "Language":
        cfg = Config(config)
        if auto_fill:
            cfg = default_config().merge(cfg)
        lang = cfg.get("nlp", {}).get("lang", "en")
        nlp = cls(lang, config=cfg)
        interp = cfg.interpolate() if not cfg.is_interpolated or True else cfg
        pipeline = list(interp["nlp"].get("pipeline", []))
        comp_cfgs = interp.get("components", {})
        for name in pipeline:
            if name not in comp_cfgs:
                raise ConfigValidationError(f"Pipeline component '{name}' has no [components.{name}] block")
            block = dict(comp_cfgs[name])
            factory = block.pop("factory", None)
            if block.pop("source", None) is not None:
                raise NotImplementedError("sourcing components from other pipelines is not supported")
            if factory is None:
                raise ConfigValidationError(f"[components.{name}] needs a 'factory' key")
            nlp.add_pipe(factory, name=name, config=block, _write_config=False)
        nlp._disabled = set(interp["nlp"].get("disabled", []) or [])
        return nlp

    def add_pipe(self, factory_name: str, name: Optional[str] = None, *, config: Optional[Dict] = None,
                 _write_config: bool = True):
        name = name or factory_name
        if name in self.component_names:
            raise ValueError(f"'{name}' already exists in pipeline")
        block = dict(config or {})
        if "model" not in block and factory_name in C.DEFAULT_MODEL_CONFIGS:
            block["model"] = json.loads(json.dumps(C.DEFAULT_MODEL_CONFIGS[factory_name]))
        factory = registry.factories.get(factory_name)
        resolved = _resolve_node(block, f"components.{name}")
        component = factory(self, name, **resolved)
        self._components.append((name, component))
        if _write_config:
            self._config.setdefault("components", {})[name] = {"factory": factory_name, **block}
            pl = self._config.setdefault("nlp", {}).setdefault("pipeline", [])
            if name not in pl:
                pl.append(name)
        self._link_listeners()
        return component

    def _link_listeners(self) -> None:
        t2vs = [c for _, c in self._components if isinstance(c, C.Tok2VecComponent)]
        for t2v in t2vs:
            for _, comp in self._components:
                if comp is not t2v and hasattr(comp, "listening_to"):
                    t2v.find_listeners(comp)

    # ---- introspection ---------------------------------------------------
    @property
    def config(self) -> Config:
        self._config.setdefault("nlp", {})["pipeline"] = self.component_names
        return self._config

    @config.setter
    def config(self, value: Config) -> None:
        self._config = Config(value)

    @property
    def meta(self) -> Dict[str, Any]:
        m = self._meta
        m.setdefault("lang", self.lang)
        m.setdefault("name", "pipeline")
        m.setdefault("version", "0.0.0")
        m.setdefault("spacy_ray_b200_version", about.__version__)
        m["pipeline"] = self.pipe_names
        m["components"] = self.component_names
        m["labels"] = {n: list(getattr(c, "labels", [])) for n, c in self._components}
        return m

    @property
    def components(self) -> List[Tuple[str, Any]]:
        return list(self._components)

    @property
    def component_names(self) -> List[str]:
        return [n for n, _ in self._components]

    @property
    def pipeline(self) -> List[Tuple[str, Any]]:
        return [(n, c) for n, c in self._components if n not in self._disabled]

    @property
    def pipe_names(self) -> List[str]:
        return [n for n, _ in self.pipeline]

    def get_pipe(self, name: str):
        for n, c in self._components:
            if n == name:
                return c
        raise KeyError(f"No component '{name}' in pipeline {self.component_names}")

    def has_pipe(self, name: str) -> bool:
        return name in self.component_names

    @contextlib.contextmanager
    def select_pipes(self, *, disable: Optional[Sequence[str]] = None, enable: Optional[Sequence[str]] = None):
        prev = set(self._disabled)
        if enable is not None:
            disable = [n for n in self.component_names if n not in enable]
        self._disabled |= set(disable or [])
        try:
            yield self
        finally:
            self._disabled = prev

    # ---- data -> device ---------------------------------------------------
    def make_doc(self, text: str) -> Doc:
        words, _ = _tokenize_with_offsets(text)
        return Doc(words)

    def make_batch(self, docs: Sequence[Doc], *, capacity_rows: Optional[int] = None) -> TokenBatch:
        ops = get_current_ops()
        batch = make_token_batch([d.to_array() for d in docs], ops.device, capacity_rows=capacity_rows)
        vectors = self.vectors
        if vectors is not None:
            # static vectors: each token's row in the table (-1 = out of vocabulary), padded-ragged layout
            import numpy as np

            from ..nn.batch import to_device

            rows = np.full((batch.n_rows,), -1, dtype=np.int64)
            for doc, s in zip(docs, batch.starts):
                r = doc.user_data.get("vec_rows")
                if r is None or doc.user_data.get("vec_table") is not vectors:
                    r = vectors.rows_for(doc.words)
                    doc.user_data["vec_rows"], doc.user_data["vec_table"] = r, vectors
                rows[s:s + len(r)] = r
            batch.extra["vec_rows"] = to_device(rows, ops.device)
        return batch

    # ---- static vectors -------------------------------------------------------
    @property
    def vectors(self):
        from ..nn.staticvectors import get_vectors

        return get_vectors()

    def load_vectors(self, path) -> None:
        """``[initialize] vectors = <path>``: .npz (``keys`` | ``words`` + ``data``) or word2vec text."""
        from ..nn.staticvectors import Vectors, set_vectors

        set_vectors(Vectors.from_disk(path))

    # ---- training -----------------------------------------------------------
    def initialize(self, get_examples: Optional[Callable[[], Iterable[Example]]] = None, *, sgd=None):
        if get_examples is None:
            get_examples = lambda: []  # noqa: E731
        cache: List[Example] = []

        def examples():
            if not cache:
                cache.extend(get_examples())
            return cache

        init_all = self._config.interpolate().get("initialize", {}) or {}
        if init_all.get("vectors"):
            self.load_vectors(init_all["vectors"])
        init_cfg = init_all.get("components", {}) or {}
        for name, comp in self._components:
            if hasattr(comp, "initialize"):
                kwargs = dict(init_cfg.get(name, {}) or {})
                comp.initialize(examples, nlp=self, **kwargs)
        self._link_listeners()
        self._fill_score_weights()
        if sgd is not None:
            self._optimizer = sgd
        return self._optimizer

    def _fill_score_weights(self) -> None:
        tr = self._config.setdefault("training", {})
        current = dict(tr.get("score_weights") or {})
        defaults = [getattr(c, "default_score_weights", {}) for _, c in self.pipeline]
        tr["score_weights"] = combine_score_weights(defaults, current)

    def create_optimizer(self):
        cfg = self._config.interpolate()
        return _resolve_node(cfg["training"]["optimizer"], "training.optimizer")

    def update(
        self,
        examples: Sequence[Example],
        _: Any = None,
        *,
        drop: float = 0.0,
        sgd: Any = None,
        losses: Optional[Dict[str, float]] = None,
        exclude: Sequence[str] = (),
        annotates: Sequence[str] = (),
    ) -> Dict[str, float]:
        """One forward/backward over ``examples`` for every trainable pipe.
        ``sgd=False`` means "accumulate gradients only" (what the training loop
        passes), ``sgd=None`` creates/uses the default optimizer."""
        if losses is None:
            losses = {}
        if not examples:
            return losses
        if sgd is None:
            if self._optimizer is None:
                self._optimizer = self.create_optimizer()
            sgd = self._optimizer
        trainer = getattr(self, "_trainer", None)
        if trainer is not None and sgd is False and not annotates and not exclude:
            # device-resident fast path (engine.Trainer): native collate -> one H2D copy ->
            # CUDA-graph replay of forward+backward+gradient exchange+optimizer
            loss = trainer.update_examples(examples)
            if loss is not None:
                for head, value in trainer.losses_dict(loss).items():
                    C._add_loss(losses, head, value)
                self._trainer_stepped = bool(getattr(trainer, "exchange", True))    # False: caller runs proxy.step()
                return losses
        batch = self.make_batch([eg.predicted for eg in examples])
        for name, comp in self.pipeline:
            if name in exclude or not getattr(comp, "is_trainable", False):
                continue
            comp.update(examples, batch=batch, drop=drop, sgd=False, losses=losses)
            if name in annotates:
                preds = comp.predict([eg.predicted for eg in examples], batch)
                comp.set_annotations([eg.predicted for eg in examples], preds)
        if sgd not in (None, False):
            for name, comp in self.pipeline:
                if name not in exclude and getattr(comp, "is_trainable", False):
                    comp.finish_update(sgd)
        return losses

    # ---- inference / evaluation --------------------------------------------
    def pipe(self, docs: Iterable[Doc], *, batch_size: int = 256) -> Iterator[Doc]:
        buf: List[Doc] = []
        for doc in docs:
            if isinstance(doc, str):
                doc = self.make_doc(doc)
            buf.append(doc)
            if len(buf) >= batch_size:
                yield from self._annotate(buf)
                buf = []
        if buf:
            yield from self._annotate(buf)

    def _annotate(self, docs: List[Doc]) -> List[Doc]:
        docs = [d for d in docs]
        nonempty = [d for d in docs if len(d) > 0]
        if nonempty:
            batch = self.make_batch(nonempty)
            for _name, comp in self.pipeline:
                preds = comp.predict(nonempty, batch)
                comp.set_annotations(nonempty, preds)
        return docs

    def __call__(self, text: Union[str, Doc]) -> Doc:
        doc = self.make_doc(text) if isinstance(text, str) else text
        return self._annotate([doc])[0]

    def evaluate(self, examples: Sequence[Example], *, batch_size: int = 256) -> Dict[str, Any]:
        examples = list(examples)
        for eg in examples:
            eg.predicted = eg.reference.copy_unannotated()
        t0 = time.perf_counter()
        n_words = 0
        for i in range(0, len(examples), batch_size):
            chunk = examples[i:i + batch_size]
            self._annotate([eg.predicted for eg in chunk])
            n_words += sum(len(eg) for eg in chunk)
        get_current_ops().synchronize()
        dt = max(time.perf_counter() - t0, 1e-9)
        scores: Dict[str, Any] = {"speed": n_words / dt}
        for _name, comp in self.pipeline:
            if hasattr(comp, "score"):
                scores.update(comp.score(examples))
        return scores

    # ---- serialisation -----------------------------------------------------
    def to_disk(self, path: Union[str, Path], *, exclude: Sequence[str] = ()) -> None:
        path = Path(path)
        path.mkdir(parents=True, exist_ok=True)
        self.config.to_disk(path / "config.cfg")
        (path / "meta.json").write_text(json.dumps(self.meta, indent=2, default=str))
        (path / "tokenizer").write_text(json.dumps({"type": "regex_whitespace_punct.v1"}))
        (path / "vocab").mkdir(exist_ok=True)
        (path / "vocab" / "strings.json").write_text(json.dumps(
            sorted({l for _, c in self._components for l in getattr(c, "labels", [])})
        ))
        if self.vectors is not None and self._uses_static_vectors():
            self.vectors.to_disk(path / "vocab" / "vectors.npz")
        for name, comp in self._components:
            if name in exclude or not hasattr(comp, "to_disk"):
                continue
            comp.to_disk(path / name)

    def _uses_static_vectors(self) -> bool:
        return any(node.name == "staticvectors" for _n, c in self._components if hasattr(c, "model")
                   for node in c.model.walk())

    def from_disk(self, path: Union[str, Path], *, exclude: Sequence[str] = ()) -> "Language":
        path = Path(path)
        if (path / "meta.json").exists():
            self._meta = json.loads((path / "meta.json").read_text())
        if (path / "vocab" / "vectors.npz").exists():
            self.load_vectors(path / "vocab" / "vectors.npz")
        for name, comp in self._components:
            if name in exclude or not hasattr(comp, "from_disk"):
                continue
            comp.from_disk(path / name)
        self._link_listeners()
        return self

    def resume_training(self, *, sgd=None):
        self._optimizer = sgd or self.create_optimizer()
        return self._optimizer


def blank(lang: str = "en", *, config: Optional[Dict] = None) -> Language:
    cfg = default_config()
    if config:
        cfg = cfg.merge(config)
    cfg["nlp"]["lang"] = lang
    return Language.from_config(cfg)


def load(path: Union[str, Path]) -> Language:
    """Load a saved pipeline directory (``config.cfg`` + per-pipe blobs)."""
    path = Path(path)
    cfg = Config().from_disk(path / "config.cfg", interpolate=False)
    from ..nn.model import reset_model_ids

    reset_model_ids()
    nlp = Language.from_config(cfg)
    nlp.from_disk(path)
    return nlp
