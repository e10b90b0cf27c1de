"""Trainable pipeline components: tok2vec, tagger, ner, parser.

These are the pipes the reference trains through spaCy's
``nlp.update`` (``/root/reference/spacy_ray/worker.py:176-189`` ->
``train_while_improving`` -> ``pipe.update``).  Factories are registered under
spaCy's factory names so ``[components.ner] factory = "ner"`` resolves.
"""
from __future__ import annotations

import json
from pathlib import Path
from typing import Any, Callable, Dict, Iterable, List, Optional, Sequence

import numpy as np
import torch

from ..config import registry
from ..models.transition_model import TransitionGold
from ..nn.batch import TokenBatch
from ..nn.layers import set_dropout_rate
from ..nn.model import Model
from . import scorer as S
from .doc import Doc, Example
from .transitions import (
    ArcEagerSystem, BiluoSystem, biluo_actions_to_spans, is_projective, spans_to_biluo_actions,
)


def _add_loss(losses, name: str, value) -> None:
    """Accumulate without forcing a device sync: tensors stay tensors until the
    logger formats them."""
    if losses is None:
        return
    prev = losses.get(name, 0.0)
    losses[name] = prev + (value.detach() if hasattr(value, "detach") else float(value))


class TrainablePipe:
    is_trainable = True
    default_score_weights: Dict[str, Any] = {}

    def __init__(self, name: str, model: Model, **cfg):
        self.name = name
        self.model = model
        self.cfg = dict(cfg)
        self._labels: List[str] = []

    # ---- labels ----------------------------------------------------------
    @property
    def labels(self) -> List[str]:
        return list(self._labels)

    def add_label(self, label: str) -> int:
        if label in self._labels:
            return 0
        self._labels.append(label)
        return 1

    # ---- lifecycle -------------------------------------------------------
    def initialize(self, get_examples: Callable[[], Iterable[Example]], *, nlp=None, labels=None) -> None:
        raise NotImplementedError

    def update(self, examples, *, batch: TokenBatch, drop: float = 0.0, sgd=None, losses=None):
        raise NotImplementedError

    def predict(self, docs: Sequence[Doc], batch: TokenBatch):
        raise NotImplementedError

    def set_annotations(self, docs: Sequence[Doc], preds) -> None:
        raise NotImplementedError

    def finish_update(self, sgd) -> None:
        self.model.finish_update(sgd)

    def score(self, examples) -> Dict[str, Any]:
        return {}

    @property
    def listening_to(self) -> List[Model]:
        return [n for n in self.model.walk() if n.name == "tok2vec_listener"]

    # ---- serialisation ---------------------------------------------------
    def to_disk(self, path: Path) -> None:
        path = Path(path)
        path.mkdir(parents=True, exist_ok=True)
        (path / "cfg").write_text(json.dumps({"labels": self._labels, **self._json_cfg()}, indent=1))
        (path / "model").write_bytes(self.model.to_bytes())

    def from_disk(self, path: Path) -> "TrainablePipe":
        path = Path(path)
        meta = json.loads((path / "cfg").read_text())
        self._labels = list(meta.get("labels", []))
        self._after_labels()
        self.model.initialize()
        self.model.from_bytes((path / "model").read_bytes())
        return self

    def _json_cfg(self) -> Dict[str, Any]:
        return {k: v for k, v in self.cfg.items() if isinstance(v, (int, float, str, bool, list, type(None)))}

    def _after_labels(self) -> None:
        pass


# ============================================================================
class Tok2VecComponent(TrainablePipe):
    """Shared embedding/encoding layer.  Downstream components that use a
    ``Tok2VecListener`` get this component's output and send gradients back."""

    def __init__(self, name: str, model: Model):
        super().__init__(name, model)
        self.listeners: List[Model] = []
        self._pending = None

    def add_listener(self, listener: Model) -> None:
        if listener not in self.listeners:
            self.listeners.append(listener)

    def find_listeners(self, component: TrainablePipe) -> None:
        for node in component.listening_to:
            up = node.attrs.get("upstream", "*")
            if up in ("*", self.name):
                self.add_listener(node)

    def initialize(self, get_examples, *, nlp=None, labels=None) -> None:
        self.model.initialize()

    def predict(self, docs, batch):
        out = self.model.predict(batch)
        for l in self.listeners:
            l.attrs["receive"](batch, out, None)
        return out

    def set_annotations(self, docs, preds) -> None:
        pass

    def update(self, examples, *, batch, drop=0.0, sgd=None, losses=None, defer_backprop: bool = False):
        """Run the forward pass now; the backward pass runs when the *last*
        listener has handed its gradient back (spaCy's listener protocol).

        ``defer_backprop=True`` (engine.Trainer, heads on concurrent streams): listeners only
        deposit their gradients; the caller sums them and runs the backward pass with
        ``finish_backprop()`` once every head's stream has been joined."""
        set_dropout_rate(self.model, drop)
        outputs, bp = self.model.begin_update(batch)
        n_listeners = len(self.listeners)
        state = {"d": None, "count": 0, "parts": []}

        def accumulate(dY):
            if defer_backprop:
                state["parts"].append(dY)
                return
            state["d"] = dY.clone() if state["d"] is None else state["d"].add_(dY)
            state["count"] += 1
            if state["count"] == n_listeners:
                bp(state["d"])
                if sgd not in (None, False):
                    self.finish_update(sgd)

        def finish():
            parts = state["parts"]
            if parts:
                d = parts[0].clone()
                for extra in parts[1:]:
                    d.add_(extra)
                bp(d)
            state["parts"] = []

        self._finish_backprop = finish
        for l in self.listeners:
            l.attrs["receive"](batch, outputs, accumulate)
        if losses is not None:
            losses.setdefault(self.name, 0.0)
        return losses

    def finish_backprop(self) -> None:
        """Sum the deposited listener gradients and run the deferred backward pass."""
        fn = getattr(self, "_finish_backprop", None)
        if fn is not None:
            fn()
            self._finish_backprop = None


# ============================================================================
class Tagger(TrainablePipe):
    default_score_weights = {"tag_acc": 1.0}

    def _after_labels(self) -> None:
        if self.model.has_dim("nO") is None and self._labels:
            self.model.set_dim("nO", len(self._labels))

    def initialize(self, get_examples, *, nlp=None, labels=None) -> None:
        if labels is not None:
            for l in labels:
                self.add_label(l)
        else:
            seen = set()
            for eg in get_examples():
                for t in (eg.reference.tags or []):
                    if t:
                        seen.add(t)
            for t in sorted(seen):
                self.add_label(t)
        if not self._labels:
            raise ValueError(f"[{self.name}] no tag labels found in the training data")
        self._after_labels()
        self.model.initialize()

    def gold_ids(self, ref: Doc) -> np.ndarray:
        """Per-token gold label ids of one reference doc (int64, -1 = no gold); cached on the doc.  This is
        the one thing the token-tagging components differ in - the generic update and the device-resident
        engine (``engine.ExampleStore``) both read it."""
        key = ("gold_ids", self.name)
        ids = ref.user_data.get(key)
        if ids is None:
            ids = np.asarray(self._compute_gold_ids(ref), dtype=np.int64)
            ref.user_data[key] = ids
        return ids

    def _compute_gold_ids(self, ref: Doc):
        index = {l: i for i, l in enumerate(self._labels)}
        tags = ref.tags or [None] * len(ref)
        return [index.get(t, -1) if t else -1 for t in tags]

    def _gold_labels(self, examples, batch: TokenBatch) -> torch.Tensor:
        arr = np.full((batch.n_rows,), -1, dtype=np.int64)
        for eg, s in zip(examples, batch.starts):
            ids = self.gold_ids(eg.reference)
            arr[s:s + len(ids)] = ids
        from ..nn.batch import to_device

        return to_device(arr, batch.device)

    def update(self, examples, *, batch, drop=0.0, sgd=None, losses=None):
        set_dropout_rate(self.model, drop)
        labels = self._gold_labels(examples, batch)
        loss, _guesses = self.model.attrs["update_with_labels"](batch, labels)
        if sgd not in (None, False):
            self.finish_update(sgd)
        _add_loss(losses, self.name, loss)
        return losses

    def predict(self, docs, batch):
        P = self.model.predict(batch)
        return P.argmax(dim=1)

    def set_annotations(self, docs, preds) -> None:
        host = preds.to("cpu").tolist()
        row = 1
        for doc in docs:
            n = len(doc)
            doc.tags = [self._labels[i] for i in host[row:row + n]]
            row += n + 1

    def score(self, examples):
        return S.score_tags(examples)


# ============================================================================
class SentenceRecognizer(Tagger):
    """``senter``: a two-label token tagger (``I`` / ``S``) whose ``S`` marks the first token of a sentence.
    Gold comes from ``Doc.sent_starts`` or, for treebank docs, from the roots of the dependency tree."""
    default_score_weights = {"sents_f": 1.0, "sents_p": 0.0, "sents_r": 0.0}

    def initialize(self, get_examples, *, nlp=None, labels=None) -> None:
        for l in ("I", "S"):
            self.add_label(l)
        self._after_labels()
        self.model.initialize()

    def _compute_gold_ids(self, ref: Doc):
        starts = ref.gold_sent_starts()
        if starts is None:
            return [-1] * len(ref)
        return [-1 if v is None else int(bool(v)) for v in starts]

    def set_annotations(self, docs, preds) -> None:
        host = preds.to("cpu").tolist()
        row = 1
        for doc in docs:
            n = len(doc)
            starts = [bool(v == 1) for v in host[row:row + n]]
            if starts:
                starts[0] = True
            doc.sent_starts = starts
            row += n + 1

    def score(self, examples):
        return S.score_sents(examples)


class Morphologizer(Tagger):
    """``morphologizer``: one label per (morphological features, UPOS) combination seen in training,
    written the way spaCy does (``"Case=Nom|Number=Sing|POS=NOUN"``); predicts ``Doc.pos`` and ``Doc.morphs``."""
    default_score_weights = {"pos_acc": 0.5, "morph_acc": 0.5}

    @staticmethod
    def _label_of(pos: Optional[str], morph: Optional[str]) -> Optional[str]:
        if not pos and not morph:
            return None
        feats = [f for f in (morph or "").split("|") if f and f != "_" and not f.startswith("POS=")]
        if pos:
            feats.append(f"POS={pos}")
        return "|".join(sorted(feats))

    @staticmethod
    def _split(label: str):
        feats = [f for f in label.split("|") if f]
        pos = next((f[4:] for f in feats if f.startswith("POS=")), None)
        return pos, "|".join(f for f in feats if not f.startswith("POS="))

    def _ref_labels(self, ref: Doc) -> List[Optional[str]]:
        n = len(ref)
        pos = ref.pos or [None] * n
        morphs = ref.morphs or [None] * n
        return [self._label_of(p, m) for p, m in zip(pos, morphs)]

    def initialize(self, get_examples, *, nlp=None, labels=None) -> None:
        if labels is not None:
            for l in labels:
                self.add_label(l)
        else:
            seen = set()
            for eg in get_examples():
                seen.update(l for l in self._ref_labels(eg.reference) if l)
            for l in sorted(seen):
                self.add_label(l)
        if not self._labels:
            raise ValueError(f"[{self.name}] no POS / morphological annotation found in the training data")
        self._after_labels()
        self.model.initialize()

    def _compute_gold_ids(self, ref: Doc):
        index = {l: i for i, l in enumerate(self._labels)}
        return [index.get(l, -1) if l else -1 for l in self._ref_labels(ref)]

    def set_annotations(self, docs, preds) -> None:
        host = preds.to("cpu").tolist()
        row = 1
        for doc in docs:
            n = len(doc)
            pairs = [self._split(self._labels[i]) for i in host[row:row + n]]
            doc.pos = [p for p, _ in pairs]
            doc.morphs = [m for _, m in pairs]
            row += n + 1

    def score(self, examples):
        out = S.score_token_attr(examples, "pos", "pos_acc")
        out.update(S.score_token_attr(examples, "morphs", "morph_acc"))
        return out


# ============================================================================
class TrainableLemmatizer(Tagger):
    """``trainable_lemmatizer``: lemmatisation as token classification over EDIT RULES learned from the
    training data (upstream's ``EditTreeLemmatizer`` classifies over edit trees with the same ``spacy.Tagger``
    model; the rule here is the suffix form of such a tree): a rule ``(case, k, suffix)`` lower-cases the form if
    ``case`` says so, strips its last ``k`` characters and appends ``suffix``.  Rules seen fewer than
    ``min_tree_freq`` times are dropped; tokens whose gold rule is unknown train nothing; with ``backoff =
    "orth"`` an unknown / inapplicable prediction falls back to the form itself."""
    default_score_weights = {"lemma_acc": 1.0}

    @staticmethod
    def rule_of(form: str, lemma: str) -> str:
        case = "L" if (lemma == lemma.lower() and form != form.lower()) else "K"
        base = form.lower() if case == "L" else form
        p = 0
        while p < len(base) and p < len(lemma) and base[p] == lemma[p]:
            p += 1
        return f"{case}|{len(base) - p}|{lemma[p:]}"

    @staticmethod
    def apply_rule(form: str, rule: str) -> Optional[str]:
        case, k, suffix = rule.split("|", 2)
        base = form.lower() if case == "L" else form
        k = int(k)
        if k > len(base):
            return None
        return base[: len(base) - k] + suffix

    def initialize(self, get_examples, *, nlp=None, labels=None) -> None:
        if labels is not None:
            for l in labels:
                self.add_label(l)
        else:
            from collections import Counter

            counts: Counter = Counter()
            for eg in get_examples():
                ref = eg.reference
                for w, lem in zip(ref.words, ref.lemmas or []):
                    if lem:
                        counts[self.rule_of(w, lem)] += 1
            min_freq = int(self.cfg.get("min_tree_freq", 3))
            for rule, n in sorted(counts.items(), key=lambda kv: (-kv[1], kv[0])):
                if n >= min_freq:
                    self.add_label(rule)
        if not self._labels:
            raise ValueError(f"[{self.name}] no lemma annotation (Doc.lemmas) found in the training data")
        self._after_labels()
        self.model.initialize()

    def _compute_gold_ids(self, ref: Doc):
        index = {l: i for i, l in enumerate(self._labels)}
        lemmas = ref.lemmas or [None] * len(ref)
        return [index.get(self.rule_of(w, lem), -1) if lem else -1 for w, lem in zip(ref.words, lemmas)]

    def set_annotations(self, docs, preds) -> None:
        host = preds.to("cpu").tolist()
        backoff = self.cfg.get("backoff", "orth")
        row = 1
        for doc in docs:
            n = len(doc)
            out = []
            for w, i in zip(doc.words, host[row:row + n]):
                lem = self.apply_rule(w, self._labels[i])
                out.append(lem if lem else (w if backoff == "orth" else None))
            doc.lemmas = out
            row += n + 1

    def score(self, examples):
        return S.score_token_attr(examples, "lemmas", "lemma_acc")


# ============================================================================
class EntityRecognizer(TrainablePipe):
    default_score_weights = {"ents_f": 1.0, "ents_p": 0.0, "ents_r": 0.0, "ents_per_type": None}

    def __init__(self, name, model, **cfg):
        super().__init__(name, model, **cfg)
        self.system: Optional[BiluoSystem] = None

    def _after_labels(self) -> None:
        self.system = BiluoSystem(self._labels)
        if self.model.has_dim("nO") is None:
            self.model.set_dim("nO", self.system.n_actions)

    def initialize(self, get_examples, *, nlp=None, labels=None) -> None:
        if labels is not None:
            for l in labels:
                self.add_label(l)
        else:
            seen = set()
            for eg in get_examples():
                for (_s, _e, lab) in eg.reference.ents:
                    seen.add(lab)
            for l in sorted(seen):
                self.add_label(l)
        if not self._labels:
            raise ValueError(f"[{self.name}] no entity labels found in the training data")
        self._after_labels()
        self.model.initialize()

    def gold_actions(self, ref: Doc) -> np.ndarray:
        key = ("ner_actions", self.name)
        acts = ref.user_data.get(key)
        if acts is None:
            if not ref.has_ents_annotation:
                acts = np.full((len(ref),), -1, dtype=np.int64)
            else:
                index = {l: i for i, l in enumerate(self._labels)}
                spans = [(s, e, index[l]) for (s, e, l) in ref.ents if l in index]
                acts = np.array(spans_to_biluo_actions(len(ref), spans), dtype=np.int64)
            ref.user_data[key] = acts
        return acts

    def _make_gold(self, examples, batch: TokenBatch) -> TransitionGold:
        per_doc = [self.gold_actions(eg.reference) for eg in examples]
        flat = np.concatenate(per_doc) if per_doc else np.zeros((0,), dtype=np.int64)
        offs = np.zeros(len(per_doc), dtype=np.int64)
        if len(per_doc) > 1:
            offs[1:] = np.cumsum([len(a) for a in per_doc])[:-1]
        from ..nn.batch import to_device

        dev = batch.device
        return TransitionGold(actions=to_device(flat, dev), offsets=to_device(offs, dev))

    def update(self, examples, *, batch, drop=0.0, sgd=None, losses=None):
        set_dropout_rate(self.model, drop)
        gold = self._make_gold(examples, batch)
        out = self.model.attrs["run"](batch, self.system, gold, True)
        if sgd not in (None, False):
            self.finish_update(sgd)
        _add_loss(losses, self.name, out.loss)
        return losses

    def predict(self, docs, batch):
        return self.model.attrs["run"](batch, self.system, None, False)

    def set_annotations(self, docs, out) -> None:
        host = out.actions_flat.to("cpu").tolist()
        pos = 0
        for doc in docs:
            n = len(doc)
            spans = biluo_actions_to_spans(host[pos:pos + n])
            doc.ents = [(s, e, self._labels[l]) for (s, e, l) in spans]
            doc.has_ents_annotation = True
            pos += n

    def score(self, examples):
        return S.score_ents(examples)


# ============================================================================
class DependencyParser(TrainablePipe):
    default_score_weights = {"dep_uas": 0.5, "dep_las": 0.5}

    def __init__(self, name, model, **cfg):
        super().__init__(name, model, **cfg)
        self.system: Optional[ArcEagerSystem] = None

    def _after_labels(self) -> None:
        self.system = ArcEagerSystem(self._labels)
        if self.model.has_dim("nO") is None:
            self.model.set_dim("nO", self.system.n_actions)

    def initialize(self, get_examples, *, nlp=None, labels=None) -> None:
        """Labels come from the PROJECTIVIZED gold trees (pseudo-projective parsing, ``models/nonproj.py``):
        a lifted arc's label is decorated ``dep||headdep``; decorated labels seen fewer than
        ``min_action_freq`` times (cfg, default 30 as upstream) fall back to the plain label."""
        if labels is not None:
            for l in labels:
                self.add_label(l)
        else:
            from collections import Counter

            from ..models import nonproj

            counts: Counter = Counter()
            for eg in get_examples():
                ref = eg.reference
                deps = ref.deps or []
                if ref.heads is not None and deps and nonproj.is_nonproj_tree(
                        [h if (h is not None and 0 <= h < len(ref)) else t for t, h in enumerate(ref.heads)]):
                    _ph, deps = nonproj.projectivize(
                        [h if (h is not None and 0 <= h < len(ref)) else t for t, h in enumerate(ref.heads)], list(deps))
                for d in deps:
                    if d:
                        counts[d] += 1
            min_freq = int(self.cfg.get("min_action_freq", 30))
            for l in sorted(counts):
                if nonproj.is_decorated(l) and counts[l] < min_freq:
                    continue
                self.add_label(l)
        if not self._labels:
            self.add_label("dep")
        self._after_labels()
        self.model.initialize()

    def _gold(self, ref: Doc):
        key = ("dep_gold", self.name)
        g = ref.user_data.get(key)
        if g is None:
            n = len(ref)
            if ref.heads is None:
                g = ([-1] * n, [-1] * n)
            else:
                from ..models import nonproj

                index = {l: i for i, l in enumerate(self._labels)}
                heads = [h if (h is not None and 0 <= h < n) else -1 for h in ref.heads]
                deps = list(ref.deps or [None] * n)
                full = [h if h >= 0 else t for t, h in enumerate(heads)]
                if not is_projective(full):
                    # pseudo-projective: lift the crossing arcs, decorate their labels (a decorated label
                    # that was pruned from the label set falls back to the plain one)
                    lifted, deco = nonproj.projectivize(full, deps)
                    heads = [lifted[t] if heads[t] >= 0 else -1 for t in range(n)]
                    deps = [d if (d in index or not nonproj.is_decorated(d)) else nonproj.decompose(d)[0] for d in deco]
                g = (heads, [index.get(d, -1) if d else -1 for d in deps])
            ref.user_data[key] = g
        return g

    def update(self, examples, *, batch, drop=0.0, sgd=None, losses=None):
        set_dropout_rate(self.model, drop)
        golds = [self._gold(eg.reference) for eg in examples]
        gold = TransitionGold(heads=[g[0] for g in golds], labels=[g[1] for g in golds])
        if batch.device.type == "cuda":        # flat device copies for the on-device oracle
            from ..nn.batch import to_device

            gold.heads_flat = to_device(np.concatenate([self._gold_np(eg.reference)[0] for eg in examples]), batch.device)
            gold.labels_flat = to_device(np.concatenate([self._gold_np(eg.reference)[1] for eg in examples]), batch.device)
        out = self.model.attrs["run"](batch, self.system, gold, True)
        if sgd not in (None, False):
            self.finish_update(sgd)
        _add_loss(losses, self.name, out.loss)
        return losses

    def predict(self, docs, batch):
        return self.model.attrs["run"](batch, self.system, None, False)

    def _gold_np(self, ref: Doc):
        key = ("dep_gold_np", self.name)
        g = ref.user_data.get(key)
        if g is None:
            heads, labels = self._gold(ref)
            # root = self; missing stays -1 (the list form marks roots as h == t already)
            g = (np.asarray(heads, dtype=np.int32), np.asarray(labels, dtype=np.int32))
            ref.user_data[key] = g
        return g

    def set_annotations(self, docs, out) -> None:
        if out.states is None and out.heads_flat is not None:     # derivations ran on the device
            heads_all = out.heads_flat.to("cpu").tolist()
            labs_all = out.labels_flat.to("cpu").tolist()
            pos = 0
            for doc in docs:
                n = len(doc)
                self._annotate(doc, heads_all[pos:pos + n], labs_all[pos:pos + n])
                pos += n
            return
        for doc, st in zip(docs, out.states):
            heads, labs = self.system.finalize(st)
            self._annotate(doc, heads, labs)

    def _annotate(self, doc: Doc, heads, labs) -> None:
        from ..models import nonproj

        deps = [self._labels[l] if l >= 0 else "ROOT" for l in labs]
        heads = list(heads)
        if any(nonproj.is_decorated(d) for d in deps):          # undo the pseudo-projective lifting
            heads, deps = nonproj.deprojectivize(heads, deps)
        doc.heads, doc.deps = heads, deps
        doc.user_data["sent_starts"] = nonproj.sentence_starts(heads)

    def score(self, examples):
        return S.score_deps(examples)


# ============================================================================
class TextCategorizer(TrainablePipe):
    """``textcat`` (mutually exclusive classes) / ``textcat_multilabel``: document categories in ``Doc.cats``.
    Labels come from the keys of the training docs' ``cats``; a label missing from a doc's ``cats`` is
    ignored for that doc (spaCy's ``not_missing`` mask).  Loss reported = mean squared error, as upstream."""
    default_score_weights = {"cats_score": 1.0, "cats_macro_f": 0.0, "cats_micro_f": 0.0, "cats_macro_p": None,
                             "cats_macro_r": None, "cats_micro_p": None, "cats_micro_r": None, "cats_f_per_type": None}
    multi_label = False

    def _after_labels(self) -> None:
        if self.model.has_dim("nO") is None and self._labels:
            self.model.set_dim("nO", len(self._labels))

    def initialize(self, get_examples, *, nlp=None, labels=None) -> None:
        if labels is not None:
            for l in labels:
                self.add_label(l)
        else:
            seen = set()
            for eg in get_examples():
                seen.update((eg.reference.cats or {}).keys())
            for l in sorted(seen):
                self.add_label(l)
        if not self._labels:
            raise ValueError(f"[{self.name}] no document categories (Doc.cats) found in the training data")
        if not self.multi_label and len(self._labels) < 2:
            raise ValueError(f"[{self.name}] mutually exclusive classes need at least two labels; "
                             "use textcat_multilabel for a single yes/no category")
        exclusive = self.model.attrs.get("exclusive_classes")
        if exclusive is not None and bool(exclusive) == self.multi_label:
            raise ValueError(f"[{self.name}] the model's exclusive_classes = {exclusive} does not fit the "
                             f"{'textcat_multilabel' if self.multi_label else 'textcat'} component")
        self._after_labels()
        self.model.initialize()

    def _truths(self, examples, device):
        n, k = len(examples), len(self._labels)
        truths = np.zeros((n, k), dtype=np.float32)
        known = np.zeros((n, k), dtype=np.float32)
        for i, eg in enumerate(examples):
            cats = eg.reference.cats or {}
            for j, l in enumerate(self._labels):
                if l in cats:
                    truths[i, j] = float(cats[l])
                    known[i, j] = 1.0
        return torch.from_numpy(truths).to(device), torch.from_numpy(known).to(device)

    def update(self, examples, *, batch, drop=0.0, sgd=None, losses=None):
        set_dropout_rate(self.model, drop)
        scores, backprop = self.model(batch, True)
        truths, known = self._truths(examples, scores.device)
        diff = (scores - truths) * known
        backprop(diff / max(len(examples), 1))
        if sgd not in (None, False):
            self.finish_update(sgd)
        _add_loss(losses, self.name, (diff * diff).mean())
        return losses

    def predict(self, docs, batch):
        return self.model.predict(batch)

    def set_annotations(self, docs, scores) -> None:
        host = scores.to("cpu").tolist()
        for doc, row in zip(docs, host):
            doc.cats = {l: float(v) for l, v in zip(self._labels, row)}

    def score(self, examples):
        return S.score_cats(examples, self._labels, multi_label=self.multi_label,
                            threshold=float(self.cfg.get("threshold", 0.5)))


class MultiLabelTextCategorizer(TextCategorizer):
    multi_label = True


# ============================================================================
class SpanCategorizer(TrainablePipe):
    """``spancat``: label candidate token spans (overlapping, multi-label) of ``Doc.spans[spans_key]``.
    Candidates come from the ``suggester`` (default: all 1-3-grams); a candidate's target is 1 for every gold
    label it carries.  Prediction keeps the labels whose score reaches ``threshold`` (at most ``max_positive``
    per span)."""

    def __init__(self, name, model, **cfg):
        super().__init__(name, model, **cfg)
        self.spans_key = str(cfg.get("spans_key", "sc"))
        sug = cfg.get("suggester")
        if sug is None:
            from ..models.spancat import ngram_suggester

            sug = ngram_suggester([1, 2, 3])
        self.suggester = sug
        k = f"spans_{self.spans_key}"
        self.default_score_weights = {f"{k}_f": 1.0, f"{k}_p": 0.0, f"{k}_r": 0.0}

    def _after_labels(self) -> None:
        if self.model.has_dim("nO") is None and self._labels:
            self.model.set_dim("nO", len(self._labels))

    def initialize(self, get_examples, *, nlp=None, labels=None) -> None:
        if labels is not None:
            for l in labels:
                self.add_label(l)
        else:
            seen = set()
            for eg in get_examples():
                seen.update(l for _s, _e, l in eg.reference.spans.get(self.spans_key, []))
            for l in sorted(seen):
                self.add_label(l)
        if not self._labels:
            raise ValueError(f"[{self.name}] no spans under Doc.spans[{self.spans_key!r}] in the training data")
        self._after_labels()
        self.model.initialize()

    def update(self, examples, *, batch, drop=0.0, sgd=None, losses=None):
        set_dropout_rate(self.model, drop)
        cands = self.suggester(batch.lengths)
        scores, backprop = self.model((batch, cands), True)
        index = {l: i for i, l in enumerate(self._labels)}
        target = np.zeros((len(cands), len(self._labels)), dtype=np.float32)
        where = {c: i for i, c in enumerate(cands)}
        for d, eg in enumerate(examples):
            for s_, e_, lab in eg.reference.spans.get(self.spans_key, []):
                i = where.get((d, s_, e_))
                if i is not None and lab in index:
                    target[i, index[lab]] = 1.0
        diff = scores - torch.from_numpy(target).to(scores.device)
        backprop(diff / max(len(examples), 1))
        if sgd not in (None, False):
            self.finish_update(sgd)
        _add_loss(losses, self.name, (diff * diff).sum() / max(len(examples), 1))
        return losses

    def predict(self, docs, batch):
        cands = self.suggester(batch.lengths)
        scores = self.model.predict((batch, cands))
        return cands, scores

    def set_annotations(self, docs, preds) -> None:
        cands, scores = preds
        host = scores.to("cpu").numpy() if len(cands) else np.zeros((0, len(self._labels)), dtype=np.float32)
        thr = float(self.cfg.get("threshold", 0.5))
        max_pos = self.cfg.get("max_positive")
        out = [[] for _ in docs]
        for (d, s_, e_), row in zip(cands, host):
            keep = [j for j in np.argsort(-row) if row[j] >= thr]
            if max_pos:
                keep = keep[: int(max_pos)]
            for j in keep:
                out[d].append((int(s_), int(e_), self._labels[j]))
        for doc, spans in zip(docs, out):
            doc.spans[self.spans_key] = sorted(spans)

    def score(self, examples):
        return S.score_spans(examples, self.spans_key)


# ============================================================================
class Sentencizer:
    """``sentencizer``: rule-based sentence boundaries (a token after sentence-final punctuation starts a
    sentence) - spaCy's non-trainable alternative to ``senter``.  Not trainable: ``nlp.update`` skips it."""
    is_trainable = False
    default_score_weights = {"sents_f": 1.0, "sents_p": 0.0, "sents_r": 0.0}
    default_punct_chars = ["!", ".", "?", "\u0589", "\u061f", "\u06d4", "\u0700", "\u0701", "\u0702", "\u0964", "\u3002",
                           "\uff01", "\uff0e", "\uff1f", "\uff61", "\u2026"]

    def __init__(self, name: str, punct_chars: Optional[Sequence[str]] = None, overwrite: bool = False, **cfg):
        self.name = name
        self.punct_chars = set(punct_chars) if punct_chars else set(self.default_punct_chars)
        self.overwrite = bool(overwrite)
        self.cfg = dict(cfg)

    @property
    def labels(self) -> List[str]:
        return []

    def predict(self, docs: Sequence[Doc], batch=None):
        out = []
        for doc in docs:
            starts = [False] * len(doc)
            if starts:
                starts[0] = True
            seen_period = False
            for i, w in enumerate(doc.words):
                is_punct = w in self.punct_chars
                if seen_period and not is_punct:
                    starts[i] = True
                    seen_period = False
                elif is_punct:
                    seen_period = True
            out.append(starts)
        return out

    def set_annotations(self, docs: Sequence[Doc], preds) -> None:
        for doc, starts in zip(docs, preds):
            if doc.sent_starts is None or self.overwrite:
                doc.sent_starts = list(starts)

    def score(self, examples):
        return S.score_sents(examples)

    def to_disk(self, path: Path) -> None:
        path = Path(path)
        path.mkdir(parents=True, exist_ok=True)
        (path / "cfg").write_text(json.dumps({"punct_chars": sorted(self.punct_chars), "overwrite": self.overwrite}))

    def from_disk(self, path: Path) -> "Sentencizer":
        meta = json.loads((Path(path) / "cfg").read_text())
        self.punct_chars = set(meta.get("punct_chars") or self.default_punct_chars)
        self.overwrite = bool(meta.get("overwrite", False))
        return self


# ---- factories ----------------------------------------------------------------
@registry.factories("tok2vec")
def make_tok2vec(nlp, name: str, model: Model) -> Tok2VecComponent:
    return Tok2VecComponent(name, model)


@registry.factories("tagger")
def make_tagger(nlp, name: str, model: Model, **cfg) -> Tagger:
    return Tagger(name, model, **cfg)


@registry.factories("senter")
def make_senter(nlp, name: str, model: Model, **cfg) -> SentenceRecognizer:
    return SentenceRecognizer(name, model, **cfg)


@registry.factories("trainable_lemmatizer")
def make_trainable_lemmatizer(nlp, name: str, model: Model, **cfg) -> TrainableLemmatizer:
    return TrainableLemmatizer(name, model, **cfg)


@registry.factories("morphologizer")
def make_morphologizer(nlp, name: str, model: Model, **cfg) -> Morphologizer:
    return Morphologizer(name, model, **cfg)


@registry.factories("sentencizer")
def make_sentencizer(nlp, name: str, punct_chars: Optional[Sequence[str]] = None, overwrite: bool = False, **cfg) -> Sentencizer:
    return Sentencizer(name, punct_chars=punct_chars, overwrite=overwrite, **cfg)


@registry.factories("spancat")
def make_spancat(nlp, name: str, model: Model, **cfg) -> SpanCategorizer:
    return SpanCategorizer(name, model, **cfg)


@registry.factories("textcat")
def make_textcat(nlp, name: str, model: Model, **cfg) -> TextCategorizer:
    return TextCategorizer(name, model, **cfg)


@registry.factories("textcat_multilabel")
def make_textcat_multilabel(nlp, name: str, model: Model, **cfg) -> MultiLabelTextCategorizer:
    return MultiLabelTextCategorizer(name, model, **cfg)


@registry.factories("ner")
def make_ner(nlp, name: str, model: Model, **cfg) -> EntityRecognizer:
    return EntityRecognizer(name, model, **cfg)


@registry.factories("parser")
def make_parser(nlp, name: str, model: Model, **cfg) -> DependencyParser:
    return DependencyParser(name, model, **cfg)


DEFAULT_MODEL_CONFIGS: Dict[str, Dict[str, Any]] = {
    "tok2vec": {
        "@architectures": "spacy.HashEmbedCNN.v2", "width": 96, "depth": 4, "embed_size": 2000,
        "window_size": 1, "maxout_pieces": 3, "subword_features": True, "pretrained_vectors": None,
    },
    "tagger": {
        "@architectures": "spacy.Tagger.v2",
        "tok2vec": {"@architectures": "spacy.HashEmbedCNN.v2", "width": 96, "depth": 4, "embed_size": 2000,
                    "window_size": 1, "maxout_pieces": 3, "subword_features": True, "pretrained_vectors": None},
    },
    "senter": {
        "@architectures": "spacy.Tagger.v2",
        "tok2vec": {"@architectures": "spacy.HashEmbedCNN.v2", "width": 12, "depth": 1, "embed_size": 2000,
                    "window_size": 1, "maxout_pieces": 2, "subword_features": True, "pretrained_vectors": None},
    },
    "trainable_lemmatizer": {
        "@architectures": "spacy.Tagger.v2",
        "tok2vec": {"@architectures": "spacy.HashEmbedCNN.v2", "width": 96, "depth": 4, "embed_size": 2000,
                    "window_size": 1, "maxout_pieces": 3, "subword_features": True, "pretrained_vectors": None},
    },
    "morphologizer": {
        "@architectures": "spacy.Tagger.v2",
        "tok2vec": {"@architectures": "spacy.HashEmbedCNN.v2", "width": 96, "depth": 4, "embed_size": 2000,
                    "window_size": 1, "maxout_pieces": 3, "subword_features": True, "pretrained_vectors": None},
    },
    "spancat": {
        "@architectures": "spacy.SpanCategorizer.v1", "hidden_size": 128,
        "tok2vec": {"@architectures": "spacy.HashEmbedCNN.v2", "width": 96, "depth": 4, "embed_size": 2000,
                    "window_size": 1, "maxout_pieces": 3, "subword_features": True, "pretrained_vectors": None},
    },
    "textcat": {
        "@architectures": "spacy.TextCatCNN.v2", "exclusive_classes": True,
        "tok2vec": {"@architectures": "spacy.HashEmbedCNN.v2", "width": 96, "depth": 4, "embed_size": 2000,
                    "window_size": 1, "maxout_pieces": 3, "subword_features": True, "pretrained_vectors": None},
    },
    "textcat_multilabel": {
        "@architectures": "spacy.TextCatCNN.v2", "exclusive_classes": False,
        "tok2vec": {"@architectures": "spacy.HashEmbedCNN.v2", "width": 96, "depth": 4, "embed_size": 2000,
                    "window_size": 1, "maxout_pieces": 3, "subword_features": True, "pretrained_vectors": None},
    },
    "ner": {
        "@architectures": "spacy.TransitionBasedParser.v2", "state_type": "ner", "extra_state_tokens": False,
        "hidden_width": 64, "maxout_pieces": 2, "use_upper": True,
        "tok2vec": {"@architectures": "spacy.HashEmbedCNN.v2", "width": 96, "depth": 4, "embed_size": 2000,
                    "window_size": 1, "maxout_pieces": 3, "subword_features": True, "pretrained_vectors": None},
    },
    "parser": {
        "@architectures": "spacy.TransitionBasedParser.v2", "state_type": "parser", "extra_state_tokens": False,
        "hidden_width": 128, "maxout_pieces": 3, "use_upper": True,
        "tok2vec": {"@architectures": "spacy.HashEmbedCNN.v2", "width": 96, "depth": 4, "embed_size": 2000,
                    "window_size": 1, "maxout_pieces": 3, "subword_features": True, "pretrained_vectors": None},
    },
}
