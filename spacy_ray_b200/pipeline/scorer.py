"""Scoring: the metrics spaCy's ``Scorer`` reports for the three heads."""
from __future__ import annotations

from collections import defaultdict
from typing import Any, Dict, Iterable, Sequence


class PRF:
    def __init__(self):
        self.tp = self.fp = self.fn = 0

    def score_set(self, cand: set, gold: set) -> None:
        self.tp += len(cand & gold)
        self.fp += len(cand - gold)
        self.fn += len(gold - cand)

    @property
    def precision(self) -> float:
        return self.tp / (self.tp + self.fp + 1e-100)

    @property
    def recall(self) -> float:
        return self.tp / (self.tp + self.fn + 1e-100)

    @property
    def fscore(self) -> float:
        p, r = self.precision, self.recall
        return 2 * p * r / (p + r + 1e-100)

    def to_dict(self) -> Dict[str, float]:
        return {"p": self.precision, "r": self.recall, "f": self.fscore}


def score_tags(examples: Iterable) -> Dict[str, Any]:
    right = total = 0
    for eg in examples:
        gold = eg.reference.tags
        pred = eg.predicted.tags
        if gold is None:
            continue
        for i, g in enumerate(gold):
            if g is None or g == "":
                continue
            total += 1
            if pred is not None and i < len(pred) and pred[i] == g:
                right += 1
    return {"tag_acc": (right / total) if total else None}


def score_ents(examples: Iterable) -> Dict[str, Any]:
    overall = PRF()
    per_type: Dict[str, PRF] = defaultdict(PRF)
    seen = False
    for eg in examples:
        if not eg.reference.has_ents_annotation:
            continue
        seen = True
        gold = {tuple(e) for e in eg.reference.ents}
        pred = {tuple(e) for e in eg.predicted.ents}
        overall.score_set(pred, gold)
        for lab in {e[2] for e in gold | pred}:
            per_type[lab].score_set({e for e in pred if e[2] == lab}, {e for e in gold if e[2] == lab})
    if not seen:
        return {"ents_p": None, "ents_r": None, "ents_f": None, "ents_per_type": None}
    return {
        "ents_p": overall.precision, "ents_r": overall.recall, "ents_f": overall.fscore,
        "ents_per_type": {k: v.to_dict() for k, v in per_type.items()},
    }


def score_deps(examples: Iterable) -> Dict[str, Any]:
    unl = PRF()
    lab = PRF()
    seen = False
    for eg in examples:
        gh, gd = eg.reference.heads, eg.reference.deps
        if gh is None:
            continue
        seen = True
        ph, pd = eg.predicted.heads, eg.predicted.deps
        gold_u = {(t, h) for t, h in enumerate(gh) if h is not None and h >= 0}
        gold_l = {(t, h, (gd[t] if gd else "")) for t, h in enumerate(gh) if h is not None and h >= 0}
        if ph is None:
            pred_u, pred_l = set(), set()
        else:
            pred_u = {(t, h) for t, h in enumerate(ph)}
            pred_l = {(t, h, (pd[t] if pd else "")) for t, h in enumerate(ph)}
        unl.score_set(pred_u, gold_u)
        lab.score_set(pred_l, gold_l)
    if not seen:
        return {"dep_uas": None, "dep_las": None}
    return {"dep_uas": unl.fscore, "dep_las": lab.fscore}


def score_token_attr(examples: Iterable, attr: str, key: str) -> Dict[str, Any]:
    """Token accuracy of a per-token string attribute of ``Doc`` (``pos``, ``morphs``, ``lemmas``)."""
    right = total = 0
    for eg in examples:
        gold = getattr(eg.reference, attr)
        pred = getattr(eg.predicted, attr)
        if gold is None:
            continue
        for i, g in enumerate(gold):
            if g is None:
                continue
            total += 1
            if pred is not None and i < len(pred) and (pred[i] or "") == (g or ""):
                right += 1
    return {key: (right / total) if total else None}


def score_sents(examples: Iterable) -> Dict[str, Any]:
    """Sentence spans (start, end) P/R/F, as spaCy's ``Scorer.score_spans(..., "sents")``."""
    prf = PRF()
    seen = False

    def spans(starts):
        idx = [i for i, v in enumerate(starts) if v] or [0]
        if idx[0] != 0:
            idx = [0] + idx
        return {(a, b) for a, b in zip(idx, idx[1:] + [len(starts)])}

    for eg in examples:
        gold = eg.reference.gold_sent_starts()
        if gold is None or any(v is None for v in gold):
            continue
        seen = True
        pred = eg.predicted.sent_starts
        prf.score_set(spans(pred) if pred is not None else set(), spans(gold))
    if not seen:
        return {"sents_p": None, "sents_r": None, "sents_f": None}
    return {"sents_p": prf.precision, "sents_r": prf.recall, "sents_f": prf.fscore}


def score_cats(examples: Iterable, labels: Sequence[str], *, multi_label: bool, threshold: float = 0.5) -> Dict[str, Any]:
    """Document categories: per-label P/R/F (argmax for exclusive classes, ``threshold`` otherwise), macro and
    micro averages; ``cats_score`` = macro F (spaCy's default for exclusive classes)."""
    per = {l: PRF() for l in labels}
    seen = False
    for eg in examples:
        gold = eg.reference.cats
        if not gold:
            continue
        seen = True
        pred = eg.predicted.cats or {}
        if multi_label:
            pos_pred = {l for l in labels if pred.get(l, 0.0) >= threshold}
        else:
            pos_pred = {max(labels, key=lambda l: pred.get(l, 0.0))} if pred else set()
        for l in labels:
            if l not in gold:
                continue
            g, p = gold[l] >= 0.5, l in pos_pred
            per[l].tp += int(g and p)
            per[l].fp += int(p and not g)
            per[l].fn += int(g and not p)
    if not seen:
        return {"cats_score": None, "cats_macro_f": None, "cats_micro_f": None, "cats_f_per_type": None}
    micro = PRF()
    for v in per.values():
        micro.tp += v.tp; micro.fp += v.fp; micro.fn += v.fn
    macro_f = sum(v.fscore for v in per.values()) / max(len(per), 1)
    return {"cats_score": macro_f, "cats_macro_f": macro_f,
            "cats_macro_p": sum(v.precision for v in per.values()) / max(len(per), 1),
            "cats_macro_r": sum(v.recall for v in per.values()) / max(len(per), 1),
            "cats_micro_p": micro.precision, "cats_micro_r": micro.recall, "cats_micro_f": micro.fscore,
            "cats_f_per_type": {k: v.to_dict() for k, v in per.items()}}


def score_spans(examples: Iterable, spans_key: str) -> Dict[str, Any]:
    """Labelled span P/R/F of one span group (spaCy's ``spans_{key}_p/r/f``)."""
    prf = PRF()
    seen = False
    for eg in examples:
        gold = eg.reference.spans.get(spans_key)
        if gold is None:
            continue
        seen = True
        prf.score_set({tuple(x) for x in eg.predicted.spans.get(spans_key, [])}, {tuple(x) for x in gold})
    k = f"spans_{spans_key}"
    if not seen:
        return {f"{k}_p": None, f"{k}_r": None, f"{k}_f": None}
    return {f"{k}_p": prf.precision, f"{k}_r": prf.recall, f"{k}_f": prf.fscore}


def weighted_score(scores: Dict[str, Any], weights: Dict[str, Any]) -> float:
    """Main score = sum_k w_k * scores[k] (missing / None scores count as 0),
    the same combination ``create_evaluation_callback`` uses upstream."""
    total = 0.0
    for k, w in (weights or {}).items():
        if w is None:
            continue
        v = scores.get(k)
        if isinstance(v, (int, float)):
            total += float(w) * float(v)
    return total


def combine_score_weights(weight_dicts: Sequence[Dict[str, Any]], overrides: Dict[str, Any] | None = None) -> Dict[str, Any]:
    """Merge per-component default weights so they sum to 1 (None = report only)."""
    overrides = overrides or {}
    result: Dict[str, Any] = {}
    active = [d for d in weight_dicts if d]
    for d in active:
        keyed = {k: v for k, v in d.items() if k not in overrides}
        tot = sum(v for v in keyed.values() if isinstance(v, (int, float)))
        for k, v in keyed.items():
            if isinstance(v, (int, float)) and tot > 0:
                result[k] = round(v / tot / len(active), 2)
            else:
                result[k] = v
    result.update(overrides)
    return result
