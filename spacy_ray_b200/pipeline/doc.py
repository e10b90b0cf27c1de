"""``Doc`` / ``Example`` and lexical attribute hashing.

spaCy's ``Doc`` is a Cython object over a vocab; the training path only needs
(a) the token strings, (b) their lexical attribute ids (NORM / PREFIX / SUFFIX /
SHAPE - the ``MultiHashEmbed`` inputs) as a ``uint64`` array, and (c) gold
annotations.  Attribute ids are 64-bit string hashes (FNV-1a + fmix64; spaCy uses
MurmurHash64A - the tables only need *a* well-mixed id, not that one).

A native implementation of the same featurisation lives in
``spacy_ray_b200/native`` (C++); ``featurize_words`` uses it when built and
falls back to the pure-Python path below (bit-identical; tested).
"""
from __future__ import annotations

from dataclasses import dataclass
from functools import lru_cache
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

_M64 = (1 << 64) - 1
_FNV_OFFSET = 0xCBF29CE484222325
_FNV_PRIME = 0x100000001B3


def _fmix64(h: int) -> int:
    h ^= h >> 33
    h = (h * 0xFF51AFD7ED558CCD) & _M64
    h ^= h >> 33
    h = (h * 0xC4CEB9FE1A85EC53) & _M64
    h ^= h >> 33
    return h


@lru_cache(maxsize=1 << 20)
def hash_string(s: str) -> int:
    """64-bit id of a string: FNV-1a over the UTF-8 bytes, then fmix64.  Never 0
    (0 is reserved for "pad row" in the batch layout)."""
    h = _FNV_OFFSET
    for byte in s.encode("utf8"):
        h = ((h ^ byte) * _FNV_PRIME) & _M64
    h = _fmix64(h)
    return h or 1


def word_shape(text: str) -> str:
    """spaCy's shape feature: letters -> X/x, digits -> d, other chars kept;
    runs longer than 4 of the same class are truncated."""
    if len(text) >= 100:
        return "LONG"
    out = []
    last = ""
    run = 0
    for ch in text:
        if ch.isalpha():
            c = "X" if ch.isupper() else "x"
        elif ch.isdigit():
            c = "d"
        else:
            c = ch
        if c == last:
            run += 1
        else:
            run = 0
            last = c
        if run < 4:
            out.append(c)
    return "".join(out)


@lru_cache(maxsize=1 << 20)
def lex_attrs(word: str) -> Tuple[int, int, int, int, int]:
    """(NORM, PREFIX, SUFFIX, SHAPE, ORTH) ids of one token."""
    return (
        hash_string(word.lower()),
        hash_string(word[:1]),
        hash_string(word[-3:]),
        hash_string(word_shape(word)),
        hash_string(word),
    )


def featurize_words_py(words: Sequence[str]) -> np.ndarray:
    arr = np.empty((len(words), 4), dtype=np.uint64)
    for i, w in enumerate(words):
        a = lex_attrs(w)
        arr[i, 0], arr[i, 1], arr[i, 2], arr[i, 3] = a[0], a[1], a[2], a[3]
    return arr


def featurize_words(words: Sequence[str]) -> np.ndarray:
    from ..native import featurize as _native

    if _native.available():
        return _native.featurize_words(words)
    return featurize_words_py(words)


class Doc:
    """A tokenised text with optional gold annotations.

    ``tags``: per-token fine-grained tags; ``pos`` / ``morphs`` / ``lemmas``: per-token UPOS, FEATS strings
    (``"Case=Nom|Number=Sing"``) and lemmas; ``ents``: ``(start, end_exclusive, label)`` token spans;
    ``heads``: absolute head index per token (root points at itself); ``deps``: per-token dependency
    label; ``sent_starts``: per-token True / False / None (unknown); ``cats``: ``{label: score}``
    document categories; ``spans``: named groups of labelled, possibly overlapping token spans."""

    __slots__ = ("words", "spaces", "tags", "ents", "heads", "deps", "_attrs", "user_data", "has_ents_annotation",
                 "pos", "morphs", "lemmas", "sent_starts", "cats", "spans")

    def __init__(
        self,
        words: Sequence[str],
        spaces: Optional[Sequence[bool]] = None,
        *,
        tags: Optional[Sequence[str]] = None,
        ents: Optional[Sequence[Tuple[int, int, str]]] = None,
        heads: Optional[Sequence[int]] = None,
        deps: Optional[Sequence[str]] = None,
        attrs: Optional[np.ndarray] = None,
        pos: Optional[Sequence[Optional[str]]] = None,
        morphs: Optional[Sequence[Optional[str]]] = None,
        lemmas: Optional[Sequence[Optional[str]]] = None,
        sent_starts: Optional[Sequence[Optional[bool]]] = None,
        cats: Optional[Dict[str, float]] = None,
        spans: Optional[Dict[str, Sequence[Tuple[int, int, str]]]] = None,
    ):
        self.words = list(words)
        self.spaces = list(spaces) if spaces is not None else [True] * len(self.words)
        self.tags = list(tags) if tags is not None else None
        self.ents = [tuple(e) for e in ents] if ents is not None else []
        self.has_ents_annotation = ents is not None
        self.heads = list(heads) if heads is not None else None
        self.deps = list(deps) if deps is not None else None
        self.pos = list(pos) if pos is not None else None
        self.morphs = list(morphs) if morphs is not None else None
        self.lemmas = list(lemmas) if lemmas is not None else None
        self.sent_starts = list(sent_starts) if sent_starts is not None else None
        self.cats = dict(cats) if cats else {}
        # named groups of (possibly overlapping) labelled token spans: {"sc": [(start, end_exclusive, label), ...]}
        self.spans = {str(k): [tuple(x) for x in v] for k, v in (spans or {}).items()}
        self._attrs = attrs
        self.user_data: Dict = {}
        for name in ("tags", "heads", "deps", "pos", "morphs", "lemmas", "sent_starts"):
            v = getattr(self, name)
            if v is not None and len(v) != len(self.words):
                raise ValueError(f"Doc: {name} has {len(v)} entries for {len(self.words)} tokens")

    def gold_sent_starts(self) -> Optional[List[Optional[bool]]]:
        """Sentence starts: the explicit annotation, else derived from the dependency tree (first token
        of every root's subtree), else None."""
        if self.sent_starts is not None:
            return self.sent_starts
        if self.heads is None:
            return None
        n = len(self.words)
        root_of = list(range(n))
        for i in range(n):
            j, guard = i, 0
            while 0 <= self.heads[j] < n and self.heads[j] != j and guard <= n:
                j, guard = self.heads[j], guard + 1
            root_of[i] = j
        seen, out = set(), [False] * n
        for i in range(n):
            if root_of[i] not in seen:
                seen.add(root_of[i])
                out[i] = True
        return out

    def __len__(self) -> int:
        return len(self.words)

    @property
    def text(self) -> str:
        return "".join(w + (" " if s else "") for w, s in zip(self.words, self.spaces)).strip()

    def to_array(self) -> np.ndarray:
        """``(n, 4)`` uint64: NORM, PREFIX, SUFFIX, SHAPE ids (cached)."""
        if self._attrs is None:
            self._attrs = featurize_words(self.words)
        return self._attrs

    def copy_unannotated(self) -> "Doc":
        return Doc(self.words, self.spaces, attrs=self._attrs)

    def to_dict(self) -> Dict:
        d: Dict = {"words": self.words, "spaces": self.spaces}
        if self.tags is not None:
            d["tags"] = self.tags
        if self.has_ents_annotation:
            d["ents"] = [list(e) for e in self.ents]
        if self.heads is not None:
            d["heads"] = self.heads
        if self.deps is not None:
            d["deps"] = self.deps
        for name in ("pos", "morphs", "lemmas", "sent_starts"):
            if getattr(self, name) is not None:
                d[name] = getattr(self, name)
        if self.cats:
            d["cats"] = self.cats
        if self.spans:
            d["spans"] = {k: [list(x) for x in v] for k, v in self.spans.items()}
        return d

    @classmethod
    def from_dict(cls, d: Dict) -> "Doc":
        words = d.get("words")
        cats = d.get("cats")
        if cats is None and isinstance(d.get("label"), str) and "answer" in d:       # Prodigy textcat rows
            cats = {d["label"]: 1.0 if d.get("answer") == "accept" else 0.0}
        if words is None:
            words, spans = _tokenize_with_offsets(d.get("text", ""))
            ents = None
            if "spans" in d or "entities" in d:
                ents = _char_spans_to_token_spans(spans, d.get("spans") or d.get("entities") or [])
            return cls(words, ents=ents, cats=cats)
        ents = d.get("ents")
        return cls(
            words, d.get("spaces"), tags=d.get("tags"),
            ents=[tuple(e) for e in ents] if ents is not None else None,
            heads=d.get("heads"), deps=d.get("deps"), pos=d.get("pos"), morphs=d.get("morphs"),
            lemmas=d.get("lemmas"), sent_starts=d.get("sent_starts"), cats=cats, spans=d.get("spans"),
        )

    def __repr__(self) -> str:
        return f"Doc({' '.join(self.words[:12])}{'...' if len(self.words) > 12 else ''})"


def _tokenize_with_offsets(text: str):
    """Whitespace + punctuation-splitting tokenizer (enough for JSONL corpora in
    the Prodigy ``{"text", "spans"}`` format that ``bin/get-data.sh`` of the
    reference downloads)."""
    import re

    words, spans = [], []
    for m in re.finditer(r"\w+(?:[-']\w+)*|[^\w\s]", text, flags=re.UNICODE):
        words.append(m.group(0))
        spans.append((m.start(), m.end()))
    return words, spans


def _char_spans_to_token_spans(tok_spans, char_spans) -> List[Tuple[int, int, str]]:
    out = []
    for sp in char_spans:
        if isinstance(sp, dict):
            cs, ce, lab = sp.get("start"), sp.get("end"), sp.get("label")
        else:
            cs, ce, lab = sp[0], sp[1], sp[2]
        idx = [i for i, (a, b) in enumerate(tok_spans) if a >= cs and b <= ce]
        if idx:
            out.append((idx[0], idx[-1] + 1, str(lab)))
    return out


class Example:
    """``predicted`` (what the pipeline sees/annotates) + ``reference`` (gold)."""

    __slots__ = ("predicted", "reference")

    def __init__(self, predicted: Doc, reference: Doc):
        if len(predicted) != len(reference):
            raise ValueError("Example: predicted and reference must share a tokenisation")
        self.predicted = predicted
        self.reference = reference

    @property
    def x(self) -> Doc:
        return self.predicted

    @property
    def y(self) -> Doc:
        return self.reference

    @classmethod
    def from_doc(cls, gold: Doc) -> "Example":
        return cls(gold.copy_unannotated(), gold)

    def __len__(self) -> int:
        return len(self.predicted)
