"""Pseudo-projective dependency parsing (Nivre & Nilsson 2005), the transform spaCy applies around
its arc-eager parser (``spacy/pipeline/_parser_internals/nonproj.pyx`` upstream; written here from the
paper's "HEAD" encoding, which is what spaCy uses).

Arc-eager can only derive projective trees.  Training data with crossing arcs is *projectivized*:
the shortest non-projective arc is repeatedly lifted (its dependent is re-attached to the head's
head) until the tree is projective, and every lifted token's label is decorated with the label of
its original head, ``dep||headdep``.  After parsing, *deprojectivize* undoes it: a token labelled
``a||b`` is moved down to the closest descendant of its current head whose own label is ``b``
(breadth first), and gets label ``a`` back.  Heads are absolute indices, a root points at itself.
"""
from __future__ import annotations

from collections import deque
from typing import List, Optional, Sequence, Tuple

DELIMITER = "||"


def _ancestors(tok: int, heads: Sequence[int]) -> List[int]:
    out, seen = [], {tok}
    h = heads[tok]
    while h != tok and h is not None and h >= 0 and h not in seen:
        out.append(h)
        seen.add(h)
        tok, h = h, heads[h]
    return out


def is_nonproj_arc(tok: int, heads: Sequence[int]) -> bool:
    """The arc head(tok) -> tok is non-projective iff some token strictly between the two is not a
    descendant of the head."""
    h = heads[tok]
    if h is None or h < 0 or h == tok:
        return False
    lo, hi = (h, tok) if h < tok else (tok, h)
    for k in range(lo + 1, hi):
        if h not in _ancestors(k, heads) and k != h:
            return True
    return False


def is_nonproj_tree(heads: Sequence[int]) -> bool:
    return any(is_nonproj_arc(t, heads) for t in range(len(heads)))


def _smallest_nonproj_arc(heads: Sequence[int]) -> Optional[int]:
    best, best_len = None, None
    for t in range(len(heads)):
        if is_nonproj_arc(t, heads):
            length = abs(heads[t] - t)
            if best is None or length < best_len:
                best, best_len = t, length
    return best


def projectivize(heads: Sequence[int], labels: Sequence[Optional[str]]) -> Tuple[List[int], List[Optional[str]]]:
    """-> (projective heads, decorated labels).  Tokens with a missing head (< 0 / None) are left alone."""
    proj = list(heads)
    n = len(proj)
    guard = 0
    while guard <= n * n:
        guard += 1
        t = _smallest_nonproj_arc(proj)
        if t is None:
            break
        h = proj[t]
        hh = proj[h]
        proj[t] = t if hh == h else hh          # lifting past a root makes the token a root
    deco = list(labels)
    for t in range(n):
        if proj[t] != heads[t] and labels[t] is not None and heads[t] is not None and heads[t] >= 0:
            head_label = labels[heads[t]]
            deco[t] = f"{labels[t]}{DELIMITER}{head_label}"
    return proj, deco


def is_decorated(label: Optional[str]) -> bool:
    return bool(label) and DELIMITER in label


def decompose(label: str) -> Tuple[str, str]:
    a, _, b = label.partition(DELIMITER)
    return a, b


def deprojectivize(heads: Sequence[int], labels: Sequence[Optional[str]]) -> Tuple[List[int], List[Optional[str]]]:
    heads, labels = list(heads), list(labels)
    n = len(heads)
    children: List[List[int]] = [[] for _ in range(n)]
    for t, h in enumerate(heads):
        if h != t and 0 <= h < n:
            children[h].append(t)
    for t in range(n):
        if not is_decorated(labels[t]):
            continue
        own, head_label = decompose(labels[t])
        labels[t] = own
        # closest descendant of the current head (breadth first, not through t itself) labelled head_label
        start = heads[t]
        queue = deque(c for c in children[start] if c != t) if start != t else deque()
        found = None
        while queue:
            c = queue.popleft()
            base = decompose(labels[c])[0] if is_decorated(labels[c]) else labels[c]
            if base == head_label:
                found = c
                break
            queue.extend(k for k in children[c] if k != t)
        if found is not None:
            if start != t and t in children[start]:
                children[start].remove(t)
            heads[t] = found
            children[found].append(t)
    return heads, labels


def sentence_starts(heads: Sequence[int]) -> List[bool]:
    """Sentence segmentation from a dependency forest: a token starts a sentence iff it is the
    leftmost token of a root's subtree (what spaCy derives from the parse; its BREAK transition
    exists only to commit to the boundary earlier)."""
    n = len(heads)
    root_of = []
    for t in range(n):
        k, guard = t, 0
        while heads[k] != k and 0 <= heads[k] < n and guard <= n:
            k, guard = heads[k], guard + 1
        root_of.append(k)
    seen, out = set(), []
    for t in range(n):
        out.append(root_of[t] not in seen)
        seen.add(root_of[t])
    return out
