"""Transition systems for the NER and dependency-parser heads.

Upstream these live in spaCy's Cython ``_parser_internals`` (C++ state
machine driven per step from the CPU - SURVEY.md 2.7 K7: "this loop is the real
bottleneck upstream").  Here each system has

* a plain-Python per-doc ``State`` implementation (readable spec, used by the
  CPU path and by differential tests), and
* for BILUO a *vectorised* formulation over a whole batch - every quantity is a
  tensor indexed by doc - which is what the persistent sm_100a kernel
  (``ops/csrc/ner_kernels.cu``) implements with one warp per doc.

Action numbering
----------------
BILUO (``L`` labels, ``4L+1`` actions): ``0 = OUT``; for label j:
``1+4j = BEGIN-j``, ``2+4j = IN-j``, ``3+4j = LAST-j``, ``4+4j = UNIT-j``.

Arc-eager (``L`` labels, ``2L+2`` actions): ``0 = SHIFT``, ``1 = REDUCE``,
``2+2j = LEFT-j``, ``3+2j = RIGHT-j``.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch

# ============================================================================
# BILUO
# ============================================================================
OUT = 0
B_, I_, L_, U_ = 0, 1, 2, 3


def biluo_action(kind: int, label: int) -> int:
    return 1 + 4 * label + kind


def biluo_decode(action: int) -> Tuple[int, int]:
    """action -> (kind, label) with kind in {B_,I_,L_,U_} or (-1,-1) for OUT."""
    if action <= 0:
        return -1, -1
    a = action - 1
    return a % 4, a // 4


def spans_to_biluo_actions(n: int, spans: Sequence[Tuple[int, int, int]]) -> List[int]:
    """Token spans ``(start, end_exclusive, label_index)`` -> per-token gold
    action ids.  Overlapping/invalid spans are dropped (first wins)."""
    acts = [OUT] * n
    taken = [False] * n
    for start, end, lab in sorted(spans):
        if start < 0 or end > n or end <= start or any(taken[start:end]):
            continue
        for t in range(start, end):
            taken[t] = True
        if end - start == 1:
            acts[start] = biluo_action(U_, lab)
        else:
            acts[start] = biluo_action(B_, lab)
            for t in range(start + 1, end - 1):
                acts[t] = biluo_action(I_, lab)
            acts[end - 1] = biluo_action(L_, lab)
    return acts


def biluo_actions_to_spans(acts: Sequence[int]) -> List[Tuple[int, int, int]]:
    spans = []
    start, lab = -1, -1
    for t, a in enumerate(acts):
        kind, l = biluo_decode(a)
        if kind == U_:
            spans.append((t, t + 1, l))
            start = -1
        elif kind == B_:
            start, lab = t, l
        elif kind == L_ and start >= 0:
            spans.append((start, t + 1, lab))
            start = -1
        elif kind == -1:
            start = -1
    return spans


@dataclass
class BiluoState:
    n: int
    i: int = 0
    ent_start: int = -1
    ent_label: int = -1
    ent_ok: bool = False
    history: List[int] = field(default_factory=list)

    @property
    def is_final(self) -> bool:
        return self.i >= self.n


class BiluoSystem:
    """NER transition system; one action per token."""

    n_features = 3

    def __init__(self, labels: Sequence[str]):
        self.labels = list(labels)
        self.n_labels = len(self.labels)
        self.n_actions = 4 * self.n_labels + 1

    # ---- names -----------------------------------------------------------
    def action_name(self, a: int) -> str:
        kind, lab = biluo_decode(a)
        if kind < 0:
            return "O"
        return "BILU"[kind] + "-" + self.labels[lab]

    # ---- per-doc reference -----------------------------------------------
    def init_state(self, n: int) -> BiluoState:
        return BiluoState(n=n)

    def features(self, s: BiluoState) -> List[int]:
        """[B0, first word of the open entity or -1, B0-1 if an entity is open]."""
        b0 = s.i if s.i < s.n else -1
        e0 = s.ent_start if s.ent_start >= 0 else -1
        last = b0 - 1 if (b0 >= 0 and e0 >= 0) else -1
        return [b0, e0, last]

    def valid(self, s: BiluoState) -> List[bool]:
        v = [False] * self.n_actions
        if s.is_final:
            return v
        is_open = s.ent_start >= 0
        not_last = s.i + 1 < s.n
        if not is_open:
            v[OUT] = True
            for j in range(self.n_labels):
                v[biluo_action(U_, j)] = True
                v[biluo_action(B_, j)] = not_last
        else:
            v[biluo_action(L_, s.ent_label)] = True
            v[biluo_action(I_, s.ent_label)] = not_last
        return v

    def gold_action(self, s: BiluoState, gold: Sequence[int]) -> int:
        """The single zero-cost action, or -1 meaning "every valid action is
        zero-cost" (gold missing, or the open entity is already wrong - 'sunk' -
        so nothing done with this token can lose anything more)."""
        g = gold[s.i]
        if g < 0:
            return -1
        kind, lab = biluo_decode(g)
        if s.ent_start < 0:
            return g if kind in (-1, B_, U_) else OUT
        if s.ent_ok and kind in (I_, L_) and lab == s.ent_label:
            return g
        return -1

    def costs(self, s: BiluoState, gold: Sequence[int]) -> List[int]:
        ga = self.gold_action(s, gold)
        v = self.valid(s)
        if ga < 0 or not v[ga]:
            return [0 if ok else 9 for ok in v]
        return [(0 if a == ga else 1) if ok else 9 for a, ok in enumerate(v)]

    def apply(self, s: BiluoState, action: int, gold: Optional[Sequence[int]] = None) -> None:
        kind, lab = biluo_decode(action)
        g = gold[s.i] if gold is not None else -2
        if kind == B_:
            s.ent_start, s.ent_label = s.i, lab
            s.ent_ok = g == action
        elif kind == I_:
            s.ent_ok = s.ent_ok and g == action
        elif kind in (L_, U_, -1):
            s.ent_start, s.ent_label, s.ent_ok = -1, -1, False
        s.history.append(action)
        s.i += 1

    def gold_sequence(self, gold: Sequence[int]) -> List[int]:
        return [g if g >= 0 else OUT for g in gold]

    # ---- vectorised (batch) formulation ------------------------------------
    def batch_init(self, lens: torch.Tensor) -> Dict[str, torch.Tensor]:
        B = lens.shape[0]
        dev = lens.device
        return {
            "i": torch.zeros(B, dtype=torch.int64, device=dev),
            "ent_start": torch.full((B,), -1, dtype=torch.int64, device=dev),
            "ent_label": torch.full((B,), -1, dtype=torch.int64, device=dev),
            "ent_ok": torch.zeros(B, dtype=torch.bool, device=dev),
            "n": lens.to(torch.int64),
        }

    def batch_active(self, st) -> torch.Tensor:
        return st["i"] < st["n"]

    def batch_features(self, st, starts: torch.Tensor) -> torch.Tensor:
        """(B, 3) *row* indices into the padded token array (-1 = missing)."""
        active = self.batch_active(st)
        b0 = torch.where(active, starts + st["i"], torch.full_like(st["i"], -1))
        is_open = st["ent_start"] >= 0
        e0 = torch.where(is_open & active, starts + st["ent_start"], torch.full_like(b0, -1))
        last = torch.where((b0 >= 0) & (e0 >= 0), b0 - 1, torch.full_like(b0, -1))
        return torch.stack([b0, e0, last], dim=1)

    def batch_valid(self, st) -> torch.Tensor:
        B = st["i"].shape[0]
        dev = st["i"].device
        A = self.n_actions
        active = self.batch_active(st)
        is_open = st["ent_start"] >= 0
        not_last = (st["i"] + 1) < st["n"]
        acts = torch.arange(A, device=dev)
        kind = torch.where(acts > 0, (acts - 1) % 4, torch.full_like(acts, -1))       # (A,)
        lab = torch.where(acts > 0, (acts - 1) // 4, torch.full_like(acts, -1))
        closed_ok = (kind == -1) | (kind == U_)                                        # OUT, UNIT
        closed_ok = closed_ok.unsqueeze(0) | ((kind == B_).unsqueeze(0) & not_last.unsqueeze(1))
        same = lab.unsqueeze(0) == st["ent_label"].unsqueeze(1)
        open_ok = same & ((kind == L_).unsqueeze(0) | ((kind == I_).unsqueeze(0) & not_last.unsqueeze(1)))
        v = torch.where(is_open.unsqueeze(1), open_ok, closed_ok.expand(B, A))
        return v & active.unsqueeze(1)

    def batch_gold_action(self, st, gold: torch.Tensor, gold_offsets: torch.Tensor) -> torch.Tensor:
        """(B,) single gold action or -1 (= all valid actions are gold).
        ``gold`` is the flat per-token gold-action array (doc order, unpadded),
        ``gold_offsets`` each doc's offset into it."""
        active = self.batch_active(st)
        idx = (gold_offsets + st["i"]).clamp(max=gold.shape[0] - 1)
        g = torch.where(active, gold[idx], torch.full_like(st["i"], -1))
        kind = torch.where(g > 0, (g - 1) % 4, torch.full_like(g, -1))
        lab = torch.where(g > 0, (g - 1) // 4, torch.full_like(g, -1))
        is_open = st["ent_start"] >= 0
        closed_gold = torch.where((kind == -1) | (kind == B_) | (kind == U_), g, torch.zeros_like(g))
        open_match = st["ent_ok"] & ((kind == I_) | (kind == L_)) & (lab == st["ent_label"])
        open_gold = torch.where(open_match, g, torch.full_like(g, -1))
        out = torch.where(is_open, open_gold, closed_gold)
        return torch.where(g < 0, torch.full_like(g, -1), out)

    def batch_apply(self, st, actions: torch.Tensor, gold: Optional[torch.Tensor], gold_offsets: Optional[torch.Tensor]):
        active = self.batch_active(st)
        kind = torch.where(actions > 0, (actions - 1) % 4, torch.full_like(actions, -1))
        lab = torch.where(actions > 0, (actions - 1) // 4, torch.full_like(actions, -1))
        if gold is not None:
            idx = (gold_offsets + st["i"]).clamp(max=max(gold.shape[0] - 1, 0))
            g = gold[idx] if gold.shape[0] else torch.full_like(actions, -2)
        else:
            g = torch.full_like(actions, -2)
        begin = active & (kind == B_)
        cont = active & (kind == I_)
        close = active & ((kind == L_) | (kind == U_) | (kind == -1))
        st["ent_start"] = torch.where(begin, st["i"], torch.where(close, torch.full_like(st["i"], -1), st["ent_start"]))
        st["ent_label"] = torch.where(begin, lab, torch.where(close, torch.full_like(lab, -1), st["ent_label"]))
        ok_begin = g == actions
        st["ent_ok"] = torch.where(begin, ok_begin, torch.where(cont, st["ent_ok"] & (g == actions), st["ent_ok"] & ~close))
        st["i"] = torch.where(active, st["i"] + 1, st["i"])
        return st


# ============================================================================
# Arc-eager
# ============================================================================
SHIFT, REDUCE = 0, 1


def arc_action(is_right: bool, label: int) -> int:
    return 2 + 2 * label + (1 if is_right else 0)


def arc_decode(a: int) -> Tuple[str, int]:
    if a == SHIFT:
        return "S", -1
    if a == REDUCE:
        return "D", -1
    a -= 2
    return ("R" if a % 2 else "L"), a // 2


@dataclass
class ArcState:
    n: int
    stack: List[int] = field(default_factory=list)
    b: int = 0
    heads: List[int] = field(default_factory=list)
    labels: List[int] = field(default_factory=list)
    lefts: List[List[int]] = field(default_factory=list)
    rights: List[List[int]] = field(default_factory=list)
    history: List[int] = field(default_factory=list)

    def __post_init__(self):
        self.heads = [-1] * self.n
        self.labels = [-1] * self.n
        self.lefts = [[] for _ in range(self.n)]
        self.rights = [[] for _ in range(self.n)]

    @property
    def is_final(self) -> bool:
        return self.b >= self.n and len(self.stack) <= 1

    def S(self, k: int) -> int:
        return self.stack[-1 - k] if len(self.stack) > k else -1

    def B(self, k: int) -> int:
        return self.b + k if self.b + k < self.n else -1


class ArcEagerSystem:
    """Arc-eager dependency parsing with a dynamic oracle (Goldberg & Nivre 2012).

    Simplifications vs spaCy: no BREAK action (one tree per doc; tokens left
    without a head at the end become roots), REDUCE is also allowed on a
    headless S0 once the buffer is empty (that token becomes a root)."""

    n_features = 8

    def __init__(self, labels: Sequence[str]):
        self.labels = list(labels)
        self.n_labels = len(self.labels)
        self.n_actions = 2 + 2 * self.n_labels

    def action_name(self, a: int) -> str:
        k, l = arc_decode(a)
        return {"S": "SHIFT", "D": "REDUCE"}.get(k) or f"{k}-{self.labels[l]}"

    def init_state(self, n: int) -> ArcState:
        return ArcState(n=n)

    def features(self, s: ArcState) -> List[int]:
        b0, b1 = s.B(0), s.B(1)
        s0, s1, s2 = s.S(0), s.S(1), s.S(2)
        lb0 = s.lefts[b0][0] if b0 >= 0 and s.lefts[b0] else -1
        ls0 = s.lefts[s0][0] if s0 >= 0 and s.lefts[s0] else -1
        rs0 = s.rights[s0][-1] if s0 >= 0 and s.rights[s0] else -1
        return [b0, b1, s0, s1, s2, lb0, ls0, rs0]

    def valid(self, s: ArcState) -> List[bool]:
        v = [False] * self.n_actions
        if s.is_final:
            return v
        has_buf = s.b < s.n
        has_stack = len(s.stack) > 0
        s0 = s.S(0)
        v[SHIFT] = has_buf
        if has_stack:
            s0_headed = s.heads[s0] >= 0
            v[REDUCE] = s0_headed or not has_buf
            if has_buf:
                for j in range(self.n_labels):
                    v[arc_action(True, j)] = True
                    v[arc_action(False, j)] = not s0_headed
        return v

    def costs(self, s: ArcState, gold_heads: Sequence[int], gold_labels: Sequence[int]) -> List[int]:
        """Dynamic-oracle cost per action (9 = invalid).  ``gold_heads[t] == t``
        marks a root; ``-1`` marks a missing annotation (never costs)."""
        v = self.valid(s)
        c = [9] * self.n_actions
        if s.is_final:
            return c
        b0, s0 = s.B(0), s.S(0)
        in_stack = set(s.stack)
        buf = range(s.b, s.n)

        def ghead(t):
            h = gold_heads[t]
            return -2 if h < 0 else (-1 if h == t else h)   # -2 missing, -1 root

        if v[SHIFT]:
            cost = 0
            # pushing B0: loses arcs between B0 and anything in the stack
            if ghead(b0) in in_stack:
                cost += 1
            cost += sum(1 for k in s.stack if s.heads[k] < 0 and ghead(k) == b0)
            c[SHIFT] = cost
        if v[REDUCE]:
            cost = sum(1 for k in buf if ghead(k) == s0)
            if s.heads[s0] < 0 and s.b >= s.n:
                cost = 0  # forced root at the end
            c[REDUCE] = cost
        if b0 >= 0 and s0 >= 0:
            # LEFT: S0 gets head B0 and is popped.  Loses S0's gold head if that is
            # the root or still in the buffer (other than B0), and every gold child
            # of S0 still in the buffer.
            gs = ghead(s0)
            base_left = sum(1 for k in buf if ghead(k) == s0)
            if gs != -2 and gs != b0 and (gs == -1 or gs > b0):
                base_left += 1
            # RIGHT: B0 gets head S0 and is pushed.  Loses B0's gold head if it is
            # the root, elsewhere in the stack or later in the buffer, and every
            # headless stack item whose gold head is B0.
            base_right = 0
            gb = ghead(b0)
            if gb != -2 and gb != s0:
                if gb == -1 or gb in in_stack or gb > b0:
                    base_right += 1
            base_right += sum(1 for k in s.stack if s.heads[k] < 0 and ghead(k) == b0)
            for j in range(self.n_labels):
                la, ra = arc_action(False, j), arc_action(True, j)
                if v[la]:
                    lab_cost = 1 if (ghead(s0) == b0 and gold_labels[s0] >= 0 and gold_labels[s0] != j) else 0
                    c[la] = base_left + lab_cost
                if v[ra]:
                    lab_cost = 1 if (gb == s0 and gold_labels[b0] >= 0 and gold_labels[b0] != j) else 0
                    c[ra] = base_right + lab_cost
        return c

    def apply(self, s: ArcState, action: int) -> None:
        kind, lab = arc_decode(action)
        if kind == "S":
            s.stack.append(s.b)
            s.b += 1
        elif kind == "D":
            s.stack.pop()
        elif kind == "L":
            child, head = s.stack.pop(), s.b
            s.heads[child], s.labels[child] = head, lab
            s.lefts[head].append(child)
            s.lefts[head].sort()
        else:
            child, head = s.b, s.stack[-1]
            s.heads[child], s.labels[child] = head, lab
            s.rights[head].append(child)
            s.stack.append(child)
            s.b += 1
        s.history.append(action)

    def finalize(self, s: ArcState) -> Tuple[List[int], List[int]]:
        heads = [h if h >= 0 else t for t, h in enumerate(s.heads)]
        return heads, list(s.labels)

    def gold_sequence(self, gold_heads: Sequence[int], gold_labels: Sequence[int]) -> List[int]:
        """A zero-cost action sequence for a projective gold tree (static
        oracle by following the dynamic oracle's zero-cost choices)."""
        s = self.init_state(len(gold_heads))
        out = []
        guard = 0
        while not s.is_final and guard < 4 * len(gold_heads) + 8:
            costs = self.costs(s, gold_heads, gold_labels)
            a = min(range(self.n_actions), key=lambda k: (costs[k], _ARC_PREF(k)))
            self.apply(s, a)
            out.append(a)
            guard += 1
        return out


def _ARC_PREF(a: int) -> int:
    # tie-break among zero-cost actions: prefer arcs, then reduce, then shift
    k, _ = arc_decode(a)
    return {"L": 0, "R": 1, "D": 2, "S": 3}[k]


def is_projective(heads: Sequence[int]) -> bool:
    n = len(heads)
    arcs = [(min(t, h), max(t, h)) for t, h in enumerate(heads) if h != t and h >= 0]
    for a1, b1 in arcs:
        for a2, b2 in arcs:
            if a1 < a2 < b1 < b2:
                return False
    return True
