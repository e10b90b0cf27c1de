"""Pseudo-projective transform: projectivize / deprojectivize round trips, parser integration."""
import random

from spacy_ray_b200.models.nonproj import (
    deprojectivize, is_nonproj_arc, is_nonproj_tree, projectivize, sentence_starts,
)
from spacy_ray_b200.models.transitions import ArcEagerSystem, is_projective


def test_known_example_from_the_literature():
    # "A hearing is scheduled on the issue today": hearing -> on is non-projective
    #        0    1      2   3        4   5    6     7
    heads = [1, 3, 3, 3, 1, 6, 4, 3]
    labels = ["det", "nsubjpass", "auxpass", "ROOT", "prep", "det", "pobj", "npadvmod"]
    assert is_nonproj_tree(heads) and is_nonproj_arc(4, heads) and not is_projective(heads)
    ph, pl = projectivize(heads, labels)
    assert not is_nonproj_tree(ph) and is_projective(ph)
    assert ph[4] == 3 and pl[4] == "prep||nsubjpass"
    assert [pl[i] for i in range(8) if i != 4] == [labels[i] for i in range(8) if i != 4]
    dh, dl = deprojectivize(ph, pl)
    assert dh == heads and dl == labels


def _random_tree(n, rng):
    root = rng.randrange(n)
    heads = [None] * n
    heads[root] = root
    order = [t for t in range(n) if t != root]
    rng.shuffle(order)
    attached = [root]
    for t in order:
        heads[t] = rng.choice(attached)
        attached.append(t)
    return heads


def _is_tree(heads):
    n = len(heads)
    for t in range(n):
        k, guard = t, 0
        while heads[k] != k and guard <= n:
            k, guard = heads[k], guard + 1
        if heads[k] != k:
            return False
    return True


def _mildly_nonprojective(n, rng):
    """A projective chain-like tree with ONE token re-attached somewhere else (treebank-like)."""
    root = rng.randrange(n)
    heads = [root if t == root else (t + 1 if t < root else t - 1) for t in range(n)]
    for _ in range(20):
        t = rng.randrange(n)
        h = rng.randrange(n)
        if t == root or h == t:
            continue
        cand = list(heads)
        cand[t] = h
        if _is_tree(cand) and is_nonproj_tree(cand):
            return cand
    return heads


def test_random_trees_projectivize_to_projective_and_roundtrip():
    rng = random.Random(0)
    exact = total = 0
    for i in range(400):
        n = rng.randint(2, 14)
        wild = i % 2 == 0
        heads = _random_tree(n, rng) if wild else _mildly_nonprojective(n, rng)
        labels = [f"l{t}" if h != t else "ROOT" for t, h in enumerate(heads)]        # distinct labels: unambiguous
        ph, pl = projectivize(heads, labels)
        assert is_projective(ph), (heads, ph)
        assert sum(1 for t in range(n) if ph[t] == t) >= 1
        if not is_nonproj_tree(heads):
            assert ph == heads and pl == labels
        dh, dl = deprojectivize(ph, pl)
        assert all("||" not in l for l in dl) and _is_tree(dh)
        if not wild:
            total += 1
            exact += dh == heads and dl == labels
        # an arc-eager derivation of the projectivized tree exists and reproduces it
        sys_ = ArcEagerSystem(sorted({l for l in pl if l != "ROOT"}))
        idx = {l: i for i, l in enumerate(sys_.labels)}
        gl = [idx.get(l, -1) for l in pl]
        st = sys_.init_state(n)
        for a in sys_.gold_sequence(ph, gl):
            sys_.apply(st, a)
        got, _ = sys_.finalize(st)
        assert got == ph
    assert exact == total, (exact, total)      # a single lifted arc with an unambiguous head label always comes back


def test_sentence_starts_from_a_forest():
    #  s1: 0 1 2 (root 1)   s2: 3 4 (root 3)   s3: 5
    assert sentence_starts([1, 1, 1, 3, 3, 5]) == [True, False, False, True, False, True]
    assert sentence_starts([0]) == [True]


def test_parser_component_projectivizes_gold_and_deprojectivizes_predictions():
    from spacy_ray_b200.pipeline.components import DependencyParser
    from spacy_ray_b200.pipeline.doc import Doc, Example

    heads = [1, 3, 3, 3, 1, 6, 4, 3]
    labels = ["det", "nsubjpass", "auxpass", "ROOT", "prep", "det", "pobj", "npadvmod"]
    gold = Doc(["A", "hearing", "is", "scheduled", "on", "the", "issue", "today"], heads=heads, deps=labels)

    class _M:                       # the component only needs has_dim / set_dim / initialize here
        def has_dim(self, n): return None
        def set_dim(self, n, v): self.nO = v
        def initialize(self): return self
        def walk(self): return []

    p = DependencyParser("parser", _M(), min_action_freq=1)
    p.initialize(lambda: [Example.from_doc(gold)])
    assert "prep||nsubjpass" in p.labels
    gh, gl = p._gold(gold)
    assert gh[4] == 3 and p.labels[gl[4]] == "prep||nsubjpass" and is_projective(gh)

    class Out:
        states = None
        heads_flat = __import__("torch").tensor(gh)
        labels_flat = __import__("torch").tensor(gl)

    doc = gold.copy_unannotated()
    p.set_annotations([doc], Out())
    assert doc.heads == heads and doc.deps[4] == "prep" and "||" not in "".join(doc.deps)
    assert doc.user_data["sent_starts"][0] is True and sum(doc.user_data["sent_starts"]) == 1
