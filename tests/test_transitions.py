import random

import torch

from spacy_ray_b200.models.transitions import (
    ArcEagerSystem, BiluoSystem, biluo_action, biluo_actions_to_spans, is_projective, spans_to_biluo_actions,
    B_, I_, L_, U_, OUT,
)


def test_biluo_encoding_roundtrip():
    spans = [(0, 1, 0), (2, 5, 1), (6, 8, 0)]
    acts = spans_to_biluo_actions(9, spans)
    assert acts[0] == biluo_action(U_, 0) and acts[2] == biluo_action(B_, 1) and acts[4] == biluo_action(L_, 1)
    assert acts[1] == OUT and acts[8] == OUT
    assert biluo_actions_to_spans(acts) == spans


def test_biluo_following_the_oracle_reproduces_gold_and_always_has_a_valid_action():
    sys = BiluoSystem(["A", "B"])
    rng = random.Random(0)
    for _ in range(50):
        n = rng.randint(1, 12)
        spans, t = [], 0
        while t < n:
            if rng.random() < 0.3:
                ln = min(n - t, rng.randint(1, 3))
                spans.append((t, t + ln, rng.randint(0, 1)))
                t += ln
            t += 1
        gold = spans_to_biluo_actions(n, spans)
        s = sys.init_state(n)
        while not s.is_final:
            v = sys.valid(s)
            assert any(v)
            costs = sys.costs(s, gold)
            a = min(range(sys.n_actions), key=lambda k: costs[k])
            assert costs[a] == 0 and v[a]
            sys.apply(s, a, gold)
        assert biluo_actions_to_spans(s.history) == spans


def test_biluo_batched_matches_per_doc_reference_under_random_policies():
    sys = BiluoSystem(["A", "B", "C"])
    rng = random.Random(1)
    lens = [rng.randint(1, 9) for _ in range(17)]
    golds = []
    for n in lens:
        spans, t = [], 0
        while t < n:
            if rng.random() < 0.35:
                ln = min(n - t, rng.randint(1, 3))
                spans.append((t, t + ln, rng.randint(0, 2)))
                t += ln
            t += 1
        golds.append(spans_to_biluo_actions(n, spans))
    flat = torch.tensor([a for g in golds for a in g])
    offs = torch.tensor([sum(lens[:i]) for i in range(len(lens))])
    starts = torch.tensor([1 + sum(lens[:i]) + i for i in range(len(lens))])
    st = sys.batch_init(torch.tensor(lens))
    states = [sys.init_state(n) for n in lens]
    for _step in range(max(lens)):
        valid = sys.batch_valid(st)
        feats = sys.batch_features(st, starts)
        ga = sys.batch_gold_action(st, flat, offs)
        chosen = torch.zeros(len(lens), dtype=torch.int64)
        for d, s in enumerate(states):
            if s.is_final:
                assert not valid[d].any()
                continue
            assert valid[d].tolist() == sys.valid(s)
            ref_feats = [(int(starts[d]) + f) if f >= 0 else -1 for f in sys.features(s)]
            assert feats[d].tolist() == ref_feats
            ref_ga = sys.gold_action(s, golds[d])
            if ref_ga >= 0 and not sys.valid(s)[ref_ga]:
                ref_ga = int(ga[d])          # both sides treat an invalid gold as "no constraint" later
            assert int(ga[d]) == sys.gold_action(s, golds[d])
            options = [a for a, ok in enumerate(sys.valid(s)) if ok]
            a = rng.choice(options)           # random (often wrong) policy exercises the 'sunk' logic
            chosen[d] = a
            sys.apply(s, a, golds[d])
        st = sys.batch_apply(st, chosen, flat, offs)
        for d, s in enumerate(states):
            assert int(st["i"][d]) == s.i and int(st["ent_start"][d]) == s.ent_start
            assert bool(st["ent_ok"][d]) == (s.ent_ok if s.ent_start >= 0 else False)


def _random_projective(n, rng):
    root = rng.randrange(n)
    heads = []
    for t in range(n):
        if t == root:
            heads.append(t)
        elif t < root:
            heads.append(root if rng.random() < 0.3 else t + 1)
        else:
            heads.append(root if rng.random() < 0.3 else t - 1)
    return heads


def test_arc_eager_oracle_reaches_the_gold_tree():
    sys = ArcEagerSystem(["x", "y"])
    rng = random.Random(0)
    for _ in range(40):
        n = rng.randint(1, 10)
        heads = _random_projective(n, rng)
        assert is_projective(heads)
        labels = [rng.randint(0, 1) if h != t else -1 for t, h in enumerate(heads)]
        seq = sys.gold_sequence(heads, labels)
        s = sys.init_state(n)
        for a in seq:
            assert sys.valid(s)[a]
            assert sys.costs(s, heads, labels)[a] == 0
            sys.apply(s, a)
        assert s.is_final
        got_heads, got_labels = sys.finalize(s)
        assert got_heads == heads
        assert all(gl == l for gl, l, h, t in zip(got_labels, labels, heads, range(n)) if h != t)


def test_arc_eager_random_walk_terminates_with_valid_actions():
    sys = ArcEagerSystem(["x"])
    rng = random.Random(3)
    for _ in range(30):
        n = rng.randint(1, 9)
        s = sys.init_state(n)
        steps = 0
        while not s.is_final:
            options = [a for a, ok in enumerate(sys.valid(s)) if ok]
            assert options
            sys.apply(s, rng.choice(options))
            steps += 1
            assert steps <= 4 * n + 4
        heads, _ = sys.finalize(s)
        assert len(heads) == n
