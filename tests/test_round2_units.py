"""CPU unit tests for the round-2 host logic: bucket planning for the overlapped exchange, LPT
shard balance, RAdam / averages semantics, resume-safe optimizer state loading, checkpoint shards
with fp32 master, equal step counts across ranks, async-proxy resume, token-balanced batches."""
import math

import numpy as np
import pytest
import torch

from spacy_ray_b200.nn import reset_model_ids
from spacy_ray_b200.nn.layers import HashEmbedCNN, fix_random_seed
from spacy_ray_b200.parallel.sync_proxy import FlatLayout
from spacy_ray_b200.parallel.util import DIVIDERS, divide_params, divide_params_lpt
from spacy_ray_b200.training.optimizer import Optimizer, radam_step_size


def _model(width=32, depth=3):
    reset_model_ids()
    fix_random_seed(0)
    return HashEmbedCNN(width, depth, 300).initialize()


def _numel(model):
    return {(n.id, p): int(n.get_param(p).numel()) for n in model.walk() for p in n.param_names if n.has_param(p)}


# ---------------------------------------------------------------- partition
@pytest.mark.parametrize("world", [1, 2, 3, 8, 40])
def test_lpt_partition_owns_every_key_once_keeps_nodes_together_and_balances(world):
    m = _model()
    sizes = _numel(m)
    shares = divide_params_lpt(m, world)
    assert len(shares) == world
    flat = [k for s in shares for k in s]
    assert sorted(flat) == sorted(sizes) and len(set(flat)) == len(flat)
    owner = {k: r for r, s in enumerate(shares) for k in s}
    for node in m.walk():
        owners = {owner[(node.id, p)] for p in node.param_names}
        assert len(owners) <= 1, "a node's parameters must share an owner"
    loads = [sum(sizes[k] for k in s) for s in shares]
    groups = {}
    for k, n in sizes.items():
        groups[k[0]] = groups.get(k[0], 0) + n
    # LPT guarantee: max load <= mean + largest item
    assert max(loads) <= sum(loads) / world + max(groups.values())
    ref_loads = [sum(sizes[k] for k in s) for s in divide_params(m, world)]
    assert max(loads) <= max(ref_loads)
    assert divide_params_lpt(m, world) == shares, "must be deterministic (every rank derives the same map)"


def test_layout_build_accepts_every_balance_mode_and_rejects_unknown():
    m = _model()
    for mode in DIVIDERS:
        lay = FlatLayout.build([("t2v", m)], 4, balance=mode)
        assert sorted(lay.keys) == sorted(_numel(m))
    with pytest.raises(ValueError):
        FlatLayout.build([("t2v", m)], 4, balance="nope")


# ---------------------------------------------------------------- bucket plan
def _layout_and_order(world=4):
    m = _model(depth=4)
    lay = FlatLayout.build([("t2v", m)], world, balance="lpt")
    # backward order: encoder blocks last-to-first (W, b, G, b per block), then mix, then the tables
    nodes = [n for n in m.walk(order="dfs_pre") if n.param_names]
    order = []
    for n in reversed(nodes):
        order.extend((n.id, p) for p in n.param_names)
    return m, lay, order


def test_plan_buckets_partitions_in_order_tables_last_and_alone():
    from spacy_ray_b200.parallel.fused_comm import plan_buckets

    m, lay, order = _layout_and_order()
    plan = plan_buckets(order, lay, n_target=4)
    flat = [k for b in plan.buckets for k in b]
    assert flat == [k for k in order if k in lay.numel], "buckets must be contiguous runs of the completion order"
    assert 3 <= plan.n <= 32
    kinds = [{k[1] == "E" for k in b} for b in plan.buckets]
    assert all(len(s) == 1 for s in kinds), "embedding tables never share a bucket with other parameters"
    assert kinds[-1] == {True}, "the tables are the last bucket"
    # a node's keys stay in one bucket
    for b in plan.buckets:
        ids = [k[0] for k in b]
        for other in plan.buckets:
            if other is not b:
                assert not set(ids) & {k[0] for k in other}
    assert all(plan.bucket_of[k] == i for i, b in enumerate(plan.buckets) for k in b)


def test_plan_buckets_handles_missing_and_duplicate_keys_and_caps_the_count():
    from spacy_ray_b200.parallel.fused_comm import plan_buckets

    m, lay, order = _layout_and_order()
    half = order[: len(order) // 2]
    plan = plan_buckets(half + half + [(999, "W")], lay, n_target=3)
    flat = [k for b in plan.buckets for k in b]
    assert sorted(flat) == sorted(lay.keys), "keys without a gradient ride in the last bucket"
    assert len(flat) == len(set(flat))
    many = plan_buckets(order, lay, n_target=1000, max_buckets=5)
    assert many.n <= 5 and sorted(k for b in many.buckets for k in b) == sorted(lay.keys)
    empty = plan_buckets([], lay, n_target=4)
    assert empty.n == 1 and sorted(empty.buckets[0]) == sorted(lay.keys)


def test_shard_tables_ranges_cover_each_ranks_keys_bucket_by_bucket():
    from spacy_ray_b200.parallel.fused_comm import plan_buckets, shard_tables

    m, lay, order = _layout_and_order(world=3)
    plan = plan_buckets(order, lay, n_target=4)
    for rank in range(3):
        t = shard_tables(lay, rank, torch.device("cpu"), plan)
        assert sorted(t["keys"]) == sorted(lay.owned_keys(rank))
        assert len(t["ranges"]) == plan.n
        prev_b, prev_k = 0, 0
        for b, (bb, be, kb, ke) in enumerate(t["ranges"]):
            assert bb == prev_b and kb == prev_k and be >= bb and ke >= kb
            prev_b, prev_k = be, ke
            assert all(plan.bucket_of[k] == b for k in t["keys"][kb:ke])
            assert set(t["blk_key"][bb:be].tolist()) == set(range(kb, ke)) or kb == ke
        assert prev_b == t["blk_key"].numel() and prev_k == len(t["keys"])
        s0 = lay.shard_start[rank]
        for i, k in enumerate(t["keys"]):
            assert int(t["key_off"][i]) == lay.offset[k] - s0
            assert int(t["key_len"][i]) % 128 == 0 and int(t["key_len"][i]) >= lay.numel[k]


# ---------------------------------------------------------------- optimizer
def test_radam_follows_the_rectified_formula_and_degenerates_early():
    b1, b2, lr = 0.9, 0.999, 0.01
    opt = Optimizer(lr, use_radam=True, L2=0.0, grad_clip=0.0)
    w = torch.tensor([0.5, -0.25, 1.0])
    g = torch.tensor([0.1, -0.2, 0.3])
    m1 = torch.zeros(3)
    m2 = torch.zeros(3)
    want = w.clone()
    saw = set()
    for t in range(1, 12):
        opt((1, "W"), w, g.clone())
        m2 = b2 * m2 + (1 - b2) * g * g
        m1 = b1 * m1 + (1 - b1) * g
        step, rect = radam_step_size(t, b1, b2)
        saw.add(rect)
        want = want - lr * step * (m1 / (m2.sqrt() + 1e-8) if rect else m1)
        assert torch.allclose(w, want, rtol=1e-5, atol=1e-7), t
    assert saw == {True, False}, "both the SGD-like warm-up and the rectified regime must be exercised"
    sma_max = 2 / (1 - b2) - 1
    assert radam_step_size(10 ** 6, b1, b2)[0] == pytest.approx(1.0, rel=1e-2) and sma_max > 5


def test_averages_follow_thinc_update_averages():
    opt = Optimizer(0.1, use_averages=True, L2=0.0, grad_clip=0.0, use_adam=False)
    w = torch.tensor([1.0, 2.0])
    ema = torch.zeros(2)
    for t in range(1, 6):
        opt((1, "W"), w, torch.tensor([0.5, -0.5]))
        decay = min((1 + t) / (10 + t), 0.9999)
        ema = ema - (1 - decay) * (ema - w)
        assert torch.allclose(opt.averages[(1, "W")], ema, rtol=1e-6)


def test_load_state_dict_copies_into_existing_views():
    """FusedSymmComm.bind exposes the kernel's flat moment buffers as per-key views; resuming must
    fill those views, not replace the dict entries."""
    opt = Optimizer(0.01)
    flat1, flat2 = torch.zeros(10), torch.zeros(10)
    key = (3, "W")
    opt.mom1[key], opt.mom2[key] = flat1[2:8].view(2, 3), flat2[2:8].view(2, 3)
    state = {"mom1": {key: torch.arange(6.0).view(2, 3)}, "mom2": {key: torch.ones(2, 3)}, "nr_update": {key: 7},
             "averages": None, "step": 0}
    opt.load_state_dict(state, device="cpu")
    assert flat1[2:8].tolist() == [0, 1, 2, 3, 4, 5] and flat2[2:8].tolist() == [1] * 6
    assert opt.mom1[key].data_ptr() == flat1[2:8].data_ptr()
    assert opt.nr_update[key] == 7


def test_load_state_dict_fast_forwards_schedules():
    def sched():
        v = 1.0
        while True:
            yield v
            v *= 0.5

    opt = Optimizer(sched())
    for _ in range(3):
        opt.step_schedules()
    lr3 = opt.learn_rate
    fresh = Optimizer(sched())
    fresh.load_state_dict({"mom1": {}, "mom2": {}, "nr_update": {}, "averages": None, "step": 3}, device="cpu")
    assert fresh.learn_rate == lr3 and fresh.step == 3


# ---------------------------------------------------------------- checkpoint shards
def test_optimizer_shards_roundtrip_master_atomic_and_drop_stale_world_sizes(tmp_path, tagger_config):
    from spacy_ray_b200.training.checkpoint import load_optimizer_shards, save_optimizer_shard
    from spacy_ray_b200.training.initialize import init_nlp

    nlp = init_nlp(tagger_config, use_gpu=-1)
    model = nlp.get_pipe("tagger").model
    keys = [(n.id, p) for n in model.walk() for p in n.param_names if n.has_param(p)]
    halves = [keys[::2], keys[1::2]]
    (tmp_path / "optim").mkdir()
    (tmp_path / "optim" / "rank0-of5.pt").write_bytes(b"stale")
    for rank in (0, 1):
        opt = Optimizer(0.01)
        for k in halves[rank]:
            opt.mom1[k] = torch.full((2,), float(rank + 1))
            opt.mom2[k] = torch.full((2,), 0.5)
            opt.nr_update[k] = 11
        master = {k: torch.full((3,), 7.0 + rank) for k in halves[rank]}
        save_optimizer_shard(tmp_path, nlp, opt, halves[rank], rank=rank, world_size=2, master=master,
                             extra={"version": 11})
    names = sorted(p.name for p in (tmp_path / "optim").iterdir())
    assert names == ["rank0-of2.pt", "rank1-of2.pt"], names          # stale world size removed, no temp files left
    # same world size: own file suffices
    opt = Optimizer(0.01)
    got = load_optimizer_shards(tmp_path, nlp, opt, halves[1], rank=1, world_size=2)
    assert got["nr_update"] == 11 and got["extra"]["version"] == 11
    assert set(got["master"]) == set(halves[1]) and float(got["master"][halves[1][0]][0]) == 8.0
    assert float(opt.mom1[halves[1][0]][0]) == 2.0
    # re-shard onto one rank: all shards are scanned
    opt1 = Optimizer(0.01)
    got1 = load_optimizer_shards(tmp_path, nlp, opt1, keys, rank=0, world_size=1)
    assert set(opt1.mom1) == set(keys) and set(got1["master"]) == set(keys)
    # same world size but a different ownership map (shard_balance changed): missing keys come from peers' files
    opt2 = Optimizer(0.01)
    load_optimizer_shards(tmp_path, nlp, opt2, halves[0][:1] + halves[1][:1], rank=0, world_size=2)
    assert set(opt2.mom1) == {halves[0][0], halves[1][0]}


# ---------------------------------------------------------------- equal step counts
def test_sharded_batches_give_every_rank_the_same_number_of_steps(tagger_config):
    from spacy_ray_b200.training.batchers import configure_minibatch
    from spacy_ray_b200.training.initialize import init_nlp
    from spacy_ray_b200.training.loop import create_train_batches

    nlp = init_nlp(tagger_config, use_gpu=-1)
    from spacy_ray_b200.config import resolve_dot_names

    cfg = nlp.config.interpolate()
    train_corpus, _dev = resolve_dot_names(cfg, ["corpora.train", "corpora.dev"])
    examples = list(train_corpus(nlp))[:101]                      # 101 docs, batches of 8 -> 13 batches
    batcher = configure_minibatch(size=8)
    seen = []
    for rank in range(3):
        batches = list(create_train_batches(nlp, lambda _n: list(examples), batcher, 2, rank=rank, world_size=3, seed=5))
        seen.append(batches)
    counts = [len(b) for b in seen]
    assert len(set(counts)) == 1 and counts[0] == 2 * (13 // 3), counts
    # disjoint within an epoch
    for ep in (0, 1):
        ids = [sorted(id(eg) for e, b in seen[r] if e == ep for eg in b) for r in range(3)]
        assert not (set(ids[0]) & set(ids[1])) and not (set(ids[1]) & set(ids[2]))


# ---------------------------------------------------------------- async resume
def test_peer_proxy_load_param_overwrites_non_owned_keys_at_a_common_version():
    from spacy_ray_b200.parallel.proxies import PeerProxy

    class Peer:
        pass

    owned, other = (1, "W"), (2, "W")
    p = PeerProxy({owned: None, other: Peer()}, Optimizer(0.01), [owned], grads_per_update=2)
    p.set_param(1, "W", torch.zeros(2))
    p.set_param(2, "W", torch.zeros(2))
    p.set_param(2, "W", torch.ones(2))                       # ignored: not owned and already present
    assert p.get_param(2, "W").tolist() == [0, 0]
    p.load_param(2, "W", torch.full((2,), 5.0), version=2)
    p.load_param(1, "W", torch.full((2,), 6.0), version=2)
    assert p.get_param(2, "W").tolist() == [5, 5] and p.get_param(1, "W").tolist() == [6, 6]
    assert p.check_version(other, 2) and p.check_version(owned, 2)


# ---------------------------------------------------------------- token-balanced batches
def test_trainer_batches_can_be_token_balanced():
    from spacy_ray_b200.engine.trainer import Trainer

    class Store:
        n_docs = 4000
        lens = np.random.default_rng(0).integers(8, 41, size=4000)

    t = Trainer.__new__(Trainer)
    t.store, t.B = Store(), 256
    target = 256 * 24
    bs = t.batches(20, seed=3, tokens_per_batch=target)
    assert all(len(b) == 256 and int(Store.lens[b].sum()) == target for b in bs)
    plain = t.batches(20, seed=3)
    assert len({int(Store.lens[b].sum()) for b in plain}) > 1
    with pytest.raises(ValueError):
        t.batches(1, seed=0, tokens_per_batch=256 * 100)
