"""GPU arc-eager kernel (parser_kernels.cu) vs the host reference loop, and end-to-end parser /
multi-task training on the sm_100a backend."""
import random

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _projective(n, rng):
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


@pytest.mark.parametrize("nO,nP,n_labels", [(64, 2, 5), (128, 2, 9), (64, 3, 20)])
def test_arc_eager_kernel_matches_reference_loop(nO, nP, n_labels):
    from spacy_ray_b200.models.transition_model import TransitionGold, _arc_steps_reference, transition_backward
    from spacy_ray_b200.models.transitions import ArcEagerSystem
    from spacy_ray_b200.nn.batch import make_token_batch
    from spacy_ray_b200.ops.b200_ops import B200Ops
    from spacy_ray_b200.ops.torch_ops import TorchOps

    ops, ref = B200Ops("cuda:0"), TorchOps("cuda:0", dtype=torch.float32)
    rng = random.Random(0)
    torch.manual_seed(0)
    system = ArcEagerSystem([f"d{i}" for i in range(n_labels)])
    lens = [rng.randint(1, 40) for _ in range(53)]
    batch = make_token_batch([np.ones((n, 4), dtype=np.uint64) for n in lens], "cuda:0")
    heads = [_projective(n, rng) for n in lens]
    labels = [[rng.randrange(n_labels) if h != t else -1 for t, h in enumerate(hs)] for hs in heads]
    gold = TransitionGold(heads=heads, labels=labels)
    nF = 8
    Tp = batch.n_rows
    Yf = (torch.randn(Tp, nF * nO * nP, device="cuda") * batch.mask).bfloat16()
    params = {
        "pad": (torch.randn(nF, nO * nP, device="cuda") * 0.3).bfloat16(),
        "b": (torch.randn(nO * nP, device="cuda") * 0.3).bfloat16(),
        "Wu": (torch.randn(system.n_actions, nO, device="cuda") * 0.3).bfloat16(),
        "bu": (torch.randn(system.n_actions, device="cuda") * 0.1).bfloat16(),
        "nF": nF, "nO": nO, "nP": nP,
    }
    rec = ops.transition_steps(system, Yf, params, batch, gold, True)
    assert rec is not None and "arc_heads" in rec
    pf = {k: (v.float() if torch.is_tensor(v) else v) for k, v in params.items()}
    rref = _arc_steps_reference(system, Yf.float(), pf, batch, gold, True)
    torch.cuda.synchronize()
    hist = rec["arc_history"].cpu().tolist()
    nsteps = rec["arc_n_steps"].cpu().tolist()
    heads_k = rec["arc_heads"].cpu().tolist()
    same_docs, pos, tok = 0, 0, 0
    for d, n in enumerate(lens):
        mine = hist[2 * tok: 2 * tok + nsteps[d]]
        theirs = rref["histories"][d]
        assert len(mine) == nsteps[d] and 1 <= nsteps[d] <= 2 * n
        if mine == theirs:
            same_docs += 1
            want_heads, _ = system.finalize(rref["states"][d])
            assert heads_k[tok:tok + n] == want_heads
        # every derivation must be a legal one: replay it through the host system
        s = system.init_state(n)
        for a in mine:
            assert system.valid(s)[a], (d, a)
            system.apply(s, a)
        assert s.is_final
        tok += n
    assert same_docs / len(lens) > 0.7, same_docs       # bf16 vs fp32 near-ties diverge a few docs
    d = rec["d_scores"].float()[:, : system.n_actions]
    assert float(d.sum(dim=1).abs().max()) < 2e-2       # each row: softmax(valid) - softmax(gold) sums to 0
    assert abs(float(rec["loss"]) - float(rref["loss"])) / max(float(rref["loss"]), 1e-6) < 0.3
    g = ops.transition_backward(rec, params, Tp)
    rec_f = {"d_scores": d, "hid": rec["hid"].float(), "which": rec["which"], "feats": rec["feats"].long()}
    gr = transition_backward(ref, rec_f, pf, Tp)
    for k in ("dYf", "dpad", "db", "dWu"):
        err = (g[k].float() - gr[k].float()).abs()
        assert float((err > 0.05 + 0.03 * gr[k].float().abs()).float().mean()) < 1e-3, k


def _train_gpu(pipeline, steps, bs=64, width=64, hidden=64):
    from conftest import multi_cfg
    from spacy_ray_b200.config import Config
    from spacy_ray_b200.worker import Worker

    cfg = Config().from_str(multi_cfg(pipeline, width=width, depth=2, n_docs=400, max_len=16, hidden=hidden),
                            interpolate=False)
    w = Worker(cfg, rank=0, num_workers=1, use_gpu=0, mode="sync", comm="auto")
    w.set_proxy(None)
    nlp = w.nlp
    exs = list(w.train_corpus(nlp))
    hist = []
    for step in range(steps):
        losses = {}
        lo = (step * bs) % (len(exs) - bs)
        nlp.update(exs[lo:lo + bs], drop=0.1, sgd=False, losses=losses)
        w.proxy.step()
        hist.append({k: float(v) for k, v in losses.items()})
    w.proxy.comm.check()
    return nlp, exs, hist


def test_parser_trains_on_gpu_with_device_side_derivations():
    nlp, exs, hist = _train_gpu(["parser"], steps=150)
    first = sum(h["parser"] for h in hist[:5])
    last = sum(h["parser"] for h in hist[-5:])
    assert last < first, (first, last)
    scores = nlp.evaluate(exs[:100])
    assert scores["dep_uas"] > 0.5, scores


def test_multitask_tagger_parser_ner_on_gpu():
    nlp, exs, hist = _train_gpu(["tagger", "parser", "ner"], steps=60)
    assert hist[-1]["tagger"] < hist[0]["tagger"]
    scores = nlp.evaluate(exs[:60])
    assert scores["tag_acc"] > 0.3 and scores["dep_uas"] is not None and scores["ents_f"] is not None


def _shared_tok2vec_cfg(n_docs=300):
    from pathlib import Path

    text = (Path(__file__).resolve().parent.parent / "configs" / "multitask_w512.cfg").read_text()
    text = text.replace("width = 512", "width = 64").replace("depth = 8", "depth = 2")
    text = text.replace("hidden_width = 128", "hidden_width = 64").replace("n_docs = 20000", f"n_docs = {n_docs}")
    return text.replace("max_len = 40", "max_len = 16")


@pytest.mark.parametrize("shared", [False, True])
def test_engine_serves_multi_head_pipelines_like_the_generic_path(shared):
    """engine.Trainer (packed staging buffer -> CUDA-graph step) must train a tagger+parser+ner
    pipeline - with per-head tok2vecs and with one shared tok2vec + listeners - to the same losses
    as the generic nlp.update path on the same batches."""
    from conftest import multi_cfg
    from spacy_ray_b200.config import Config
    from spacy_ray_b200.engine import Trainer
    from spacy_ray_b200.nn.layers import fix_random_seed
    from spacy_ray_b200.worker import Worker

    text = _shared_tok2vec_cfg() if shared else multi_cfg(["tagger", "parser", "ner"], width=64, depth=2, n_docs=300,
                                                          max_len=16, hidden=64)

    def make():
        fix_random_seed(0)
        w = Worker(Config().from_str(text, interpolate=False), rank=0, num_workers=1, use_gpu=0, mode="sync",
                   comm="auto")
        w.set_proxy(None)
        return w, list(w.train_corpus(w.nlp))

    bs, steps = 32, 12
    w_gen, exs = make()
    generic = []
    for step in range(steps):
        losses = {}
        w_gen.nlp.update(exs[step * 8: step * 8 + bs], drop=0.0, sgd=False, losses=losses)
        w_gen.proxy.step()
        generic.append({k: float(v) for k, v in losses.items()})
    w_fast, exs2 = make()
    trainer = Trainer(w_fast.nlp, w_fast.proxy, exs2, docs_per_batch=bs, dropout=0.0, prefetch=False)
    assert trainer.loss_names == ["tagger", "parser", "ner"]
    fast = []
    for step in range(steps):
        ids = np.arange(step * 8, step * 8 + bs)
        trainer.prepare(ids)
        vec = trainer.step_async()
        fast.append({n: float(v) for n, v in trainer.losses_dict(vec).items()})
    assert len(trainer._graphs) >= 1
    trainer.close()
    for g, f in zip(generic, fast):
        for head in ("tagger", "parser", "ner"):
            assert abs(g[head] - f[head]) <= 0.08 * abs(g[head]) + 0.02, (head, generic, fast)
    assert fast[-1]["tagger"] < fast[0]["tagger"]


def test_train_step_reads_each_loss_one_step_late():
    """Trainer.train_step(lag=1) must report exactly the losses a blocking run (lag=0) reports, shifted by
    one call; flush_loss() returns the last one."""
    from conftest import multi_cfg
    from spacy_ray_b200.config import Config
    from spacy_ray_b200.engine import Trainer
    from spacy_ray_b200.nn.layers import fix_random_seed
    from spacy_ray_b200.worker import Worker

    text = multi_cfg(["ner"], width=64, depth=2, n_docs=200, max_len=16, hidden=64)

    def run(lag):
        fix_random_seed(0)
        w = Worker(Config().from_str(text, interpolate=False), rank=0, num_workers=1, use_gpu=0, mode="sync",
                   comm="auto")
        w.set_proxy(None)
        exs = list(w.train_corpus(w.nlp))
        tr = Trainer(w.nlp, w.proxy, exs, docs_per_batch=32, dropout=0.0, prefetch=False)
        out = [tr.train_step(np.arange(i * 8, i * 8 + 32), lag=lag) for i in range(6)]
        last = tr.flush_loss()
        tr.close()
        return out, last

    blocking, last0 = run(0)
    lagged, last1 = run(1)
    assert last0 is None or last0 == pytest.approx(blocking[-1])
    assert lagged[0] == pytest.approx(blocking[0], rel=3e-2)            # first call has nothing older to report
    for i in range(1, 6):
        assert lagged[i] == pytest.approx(blocking[i - 1], rel=3e-2), (i, lagged, blocking)
    assert last1 == pytest.approx(blocking[-1], rel=3e-2)


def test_prepared_batches_are_consumed_in_submission_order_with_two_collate_workers():
    """Two collate workers may finish out of order (a 4-doc batch behind a 64-doc one); ``step_async`` must
    still run the batches in the order ``prepare`` was called, and preparing past the staging ring must raise."""
    from conftest import multi_cfg
    from spacy_ray_b200.config import Config
    from spacy_ray_b200.engine import Trainer
    from spacy_ray_b200.worker import Worker

    text = multi_cfg(["ner"], width=64, depth=1, n_docs=256, max_len=16, hidden=64)
    w = Worker(Config().from_str(text, interpolate=False), rank=0, num_workers=1, use_gpu=0, mode="sync", comm="auto")
    w.set_proxy(None)
    exs = list(w.train_corpus(w.nlp))
    tr = Trainer(w.nlp, w.proxy, exs, docs_per_batch=64, dropout=0.0, prefetch=True, prefetch_workers=2)
    assert len(tr._threads) == 2 and len(tr.stages) >= 4
    sizes = [64, 4, 48, 2, 64, 8, 1, 33]
    i = 0
    while i < len(sizes):
        pair = sizes[i:i + 2]
        for n in pair:
            tr.prepare(np.arange(n, dtype=np.int64))
        for n in pair:
            tr.step_async()
            assert tr.last["docs"] == n, (sizes, i, tr.last)
        i += 2
    for n in (3, 5, 7):
        tr.prepare(np.arange(n, dtype=np.int64))
    with pytest.raises(RuntimeError):
        tr.prepare(np.arange(2, dtype=np.int64))
    for n in (3, 5, 7):
        tr.step_async()
        assert tr.last["docs"] == n
    torch.cuda.synchronize()
    tr.close()


def test_learning_rate_changes_reach_the_captured_step():
    """The fused exchange kernel reads its hyper-parameters from a device tensor; a schedule that
    changes the learning rate between steps must take effect on CUDA-graph replays too."""
    from conftest import multi_cfg
    from spacy_ray_b200.config import Config
    from spacy_ray_b200.engine import Trainer
    from spacy_ray_b200.worker import Worker

    text = multi_cfg(["ner"], width=64, depth=2, n_docs=200, max_len=16, hidden=64)
    w = Worker(Config().from_str(text, interpolate=False), rank=0, num_workers=1, use_gpu=0, mode="sync", comm="auto")
    w.set_proxy(None)
    exs = list(w.train_corpus(w.nlp))
    tr = Trainer(w.nlp, w.proxy, exs, docs_per_batch=32, dropout=0.0, prefetch=False)
    ids = np.arange(32)
    for _ in range(4):                                   # eager first step, capture, replays
        tr.train_step(ids, lag=0)
    assert len(tr._graphs) >= 1
    before = w.proxy.param_flat.clone()
    tr.train_step(ids, lag=0)
    torch.cuda.synchronize()
    assert not torch.equal(before, w.proxy.param_flat)   # still learning
    w.optimizer.learn_rate = 0.0
    w.optimizer.L2 = 0.0
    frozen = w.proxy.param_flat.clone()
    for _ in range(3):
        tr.train_step(ids, lag=0)
    torch.cuda.synchronize()
    assert torch.equal(frozen, w.proxy.param_flat)       # lr = 0 reached the replayed kernel
    nr = list(w.optimizer.nr_update.values())
    assert nr and min(nr) >= 8                           # host-side update counters follow the replays
    tr.close()


@pytest.mark.parametrize("nO,nP,n_labels", [(64, 2, 5), (64, 3, 20)])
def test_arc_eager_kernel_teacher_forced_matches_reference_tightly(nO, nP, n_labels):
    """With teacher forcing both implementations follow the oracle's derivation, so the step records
    line up one to one and loss / d_scores can be compared at bf16 resolution instead of 30 %."""
    from spacy_ray_b200.models.transition_model import TransitionGold, _arc_steps_reference
    from spacy_ray_b200.models.transitions import ArcEagerSystem
    from spacy_ray_b200.nn.batch import make_token_batch
    from spacy_ray_b200.ops.b200_ops import B200Ops

    ops = B200Ops("cuda:0")
    rng = random.Random(1)
    torch.manual_seed(1)
    system = ArcEagerSystem([f"d{i}" for i in range(n_labels)])
    lens = [rng.randint(1, 40) for _ in range(41)]
    batch = make_token_batch([np.ones((n, 4), dtype=np.uint64) for n in lens], "cuda:0")
    heads = [_projective(n, rng) for n in lens]
    labels = [[rng.randrange(n_labels) if h != t else -1 for t, h in enumerate(hs)] for hs in heads]
    gold = TransitionGold(heads=heads, labels=labels, teacher_forced=True)
    nF, Tp = 8, batch.n_rows
    Yf = (torch.randn(Tp, nF * nO * nP, device="cuda") * batch.mask).bfloat16()
    params = {
        "pad": (torch.randn(nF, nO * nP, device="cuda") * 0.3).bfloat16(),
        "b": (torch.randn(nO * nP, device="cuda") * 0.3).bfloat16(),
        "Wu": (torch.randn(system.n_actions, nO, device="cuda") * 0.3).bfloat16(),
        "bu": (torch.randn(system.n_actions, device="cuda") * 0.1).bfloat16(),
        "nF": nF, "nO": nO, "nP": nP,
    }
    rec = ops.transition_steps(system, Yf, params, batch, gold, True)
    pf = {k: (v.float() if torch.is_tensor(v) else v) for k, v in params.items()}
    rref = _arc_steps_reference(system, Yf.float(), pf, batch, gold, True)
    torch.cuda.synchronize()
    hist = rec["arc_history"].cpu().tolist()
    nsteps = rec["arc_n_steps"].cpu().tolist()
    tok = 0
    # identical derivations, and they reach the gold trees
    for d, n in enumerate(lens):
        assert hist[2 * tok: 2 * tok + nsteps[d]] == rref["histories"][d], d
        want_heads, _ = system.finalize(rref["states"][d])
        assert want_heads == heads[d]
        tok += n
    assert abs(float(rec["loss"]) - float(rref["loss"])) <= 1e-2 * max(float(rref["loss"]), 1e-6)
    # step records: the reference is step-major over live docs, the kernel doc-major; line them up
    A = system.n_actions
    dk = rec["d_scores"].float()[:, :A].cpu()
    order, tok = [], 0
    step_of = []
    for d, n in enumerate(lens):
        step_of.append([2 * tok + s for s in range(nsteps[d])])
        tok += n
    k = 0
    while True:
        live = [d for d in range(len(lens)) if k < nsteps[d]]
        if not live:
            break
        order.extend(step_of[d][k] for d in live)
        k += 1
    dr = rref["d_scores"].cpu()
    assert dr.shape[0] == len(order)
    err = (dk[order] - dr).abs()
    assert float(err.max()) < 1e-2 * max(float(dr.abs().max()), 1e-3) + 2e-3, float(err.max())


def test_arc_eager_kernel_handles_docs_longer_than_128_tokens():
    """The per-warp parser state is sized at launch for the longest doc of the batch (round 1 fell
    back to the host loop beyond 128 tokens)."""
    from spacy_ray_b200.models.transition_model import TransitionGold, _arc_steps_reference
    from spacy_ray_b200.models.transitions import ArcEagerSystem
    from spacy_ray_b200.nn.batch import make_token_batch
    from spacy_ray_b200.ops.b200_ops import B200Ops

    ops = B200Ops("cuda:0")
    rng = random.Random(2)
    torch.manual_seed(2)
    n_labels, nO, nP, nF = 7, 64, 2, 8
    system = ArcEagerSystem([f"d{i}" for i in range(n_labels)])
    assert ops.arc_eager_capacity(nO, nP, system.n_actions) >= 512
    lens = [300, 5, 129, 511, 40, 1, 257]
    batch = make_token_batch([np.ones((n, 4), dtype=np.uint64) for n in lens], "cuda:0")
    heads = [_projective(n, rng) for n in lens]
    labels = [[rng.randrange(n_labels) if h != t else -1 for t, h in enumerate(hs)] for hs in heads]
    gold = TransitionGold(heads=heads, labels=labels, teacher_forced=True)
    Tp = batch.n_rows
    Yf = (torch.randn(Tp, nF * nO * nP, device="cuda") * batch.mask).bfloat16()
    params = {
        "pad": (torch.randn(nF, nO * nP, device="cuda") * 0.3).bfloat16(),
        "b": (torch.randn(nO * nP, device="cuda") * 0.3).bfloat16(),
        "Wu": (torch.randn(system.n_actions, nO, device="cuda") * 0.3).bfloat16(),
        "bu": (torch.randn(system.n_actions, device="cuda") * 0.1).bfloat16(),
        "nF": nF, "nO": nO, "nP": nP,
    }
    rec = ops.transition_steps(system, Yf, params, batch, gold, True)
    assert rec is not None and "arc_heads" in rec, "long docs must stay on the device kernel"
    pf = {k: (v.float() if torch.is_tensor(v) else v) for k, v in params.items()}
    rref = _arc_steps_reference(system, Yf.float(), pf, batch, gold, True)
    torch.cuda.synchronize()
    got = rec["arc_heads"].cpu().tolist()
    tok = 0
    for d, n in enumerate(lens):
        assert got[tok:tok + n] == heads[d], d           # teacher forced: the gold tree comes out
        tok += n
    assert abs(float(rec["loss"]) - float(rref["loss"])) <= 1e-2 * max(float(rref["loss"]), 1e-6)
