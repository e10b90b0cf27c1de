"""The bucketed exchange kernels (``bucket_reduce_kernel`` / ``bucket_update_kernel``) on ONE GPU: with a single rank the
reduce-scatter is the identity, so every launch is the sharded optimizer - checked per key against
``training.optimizer.Optimizer`` (thinc semantics) on the fp32 reference backend.  The multi-GPU
forms of the same checks are in ``test_multigpu.py``."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _layout(sizes, world=1):
    from spacy_ray_b200.parallel.sync_proxy import ALIGN, FlatLayout, _round_up

    keys = [(i + 1, "E" if i >= len(sizes) - 2 else "W") for i in range(len(sizes))]
    owner, offset, numel, shape = {}, {}, {}, {}
    pos = 0
    for k, n in zip(keys, sizes):
        owner[k], numel[k], shape[k], offset[k] = 0, n, (n,), pos
        pos += _round_up(n, ALIGN)
    cap = _round_up(pos, ALIGN)
    return FlatLayout(keys, owner, offset, numel, shape, [0], [pos], cap, world)


def _mk(opt, sizes, n_buckets):
    from spacy_ray_b200.parallel.fused_comm import FusedSymmComm
    from spacy_ray_b200.parallel.sync_proxy import ShardedSyncProxy

    dev = torch.device("cuda:0")
    layout = _layout(sizes)
    comm = FusedSymmComm(0, 1, layout, dev, optimizer=opt, n_buckets=n_buckets)
    proxy = ShardedSyncProxy(layout, opt, rank=0, world_size=1, device=dev, comm=comm,
                             param_dtype=torch.bfloat16, buffers=comm.buffers)
    comm.bind(proxy)
    g = torch.Generator().manual_seed(0)
    for k in layout.keys:
        o, n = layout.offset[k], layout.numel[k]
        w = (torch.randn(n, generator=g) * 0.3).to(dev)
        proxy.param_flat[o:o + n] = w.bfloat16()
        proxy.master[o:o + n] = proxy.param_flat[o:o + n].float()
    return layout, comm, proxy


SIZES = [128 * 3, 4096 + 128, 77, 128 * 70, 5000, 4096 * 3 + 5, 640, 9000]


@pytest.mark.parametrize("kind", ["adam", "adam_l2grad", "radam", "sgd", "adam_avg"])
@pytest.mark.parametrize("n_buckets", [1, 4])
def test_fused_bucket_kernel_matches_optimizer_per_key(kind, n_buckets):
    from spacy_ray_b200.ops.torch_ops import TorchOps
    from spacy_ray_b200.training.optimizer import Optimizer

    def make():
        kw = dict(L2=0.01, grad_clip=1.0)
        if kind == "adam_l2grad":
            return Optimizer(0.01, L2_is_weight_decay=False, **kw)
        if kind == "radam":
            return Optimizer(0.01, use_radam=True, **kw)
        if kind == "sgd":
            return Optimizer(0.05, use_adam=False, **kw)
        if kind == "adam_avg":
            return Optimizer(0.01, use_averages=True, **kw)
        return Optimizer(0.01, **kw)

    opt = make()
    layout, comm, proxy = _mk(opt, SIZES, n_buckets)
    ref_opt = make()
    ref_opt.ops = TorchOps("cuda:0", dtype=torch.float32)
    ref_w = {k: proxy.master[layout.offset[k]:layout.offset[k] + layout.numel[k]].clone() for k in layout.keys}
    gen = torch.Generator().manual_seed(1)
    for step in range(8):                      # RAdam rectification switches on at step 6 (beta2 = 0.999)
        grads = {}
        for i, k in enumerate(layout.keys):
            n = layout.numel[k]
            # key 0: norm far above the clip threshold, key 2: far below, key 4: zero gradient
            scale = 3.0 if i == 0 else (1e-3 if i == 2 else (0.0 if i == 4 else 0.05))
            grads[k] = (torch.randn(n, generator=gen) * scale).cuda()
        proxy.begin_step(overlap=True)
        for k in reversed(layout.keys):        # the "backward pass": gradients arrive key by key
            proxy.inc_grad(k[0], k[1], grads[k])
        proxy.step()
        torch.cuda.synchronize()
        comm.check()
        for k in layout.keys:
            ref_opt(k, ref_w[k], grads[k].clone())
    assert comm.plan.n >= (3 if n_buckets == 4 else 1)
    assert comm.launches == 8 * sum(comm.kernels_for(b) for b in range(comm.plan.n))
    assert float(proxy.grad_flat.abs().sum()) == 0.0, "gradient buffer not cleared"
    for k in layout.keys:
        o, n = layout.offset[k], layout.numel[k]
        got = proxy.master[o:o + n]
        want = ref_w[k]
        err = (got - want).abs().max().item()
        assert err <= 2e-6 + 2e-5 * want.abs().max().item(), (kind, k, err)
        pb = proxy.param_flat[o:o + n].float()
        assert torch.equal(pb, got.bfloat16().float()), "published bf16 weights != rounded master"
        if kind != "sgd":
            assert torch.allclose(opt.mom1[k].reshape(-1), ref_opt.mom1[k], rtol=1e-4, atol=1e-7)
            assert torch.allclose(opt.mom2[k].reshape(-1), ref_opt.mom2[k], rtol=1e-4, atol=1e-9)
        if kind == "adam_avg":
            assert torch.allclose(opt.averages[k].reshape(-1), ref_opt.averages[k], rtol=1e-4, atol=1e-6)
    assert int(comm.step_t.item()) == 8 and int(comm.epoch.item()) == 8


def test_bucket_launch_order_is_plan_order_even_if_completion_is_not():
    from spacy_ray_b200.training.optimizer import Optimizer

    opt = Optimizer(0.01)
    layout, comm, proxy = _mk(opt, SIZES, 4)
    g = {k: torch.ones(layout.numel[k], device="cuda") * 0.01 for k in layout.keys}
    for k in reversed(layout.keys):            # step 0 fixes the plan (reverse key order)
        proxy.inc_grad(k[0], k[1], g[k])
    proxy.step()
    first = [ks[0] for ks in comm.plan.buckets]
    proxy.begin_step(overlap=True)
    launched_before = comm.launches
    for k in layout.keys:                      # now complete the LAST bucket first
        proxy.inc_grad(k[0], k[1], g[k])
        if k == layout.keys[0]:
            assert comm.launches == launched_before, "a later bucket must wait for its predecessors"
    proxy.step()
    torch.cuda.synchronize()
    comm.check()
    assert comm.launches == launched_before + sum(comm.kernels_for(b) for b in range(comm.plan.n))
    assert first[0] == layout.keys[-1] or first[0][1] == "W"
    assert float(proxy.grad_flat.abs().sum()) == 0.0


def test_resume_on_the_fused_path_restores_moments_counter_and_master(tmp_path):
    """Train, checkpoint, resume in a fresh Worker: the next losses must follow the uninterrupted
    run (the fp32 atomics of the split-K weight-gradient GEMMs make bit equality impossible; 2e-3)."""
    import numpy as np
    from conftest import multi_cfg
    from spacy_ray_b200.config import Config
    from spacy_ray_b200.engine import Trainer
    from spacy_ray_b200.worker import Worker

    text = multi_cfg(["ner"], width=64, depth=2, n_docs=512, max_len=16, hidden=64)

    def make(resume=None):
        w = Worker(Config().from_str(text, interpolate=False), rank=0, num_workers=1, use_gpu=0,
                   resume_path=resume)
        w.set_proxy(None)
        exs = list(w.train_corpus(w.nlp))
        tr = Trainer(w.nlp, w.proxy, exs, docs_per_batch=64, dropout=0.0, prefetch=False)
        return w, tr

    ids = [np.arange(i * 64, (i + 1) * 64, dtype=np.int64) % 512 for i in range(16)]
    wa, ta = make()
    for b in ids[:10]:
        ta.train_step(b, lag=0)
    wa.save_checkpoint({"epoch": 0, "step": 9, "score": 0.0, "words": 0, "seconds": 0, "losses": {}, "other_scores": {}},
                       tmp_path / "ckpt")
    tail_a = [ta.train_step(b, lag=0) for b in ids[10:14]]
    m1_a = wa.proxy.comm.m1.clone()
    ta.close()

    wb, tb = make(resume=tmp_path / "ckpt")
    comm = wb.proxy.comm
    assert int(comm.step_t.item()) == 10, "device-side update counter not restored"
    assert float(comm.m1.abs().sum()) > 0 and float(comm.m2.abs().sum()) > 0, "moments not restored into the kernel's buffers"
    k0 = wb.proxy.owned_keys()[0]
    assert wb.optimizer.mom1[k0].data_ptr() >= comm.m1.data_ptr(), "optimizer moments detached from the kernel buffers"
    tail_b = [tb.train_step(b, lag=0) for b in ids[10:14]]
    tb.close()
    for a, b in zip(tail_a, tail_b):
        assert abs(a - b) <= 2e-3 * max(abs(a), 1e-6), (tail_a, tail_b)
    assert torch.allclose(comm.m1, m1_a, rtol=5e-2, atol=1e-5)
