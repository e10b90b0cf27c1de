import torch

from spacy_ray_b200.nn import reset_model_ids
from spacy_ray_b200.nn.layers import HashEmbedCNN
from spacy_ray_b200.parallel.sync_proxy import ALIGN, FlatLayout, ShardedSyncProxy
from spacy_ray_b200.parallel.util import divide_params, set_params_proxy
from spacy_ray_b200.training.optimizer import Optimizer


class FakeGroupComm:
    """All ranks in one process: collectives over a shared list of proxies."""

    def __init__(self, rank, world):
        self.rank, self.world_size, self.world = rank, len(world), world

    def reduce_scatter(self, grad_flat, layout):
        cap = layout.shard_cap
        total = sum(getattr(self, "snapshot", None) or [p.grad_flat for p in self.world])
        return total[self.rank * cap:(self.rank + 1) * cap].clone()

    def all_gather(self, param_flat, layout):
        cap = layout.shard_cap
        for r, p in enumerate(self.world):
            param_flat[r * cap:(r + 1) * cap] = p.param_flat[r * cap:(r + 1) * cap]


def build(n):
    models, proxies = [], []
    for r in range(n):
        reset_model_ids()
        from spacy_ray_b200.nn.layers import fix_random_seed
        fix_random_seed(0)
        m = HashEmbedCNN(32, 2, 300).initialize()
        models.append(m)
    layout = FlatLayout.build([("t2v", models[0])], n)
    for r in range(n):
        opt = Optimizer(0.01, L2=0.0, grad_clip=1.0)
        p = ShardedSyncProxy(layout, opt, rank=r, world_size=n, device="cpu", comm=None)
        proxies.append(p)
    for r in range(n):
        proxies[r].comm = FakeGroupComm(r, proxies)
        set_params_proxy(models[r], proxies[r])
    return models, proxies, layout


def test_layout_is_owner_major_aligned_and_matches_divide_params():
    models, proxies, layout = build(3)
    shares = divide_params(models[0], 3)
    for r in range(3):
        assert layout.owned_keys(r) == shares[r]
        for k in shares[r]:
            assert layout.shard_start[r] <= layout.offset[k] < layout.shard_start[r] + layout.shard_cap
            assert layout.offset[k] % ALIGN == 0
    assert layout.total == 3 * layout.shard_cap


def test_params_are_views_of_the_flat_buffer():
    models, proxies, layout = build(2)
    node = next(n for n in models[0].walk() if n.param_names)
    name = node.param_names[0]
    view = node.get_param(name)
    proxies[0].param_flat.zero_()
    assert float(view.abs().sum()) == 0.0


def test_sync_step_equals_single_process_adam_on_summed_grads():
    n = 2
    models, proxies, layout = build(n)
    reset_model_ids()
    from spacy_ray_b200.nn.layers import fix_random_seed
    fix_random_seed(0)
    ref = HashEmbedCNN(32, 2, 300).initialize()
    ref_opt = Optimizer(0.01, L2=0.0, grad_clip=1.0)
    gen = torch.Generator().manual_seed(1)
    grads = {}
    for node_r, nodes in zip(ref.walk(), zip(*[m.walk() for m in models])):
        for name in node_r.param_names:
            per_rank = [torch.randn(node_r.get_param(name).shape, generator=gen) for _ in range(n)]
            for node, g in zip(nodes, per_rank):
                node.inc_grad(name, g)
            grads[(node_r.id, name)] = sum(per_rank)
    snap = [p.grad_flat.clone() for p in proxies]      # ranks step one after another here
    for p in proxies:
        p.comm.snapshot = snap
        p.step()
    for p in proxies:          # the fake comm runs ranks one after another: re-gather once all have stepped
        p.sync_from_owner()
    for node_r, nodes in zip(ref.walk(), zip(*[m.walk() for m in models])):
        for name in node_r.param_names:
            w = node_r.get_param(name)
            ref_opt((node_r.id, name), w, grads[(node_r.id, name)].clone())
            for node in nodes:
                assert torch.allclose(node.get_param(name), w, atol=1e-6), (node_r.name, name)
    assert all(p.version == 1 for p in proxies)
    assert all(float(p.grad_flat.abs().sum()) == 0.0 for p in proxies)
