import torch

from spacy_ray_b200.nn import Model, reset_model_ids
from spacy_ray_b200.nn.layers import HashEmbedCNN, Maxout
from spacy_ray_b200.parallel.util import set_params_proxy


class RecordingProxy:
    def __init__(self):
        self.params, self.grads, self.calls = {}, {}, []

    def set_param(self, id, name, value):
        self.params[(id, name)] = value
        self.calls.append(("set_param", id, name))

    def get_param(self, id, name):
        self.calls.append(("get_param", id, name))
        return self.params[(id, name)]

    def inc_grad(self, id, name, value):
        self.calls.append(("inc_grad", id, name))
        self.grads[(id, name)] = self.grads.get((id, name), 0) + value

    def set_grad(self, id, name, value):
        self.grads[(id, name)] = value


def test_ids_are_deterministic_after_reset():
    reset_model_ids()
    a = [n.id for n in HashEmbedCNN(32, 2, 300).walk()]
    reset_model_ids()
    b = [n.id for n in HashEmbedCNN(32, 2, 300).walk()]
    assert a == b and len(set(a)) == len(a)


def test_walk_is_breadth_first_and_unique():
    m = HashEmbedCNN(32, 2, 300)
    nodes = list(m.walk())
    assert nodes[0] is m
    assert len({id(n) for n in nodes}) == len(nodes)
    depth = {id(m): 0}
    for n in nodes:
        for c in n.layers:
            depth.setdefault(id(c), depth[id(n)] + 1)
    assert [depth[id(n)] for n in nodes] == sorted(depth[id(n)] for n in nodes)


def test_paramserver_proxy_hook():
    m = Maxout(4, 6, 2).initialize()
    X = torch.randn(5, 6)
    proxy = RecordingProxy()
    set_params_proxy(m, proxy)
    assert set(proxy.params) == {(m.id, "W"), (m.id, "b")}
    Y, bp = m.begin_update(X)
    assert ("get_param", m.id, "W") in proxy.calls
    bp(torch.ones_like(Y))
    # gradients go to the proxy only: nothing stored locally, finish_update is a no-op
    assert not m.has_grad("W") and (m.id, "W") in proxy.grads
    called = []
    m.finish_update(lambda k, w, g: called.append(k) or (w, g))
    assert called == []


def test_finish_update_without_proxy():
    m = Maxout(4, 6, 2).initialize()
    Y, bp = m.begin_update(torch.randn(3, 6))
    bp(torch.ones_like(Y))
    assert m.has_grad("W")
    seen = []
    m.finish_update(lambda k, w, g: seen.append(k) or (w, g))
    assert sorted(seen) == sorted([(m.id, "W"), (m.id, "b")]) and not m.has_grad("W")


def test_to_from_bytes_roundtrip():
    reset_model_ids()
    a = HashEmbedCNN(32, 2, 300).initialize()
    b = HashEmbedCNN(32, 2, 300).initialize()
    b.from_bytes(a.to_bytes())
    for na, nb in zip(a.walk(), b.walk()):
        for p in na.param_names:
            assert torch.equal(na.get_param(p), nb.get_param(p))


def test_multi_hash_embed_rejects_attrs_the_featuriser_does_not_produce():
    import pytest

    from spacy_ray_b200.nn.layers import MultiHashEmbed

    MultiHashEmbed(32, attrs=["LOWER", "SHAPE"], rows=[100, 50])
    with pytest.raises(ValueError, match="unsupported attrs"):
        MultiHashEmbed(32, attrs=["NORM", "ORTH"], rows=[100, 50])
