"""The asynchronous peer-proxy protocol (reference semantics),
exercised with N proxies wired together in-process."""
import torch

from spacy_ray_b200.parallel.proxies import PeerProxy, RayOptimizer, RayPeerProxy


class Peer:
    """Plain-object stand-in for a worker actor (methods are called directly)."""

    def __init__(self):
        self.proxy = None
        self.dropped = 0

    def inc_grad(self, key, version, value):
        if not self.proxy.receive_grad(key, version, value):
            self.dropped += 1

    def set_param(self, key, version, value):
        self.proxy.receive_param(key, version, value)


def sgd(lr=1.0):
    calls = []

    def opt(key, w, g):
        calls.append(key)
        w -= lr * g
        return w, g

    opt.calls = calls
    return opt


def make_world(n, quorum, keys=((1, "W"), (2, "W"))):
    peers = [Peer() for _ in range(n)]
    owner = {k: i % n for i, k in enumerate(keys)}
    opts = [sgd() for _ in range(n)]
    for r, p in enumerate(peers):
        pm = {k: peers[o] for k, o in owner.items()}
        p.proxy = PeerProxy(pm, opts[r], [k for k, o in owner.items() if o == r], grads_per_update=quorum,
                            all_peers=peers, self_index=r)
        for k in keys:
            p.proxy.set_param(k[0], k[1], torch.zeros(3))
    return peers, owner, opts


def test_aliases():
    assert RayPeerProxy is PeerProxy and RayOptimizer is PeerProxy


def test_install_gives_version_one_everywhere():
    peers, owner, _ = make_world(2, 2)
    for p in peers:
        for k in owner:
            assert p.proxy.check_version(k, 1) is True
            assert p.proxy.check_version(k, 2) is False
        assert p.proxy.check_version((99, "x"), 1) is None


def test_quorum_update_and_param_push_is_staged():
    peers, owner, opts = make_world(2, 2)
    k = (1, "W")                       # owned by rank 0
    p0, p1 = peers[0].proxy, peers[1].proxy
    p0.inc_grad(1, "W", torch.ones(3))           # owner's own gradient: count 1 < quorum
    assert torch.equal(p0.get_param(1, "W"), torch.zeros(3)) and opts[0].calls == []
    p1.inc_grad(1, "W", torch.ones(3) * 2)       # pushed to the owner, same version -> accepted
    w = p0.get_param(1, "W")                     # lazily steps at the next read
    assert opts[0].calls == [k] and torch.allclose(w, torch.full((3,), -3.0))
    assert p0.check_version(k, 2)
    # rank 1 has the new value staged, not adopted: its version is still 1 until it reads
    assert p1.check_version(k, 1) is True
    assert torch.allclose(p1.get_param(1, "W"), torch.full((3,), -3.0))
    assert p1.check_version(k, 2) is True


def test_stale_gradients_are_dropped_and_counted():
    peers, owner, opts = make_world(3, 2)
    k = (1, "W")
    p0, p1, p2 = (p.proxy for p in peers)
    p0.inc_grad(1, "W", torch.ones(3))
    p1.inc_grad(1, "W", torch.ones(3))
    p0.get_param(1, "W")                          # owner steps: version 2
    p2.inc_grad(1, "W", torch.ones(3))            # rank 2 still at version 1 -> stale
    assert peers[0].dropped == 1 and p0.n_grads_discarded == 1
    assert p0.percent_grads_used == 2 / 3


def test_synchronous_special_case_quorum_equals_world():
    n = 4
    peers, owner, opts = make_world(n, n, keys=((1, "W"),))
    for p in peers:
        p.proxy.inc_grad(1, "W", torch.ones(3))
    w = peers[0].proxy.get_param(1, "W")
    assert torch.allclose(w, torch.full((3,), -float(n)))
    for p in peers[1:]:
        assert torch.allclose(p.proxy.get_param(1, "W"), w)
    assert peers[0].proxy.n_grads_discarded == 0


def test_non_owner_never_steps_and_set_grad_is_owner_only():
    peers, owner, opts = make_world(2, 1)
    p1 = peers[1].proxy
    p1.set_grad(1, "W", torch.ones(3))            # not the owner: ignored
    p1.get_param(1, "W")
    assert opts[1].calls == []
    peers[0].proxy.set_grad(1, "W", torch.ones(3))
    peers[0].proxy.get_param(1, "W")
    assert opts[0].calls == [(1, "W")]
