import pytest

from spacy_ray_b200.nn import reset_model_ids
from spacy_ray_b200.nn.layers import HashEmbedCNN
from spacy_ray_b200.parallel.util import divide_params, divide_params_balanced, make_key


def _model():
    reset_model_ids()
    return HashEmbedCNN(32, 3, 300).initialize()


@pytest.mark.parametrize("n", [1, 2, 3, 4, 8, 16, 64])
def test_every_key_owned_exactly_once(n):
    m = _model()
    shares = divide_params(m, n)
    assert len(shares) == n
    flat = [k for s in shares for k in s]
    expected = [make_key(node.id, p) for node in m.walk() for p in node.param_names]
    assert sorted(flat) == sorted(expected) and len(set(flat)) == len(flat)


@pytest.mark.parametrize("n", [2, 3, 4])
def test_node_grouping_and_leftovers_to_last_rank(n):
    m = _model()
    shares = divide_params(m, n)
    owner = {k: r for r, s in enumerate(shares) for k in s}
    for node in m.walk():
        assert len({owner[make_key(node.id, p)] for p in node.param_names} or {0}) <= 1
    groups = [node for node in m.walk() if node.param_names]
    per = max(1, len(groups) // n)
    counts = [len({k[0] for k in s}) for s in shares]
    assert counts[:-1] == [per] * (n - 1)
    assert counts[-1] == len(groups) - per * (n - 1)


def test_fewer_groups_than_workers_leaves_trailing_ranks_empty():
    m = _model()
    n_groups = len([node for node in m.walk() if node.param_names])
    shares = divide_params(m, n_groups + 5)
    assert all(len(s) > 0 for s in shares[:n_groups]) and all(len(s) == 0 for s in shares[n_groups:])


def test_balanced_partition_is_complete_and_flatter():
    m = _model()
    sizes = {make_key(n.id, p): n.get_param(p).numel() for n in m.walk() for p in n.param_names}
    a = divide_params(m, 4)
    b = divide_params_balanced(m, 4)
    assert sorted(k for s in b for k in s) == sorted(sizes)
    load = lambda shares: max(sum(sizes[k] for k in s) for s in shares)
    assert load(b) <= load(a)
