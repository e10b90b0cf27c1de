"""Worker-level tests.  ``test_worker_init`` mirrors the reference's only test
(``spacy_ray/tests/test_worker.py:24-30``): a blank config with empty corpus paths
and an injected mock ``ray``."""
import sys
import types

import pytest
import torch

from spacy_ray_b200 import Language, blank
from spacy_ray_b200.worker import Evaluator, FakeOptimizer, Worker

mock_ray = types.SimpleNamespace(get=lambda *a, **k: None, init=lambda *a, **k: None, remote=lambda *a, **k: None)


def test_worker_init():
    nlp = blank("en")
    nlp.config["paths"]["train"] = ""
    nlp.config["paths"]["dev"] = ""
    worker = Worker(nlp.config, rank=1, num_workers=2, use_gpu=-1, ray=mock_ray)
    assert isinstance(worker.nlp, Language)
    assert worker.ray is mock_ray and worker.get_quorum() == 2
    assert worker.get_percent_grads_used() is None


def test_fake_optimizer_and_evaluator_contracts():
    class Real:
        n = 0

        def step_schedules(self):
            self.n += 1

    real = Real()
    fo = FakeOptimizer(real)
    w, g = torch.ones(2), torch.ones(2)
    assert fo(("k", "W"), w, g) == (w, g) and fo.averages == {}
    fo.step_schedules()
    assert real.n == 1
    ev = Evaluator()
    assert ev.get_scores() is None and ev.get_scores(1) is None
    ev.set_scores({"a": 1}, 1)
    assert ev.get_scores(1) == {"a": 1} and ev.get_scores(2) is None and ev.get_scores() == {"a": 1}


def test_single_worker_trains_and_writes_checkpoints(tagger_config, tmp_path):
    w = Worker(tagger_config, rank=0, num_workers=1, use_gpu=-1, mode="sync", comm="local", output_path=tmp_path)
    w.set_proxy(None)
    owned = w.get_owned_keys()
    assert len(owned) == len(set(owned)) > 0
    w.train(None, None)
    w.join(timeout=120)
    assert not w.is_running() and w.get_error() is None
    assert (tmp_path / "model-best" / "config.cfg").exists() and (tmp_path / "model-last" / "tagger" / "model").exists()
    assert (tmp_path / "model-last" / "optim" / "rank0-of1.pt").exists()
    assert w.get_percent_grads_used() == 1.0


def test_resume_restores_weights_and_optimizer_state(tagger_config, tmp_path):
    w = Worker(tagger_config, rank=0, num_workers=1, use_gpu=-1, comm="local", output_path=tmp_path)
    w.set_proxy(None)
    w.train(None, None)
    w.join(timeout=120)
    w2 = Worker(tagger_config, rank=0, num_workers=1, use_gpu=-1, comm="local", resume_path=tmp_path / "model-last")
    w2.set_proxy(None)
    assert torch.equal(w.proxy.param_flat, w2.proxy.param_flat)
    k = w.get_owned_keys()[0]
    assert torch.allclose(w.optimizer.mom1[k], w2.optimizer.mom1[k]) and w2.optimizer.nr_update[k] == w.optimizer.nr_update[k]


@pytest.mark.slow
def test_two_workers_sync_gloo_end_identical(tagger_config, tmp_path):
    from spacy_ray_b200.train_cli import ray_train

    ray_train(tagger_config, num_workers=2, use_gpu=-1, mode="sync", comm="dist", output_path=tmp_path)
    a = torch.load(tmp_path / "model-last" / "optim" / "rank0-of2.pt", weights_only=False)
    b = torch.load(tmp_path / "model-last" / "optim" / "rank1-of2.pt", weights_only=False)
    assert a["world_size"] == 2 and set(a["mom1"]).isdisjoint(set(b["mom1"]))   # disjoint ownership
    assert (tmp_path / "model-best" / "meta.json").exists()


@pytest.mark.slow
def test_two_workers_async_reference_protocol_runs(tagger_config):
    from spacy_ray_b200.train_cli import ray_train

    ray_train(tagger_config, num_workers=2, use_gpu=-1, mode="async", quorum=2)


@pytest.mark.slow
def test_worker_failure_is_reported_not_hung(tagger_config):
    from spacy_ray_b200.train_cli import ray_train

    with pytest.raises(RuntimeError, match="injected fault"):
        ray_train(tagger_config, num_workers=2, use_gpu=-1, mode="sync", comm="dist", inject_fault="1:3")
