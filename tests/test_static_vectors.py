"""Static vectors (``include_static_vectors = true``): table loading, row lookup, the projection layer
and a pipeline that trains with them and round-trips through disk."""
import numpy as np
import torch

from spacy_ray_b200.nn.staticvectors import Vectors, get_vectors, set_vectors
from spacy_ray_b200.training.docbin import hash_string


def _table(words, dim=12, seed=0):
    rng = np.random.default_rng(seed)
    return Vectors(rng.standard_normal((len(words), dim)).astype(np.float32), words=words)


def test_vectors_lookup_formats_roundtrip(tmp_path):
    v = _table(["apple", "pear", "kiwi"])
    assert v.rows_for(["pear", "zzz", "apple"]).tolist() == [1, -1, 0]
    assert v.key2row[hash_string("kiwi")] == 2
    v.to_disk(tmp_path / "v.npz")
    v2 = Vectors.from_disk(tmp_path / "v.npz")
    assert np.array_equal(v2.data, v.data) and v2.rows_for(["kiwi"]).tolist() == [2]
    txt = tmp_path / "v.vec"
    txt.write_text("3 12\n" + "\n".join(w + " " + " ".join(f"{x:.6f}" for x in row) for w, row in zip(["apple", "pear", "kiwi"], v.data)))
    v3 = Vectors.from_disk(txt)
    assert v3.shape == (3, 12) and np.allclose(v3.data, v.data, atol=1e-5)
    np.savez(tmp_path / "w.npz", words=np.array(["apple", "pear", "kiwi"]), data=v.data)
    assert Vectors.from_disk(tmp_path / "w.npz").rows_for(["pear"]).tolist() == [1]


def test_static_vectors_layer_forward_backward_matches_autograd():
    from spacy_ray_b200.nn.batch import make_token_batch
    from spacy_ray_b200.nn.staticvectors import StaticVectors

    words = ["a", "b", "c", "d"]
    set_vectors(_table(words, dim=6))
    try:
        layer = StaticVectors(5).initialize()
        batch = make_token_batch([np.ones((3, 4), dtype=np.uint64), np.ones((2, 4), dtype=np.uint64)], "cpu")
        rows = np.full((batch.n_rows,), -1, dtype=np.int64)
        rows[1:4] = [0, 3, -1]
        rows[5:7] = [2, 2]
        batch.extra["vec_rows"] = torch.from_numpy(rows)
        Y, bp = layer(batch, True)
        W = layer.get_param("W").clone().requires_grad_(True)
        T = torch.from_numpy(get_vectors().data)
        ok = torch.from_numpy(rows >= 0).float().unsqueeze(1)
        want = (T[torch.from_numpy(rows).clamp(min=0)] * ok) @ W.t()
        assert torch.allclose(Y, want.detach(), atol=1e-5)
        assert float(Y[0].abs().sum()) == 0 and float(Y[3].abs().sum()) == 0      # pad row, OOV token
        dY = torch.randn_like(Y)
        want.backward(dY)

        got = {}

        class Proxy:
            def get_param(self, i, n): return layer._params._params[(i, n)]
            def set_param(self, i, n, v): pass
            def inc_grad(self, i, n, v): got[n] = v
            def set_grad(self, i, n, v): got[n] = v

        layer._params.proxy = Proxy()
        bp(dY)
        assert torch.allclose(got["W"].float(), W.grad, atol=1e-4)
    finally:
        set_vectors(None)


def test_pipeline_trains_with_static_vectors_and_roundtrips(tmp_path):
    from conftest import multi_cfg
    from spacy_ray_b200.config import Config
    from spacy_ray_b200.pipeline.language import load
    from spacy_ray_b200.training.corpus import SyntheticCorpus
    from spacy_ray_b200.training.initialize import init_nlp

    corpus = SyntheticCorpus(60, seed=1, min_len=4, max_len=10, vocab_size=300, tasks=("tagger",))
    vocab = sorted({w for d in corpus.docs() for w in d.words})
    _table(vocab, dim=16, seed=3).to_disk(tmp_path / "vectors.npz")
    text = multi_cfg(["tagger"], width=32, depth=1, n_docs=60, max_len=10)
    text = text.replace("pretrained_vectors = null", "pretrained_vectors = true")
    text += f'\n[initialize]\nvectors = "{tmp_path / "vectors.npz"}"\n'
    try:
        nlp = init_nlp(Config().from_str(text, interpolate=False), use_gpu=-1)
        model = nlp.get_pipe("tagger").model
        assert any(n.name == "staticvectors" for n in model.walk())
        mix = next(n for n in model.walk() if n.name == "maxout")
        assert mix.get_dim("nI") == 32 * 5                       # 4 hash embeds + the vectors projection
        from spacy_ray_b200.config import resolve_dot_names

        train, _ = resolve_dot_names(nlp.config.interpolate(), ["corpora.train", "corpora.dev"])
        exs = list(train(nlp))
        opt = nlp.create_optimizer()
        first = last = None
        for i in range(25):
            losses = {}
            nlp.update(exs[:32], drop=0.0, sgd=opt, losses=losses)
            first = float(losses["tagger"]) if first is None else first
            last = float(losses["tagger"])
        assert last < 0.85 * first, (first, last)
        sv = next(n for n in model.walk() if n.name == "staticvectors")
        assert float(sv.get_param("W").abs().sum()) > 0
        before = nlp.evaluate(exs[:20])["tag_acc"]
        nlp.to_disk(tmp_path / "model")
        assert (tmp_path / "model" / "vocab" / "vectors.npz").exists()
        set_vectors(None)
        nlp2 = load(tmp_path / "model")
        assert get_vectors() is not None
        assert abs(nlp2.evaluate(exs[:20])["tag_acc"] - before) < 1e-6
    finally:
        set_vectors(None)
