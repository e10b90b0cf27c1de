import torch

from conftest import multi_cfg
from spacy_ray_b200.config import Config
from spacy_ray_b200.pipeline import load
from spacy_ray_b200.training import init_nlp


def _train(pipeline, steps=25, bs=16):
    cfg = Config().from_str(multi_cfg(pipeline), interpolate=False)
    nlp = init_nlp(cfg)
    from spacy_ray_b200.config import resolve_dot_names
    (corpus,) = resolve_dot_names(nlp.config.interpolate(), ["corpora.train"])
    exs = list(corpus(nlp))
    opt = nlp.create_optimizer()
    history = []
    for step in range(steps):
        losses = {}
        lo = (step * bs) % (len(exs) - bs)
        nlp.update(exs[lo:lo + bs], drop=0.0, sgd=opt, losses=losses)
        history.append({k: float(v) for k, v in losses.items()})
    return nlp, exs, history


def test_tagger_learns():
    nlp, exs, hist = _train(["tagger"], steps=40)
    assert hist[-1]["tagger"] < 0.6 * hist[0]["tagger"]
    scores = nlp.evaluate(exs[:40])
    assert scores["tag_acc"] > 0.5


def test_ner_loss_goes_down_and_predicts_spans():
    nlp, exs, hist = _train(["ner"], steps=110)
    first = sum(h["ner"] for h in hist[:5])
    last = sum(h["ner"] for h in hist[-5:])
    assert last < first
    scores = nlp.evaluate(exs[:40])
    assert scores["ents_f"] is not None and scores["ents_f"] > 0.5


def test_parser_loss_goes_down():
    nlp, exs, hist = _train(["parser"], steps=130)
    assert sum(h["parser"] for h in hist[-5:]) < sum(h["parser"] for h in hist[:5])
    scores = nlp.evaluate(exs[:30])
    assert scores["dep_uas"] > 0.6 and scores["dep_las"] > 0.5


def test_checkpoint_roundtrip_gives_identical_predictions(tmp_path):
    nlp, exs, _ = _train(["tagger", "ner"], steps=8)
    docs = [eg.reference.copy_unannotated() for eg in exs[:10]]
    before = [(d.tags, d.ents) for d in nlp.pipe(docs)]
    nlp.to_disk(tmp_path / "model")
    assert (tmp_path / "model" / "config.cfg").exists() and (tmp_path / "model" / "ner" / "model").exists()
    nlp2 = load(tmp_path / "model")
    docs2 = [eg.reference.copy_unannotated() for eg in exs[:10]]
    after = [(d.tags, d.ents) for d in nlp2.pipe(docs2)]
    assert before == after


def test_shared_tok2vec_listener_trains_both_heads():
    cfg_text = """
[nlp]
lang = "en"
pipeline = ["tok2vec","tagger","ner"]

[components]

[components.tok2vec]
factory = "tok2vec"

[components.tok2vec.model]
@architectures = "spacy.HashEmbedCNN.v2"
width = 32
depth = 2
embed_size = 300
window_size = 1
maxout_pieces = 3
subword_features = true
pretrained_vectors = null

[components.tagger]
factory = "tagger"

[components.tagger.model]
@architectures = "spacy.Tagger.v2"

[components.tagger.model.tok2vec]
@architectures = "spacy.Tok2VecListener.v1"
width = 32
upstream = "*"

[components.ner]
factory = "ner"

[components.ner.model]
@architectures = "spacy.TransitionBasedParser.v2"
state_type = "ner"
hidden_width = 32
maxout_pieces = 2

[components.ner.model.tok2vec]
@architectures = "spacy.Tok2VecListener.v1"
width = 32
upstream = "*"

[corpora]

[corpora.train]
@readers = "spacy_ray_b200.SyntheticCorpus.v1"
n_docs = 60
seed = 1
max_len = 10
vocab_size = 300

[corpora.dev]
@readers = "spacy_ray_b200.SyntheticCorpus.v1"
n_docs = 20
seed = 2
max_len = 10
vocab_size = 300
"""
    nlp = init_nlp(Config().from_str(cfg_text, interpolate=False))
    t2v = nlp.get_pipe("tok2vec")
    assert len(t2v.listeners) == 2
    from spacy_ray_b200.config import resolve_dot_names
    (corpus,) = resolve_dot_names(nlp.config.interpolate(), ["corpora.train"])
    exs = list(corpus(nlp))
    node = next(n for n in t2v.model.walk() if n.name == "maxout")
    before = node.get_param("W").clone()
    opt = nlp.create_optimizer()
    losses = {}
    for i in range(6):
        nlp.update(exs[i * 8:(i + 1) * 8], sgd=opt, losses=losses)
    assert not torch.equal(before, node.get_param("W"))       # gradients reached the shared layer
    assert float(losses["tagger"]) > 0 and float(losses["ner"]) > 0
    assert nlp.evaluate(exs[:10])["tag_acc"] is not None


def test_shared_tok2vec_deferred_backprop_matches_immediate():
    """Tok2VecComponent.update(defer_backprop=True) + finish_backprop() (what engine.Trainer uses to
    run listener heads on concurrent streams) must produce the same gradients as the default
    'backward fires with the last listener' protocol."""
    from pathlib import Path

    import torch

    from spacy_ray_b200.config import Config, resolve_dot_names
    from spacy_ray_b200.nn.layers import fix_random_seed
    from spacy_ray_b200.training.initialize import init_nlp

    text = (Path(__file__).resolve().parent.parent / "configs" / "multitask_w512.cfg").read_text()
    text = text.replace("width = 512", "width = 32").replace("depth = 8", "depth = 1")
    text = text.replace("hidden_width = 128", "hidden_width = 32").replace("n_docs = 20000", "n_docs = 24")
    text = text.replace("max_len = 40", "max_len = 9")
    cfg = Config().from_str(text, interpolate=False)

    def grads(defer: bool):
        fix_random_seed(0)
        nlp = init_nlp(cfg)
        corpus = resolve_dot_names(cfg.interpolate(), ["corpora.train"])[0]
        examples = list(corpus(nlp))[:16]
        batch = nlp.make_batch([eg.predicted for eg in examples])
        t2v = nlp.get_pipe("tok2vec")
        gen = torch.Generator().manual_seed(1)
        for name in ("tagger", "parser", "ner"):        # output layers start at zero: no gradient would reach tok2vec
            for node in nlp.get_pipe(name).model.walk():
                for pname in node.param_names:
                    w = node.get_param(pname)
                    node.set_param(pname, w + 0.1 * torch.randn(w.shape, generator=gen).to(w.dtype))
        t2v.update(examples, batch=batch, drop=0.0, sgd=False, losses={}, defer_backprop=defer)
        for name in ("tagger", "parser", "ner"):
            nlp.get_pipe(name).update(examples, batch=batch, drop=0.0, sgd=False, losses={})
        if defer:
            t2v.finish_backprop()
        out = {}
        for node in t2v.model.walk():
            for pname in node.param_names:
                g = node.get_grad(pname) if node.has_grad(pname) else None
                if g is not None:
                    out[(node.name, node.id, pname)] = g.clone()
        return out

    a, b = grads(False), grads(True)
    assert a and a.keys() == b.keys()
    assert any(float(v.abs().sum()) > 0 for v in a.values())
    for k in a:
        torch.testing.assert_close(a[k], b[k], rtol=1e-5, atol=1e-6)
