"""``senter`` and ``morphologizer`` (token taggers with their own gold / annotation) and the Doc / DocBin
fields behind them (pos, morphs, lemmas, sent_starts, cats)."""
import random

import pytest

from spacy_ray_b200.config import Config
from spacy_ray_b200.pipeline.doc import Doc, Example

CFG = """
[nlp]
lang = "en"
pipeline = ["tok2vec", "senter", "morphologizer"]

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

[components.senter]
factory = "senter"

[components.senter.model]
@architectures = "spacy.Tagger.v2"

[components.senter.model.tok2vec]
@architectures = "spacy.Tok2VecListener.v1"
width = 32
upstream = "*"

[components.morphologizer]
factory = "morphologizer"

[components.morphologizer.model]
@architectures = "spacy.Tagger.v2"

[components.morphologizer.model.tok2vec]
@architectures = "spacy.Tok2VecListener.v1"
width = 32
upstream = "*"

[corpora]

[corpora.train]
@readers = "spacy_ray_b200.SyntheticCorpus.v1"
n_docs = 4

[corpora.dev]
@readers = "spacy_ray_b200.SyntheticCorpus.v1"
n_docs = 4

[training]
max_steps = 1
"""

NOUNS = ["cat", "dog", "tree", "house", "river", "stone"]
VERBS = ["sees", "likes", "finds", "paints"]
DETS = ["the", "a"]


def _sentence(rng):
    words, pos, morphs = [], [], []
    for role in ("subj", "obj"):
        plural = rng.random() < 0.4
        noun = rng.choice(NOUNS) + ("s" if plural else "")
        if role == "subj":
            words.append(rng.choice(DETS).capitalize())
        else:
            words.append(rng.choice(VERBS))
            pos.append("VERB")
            morphs.append("Tense=Pres")
            words.append(rng.choice(DETS))
        pos.append("DET")
        morphs.append("")
        words.append(noun)
        pos.append("NOUN")
        morphs.append("Number=Plur" if plural else "Number=Sing")
    words.append(".")
    pos.append("PUNCT")
    morphs.append("")
    return words, pos, morphs


def _docs(n, seed):
    rng = random.Random(seed)
    docs = []
    for _ in range(n):
        words, pos, morphs, starts = [], [], [], []
        for _s in range(rng.randint(1, 3)):
            w, p, m = _sentence(rng)
            starts += [True] + [False] * (len(w) - 1)
            words += w
            pos += p
            morphs += m
        docs.append(Doc(words, pos=pos, morphs=morphs, sent_starts=starts))
    return docs


def test_senter_and_morphologizer_learn_on_a_toy_grammar(tmp_path):
    from spacy_ray_b200.nn.layers import fix_random_seed
    from spacy_ray_b200.pipeline.language import Language

    fix_random_seed(0)
    nlp = Language.from_config(Config().from_str(CFG, interpolate=False))
    train = [Example.from_doc(d) for d in _docs(200, 0)]
    dev = [Example.from_doc(d) for d in _docs(40, 1)]
    nlp.initialize(lambda: train)
    morph = nlp.get_pipe("morphologizer")
    assert "Number=Plur|POS=NOUN" in morph.labels and "POS=DET" in morph.labels and "POS=PUNCT" in morph.labels
    assert nlp.get_pipe("senter").labels == ["I", "S"]
    opt = nlp.create_optimizer()
    first = last = None
    for step in range(60):
        losses = {}
        lo = (step * 16) % (len(train) - 16)
        nlp.update(train[lo:lo + 16], drop=0.0, sgd=opt, losses=losses)
        if step == 0:
            first = dict(losses)
        last = dict(losses)
    assert float(last["senter"]) < 0.5 * float(first["senter"])
    assert float(last["morphologizer"]) < 0.5 * float(first["morphologizer"])
    scores = nlp.evaluate(dev)
    assert scores["sents_f"] > 0.9 and scores["pos_acc"] > 0.9 and scores["morph_acc"] > 0.8, scores
    doc = next(iter(nlp.pipe([dev[0].reference.copy_unannotated()])))
    assert doc.sent_starts[0] is True and len(doc.pos) == len(doc) and len(doc.morphs) == len(doc)
    # checkpoint round trip keeps labels and predictions
    from spacy_ray_b200.pipeline import load

    nlp.to_disk(tmp_path / "m")
    nlp2 = load(tmp_path / "m")
    doc2 = next(iter(nlp2.pipe([dev[0].reference.copy_unannotated()])))
    assert (doc2.pos, doc2.morphs, doc2.sent_starts) == (doc.pos, doc.morphs, doc.sent_starts)


def test_senter_gold_falls_back_to_the_dependency_roots():
    d = Doc(["a", "b", "c", "d"], heads=[1, 1, 3, 3], deps=["x", "ROOT", "x", "ROOT"])
    assert d.gold_sent_starts() == [True, False, True, False]
    assert Doc(["a"]).gold_sent_starts() is None
    with pytest.raises(ValueError):
        Doc(["a", "b"], pos=["X"])


def test_docbin_round_trips_the_new_fields(tmp_path):
    from spacy_ray_b200.training.docbin import DocBin

    d = Doc(["Cats", "sleep", "."], pos=["NOUN", "VERB", "PUNCT"], morphs=["Number=Plur", "Tense=Pres", ""],
            lemmas=["cat", "sleep", "."], sent_starts=[True, False, False], cats={"POSITIVE": 1.0, "NEGATIVE": 0.0})
    db = DocBin(docs=[d, Doc(["x"])])
    db.to_disk(tmp_path / "c.spacy")
    a, b = list(DocBin().from_disk(tmp_path / "c.spacy").get_docs())
    assert a.pos == d.pos and a.lemmas == d.lemmas and a.sent_starts == d.sent_starts and a.cats == d.cats
    assert a.morphs == ["Number=Plur", "Tense=Pres", None]
    assert b.pos is None and b.cats == {} and b.sent_starts is None
    assert Doc.from_dict(d.to_dict()).to_dict() == d.to_dict()


TEXTCAT_CFG = """
[nlp]
lang = "en"
pipeline = ["{factory}"]

[components]

[components.{factory}]
factory = "{factory}"
threshold = 0.5

[components.{factory}.model]
@architectures = "spacy.TextCatCNN.v2"
exclusive_classes = {exclusive}
{extra}

[components.{factory}.model.tok2vec]
@architectures = "spacy.HashEmbedCNN.v2"
width = 32
depth = 1
embed_size = 300
window_size = 1
maxout_pieces = 3
subword_features = true
pretrained_vectors = null
"""

GOOD = ["great", "lovely", "superb", "fine"]
BAD = ["awful", "dreadful", "poor", "bad"]
FILL = ["the", "film", "was", "really", "and", "plot", "acting", "music"]


def _review(rng, multilabel):
    pos = rng.random() < 0.5
    words = [rng.choice(FILL) for _ in range(rng.randint(3, 8))]
    words.insert(rng.randrange(len(words) + 1), rng.choice(GOOD if pos else BAD))
    long = len(words) >= 7
    if multilabel:
        cats = {"POSITIVE": float(pos), "LONG": float(long)}
    else:
        cats = {"POSITIVE": float(pos), "NEGATIVE": float(not pos)}
    return Doc(words, cats=cats)


@pytest.mark.parametrize("factory,exclusive,extra", [("textcat", "true", ""),
                                                     ("textcat_multilabel", "false", "use_reduce_max = true")])
def test_textcat_learns(factory, exclusive, extra, tmp_path):
    from spacy_ray_b200.nn.layers import fix_random_seed
    from spacy_ray_b200.pipeline import load
    from spacy_ray_b200.pipeline.language import Language

    fix_random_seed(0)
    cfg = TEXTCAT_CFG.format(factory=factory, exclusive=exclusive, extra=extra)
    nlp = Language.from_config(Config().from_str(cfg, interpolate=False))
    rng = random.Random(3)
    train = [Example.from_doc(_review(rng, factory != "textcat")) for _ in range(300)]
    dev = [Example.from_doc(_review(rng, factory != "textcat")) for _ in range(60)]
    nlp.initialize(lambda: train)
    pipe = nlp.get_pipe(factory)
    assert pipe.labels == sorted(train[0].reference.cats)
    opt = nlp.create_optimizer()
    hist = []
    for step in range(80 if factory == "textcat" else 160):
        losses = {}
        lo = (step * 16) % (len(train) - 16)
        nlp.update(train[lo:lo + 16], drop=0.0, sgd=opt, losses=losses)
        hist.append(float(losses[factory]))
    assert sum(hist[-5:]) < 0.5 * sum(hist[:5]), hist
    scores = nlp.evaluate(dev)
    assert scores["cats_score"] > (0.9 if factory == "textcat" else 0.85), scores
    doc = next(iter(nlp.pipe([dev[0].reference.copy_unannotated()])))
    assert set(doc.cats) == set(pipe.labels) and all(0.0 <= v <= 1.0 for v in doc.cats.values())
    if factory == "textcat":
        assert abs(sum(doc.cats.values()) - 1.0) < 1e-3
    nlp.to_disk(tmp_path / "m")
    doc2 = next(iter(load(tmp_path / "m").pipe([dev[0].reference.copy_unannotated()])))
    assert doc2.cats == pytest.approx(doc.cats)


def test_textcat_rejects_a_model_of_the_other_kind():
    from spacy_ray_b200.pipeline.language import Language

    cfg = TEXTCAT_CFG.format(factory="textcat", exclusive="false", extra="")
    nlp = Language.from_config(Config().from_str(cfg, interpolate=False))
    with pytest.raises(ValueError):
        nlp.initialize(lambda: [Example.from_doc(Doc(["a"], cats={"A": 1.0, "B": 0.0}))])


BOW_CFG = """
[nlp]
lang = "en"
pipeline = ["textcat"]

[components]

[components.textcat]
factory = "textcat"

[components.textcat.model]
@architectures = "spacy.TextCatBOW.v2"
exclusive_classes = true
ngram_size = 2
no_output_layer = false
nO = null
"""

ENSEMBLE_CFG = """
[nlp]
lang = "en"
pipeline = ["textcat"]

[components]

[components.textcat]
factory = "textcat"

[components.textcat.model]
@architectures = "spacy.TextCatEnsemble.v2"
nO = null

[components.textcat.model.linear_model]
@architectures = "spacy.TextCatBOW.v2"
exclusive_classes = true
ngram_size = 1
no_output_layer = false

[components.textcat.model.tok2vec]
@architectures = "spacy.Tok2Vec.v2"

[components.textcat.model.tok2vec.embed]
@architectures = "spacy.MultiHashEmbed.v2"
width = 32
rows = [500, 250, 250, 250]
attrs = ["NORM", "PREFIX", "SUFFIX", "SHAPE"]
include_static_vectors = false

[components.textcat.model.tok2vec.encode]
@architectures = "spacy.MaxoutWindowEncoder.v2"
width = 32
window_size = 1
maxout_pieces = 3
depth = 1
"""


@pytest.mark.parametrize("cfg_text", [BOW_CFG, ENSEMBLE_CFG], ids=["bow", "ensemble"])
def test_stock_textcat_architectures_train(cfg_text, tmp_path):
    """The two architectures `spacy init config` writes for textcat (efficiency: TextCatBOW, accuracy:
    TextCatEnsemble) resolve and learn."""
    from spacy_ray_b200.nn.layers import fix_random_seed
    from spacy_ray_b200.pipeline import load
    from spacy_ray_b200.pipeline.language import Language

    fix_random_seed(0)
    nlp = Language.from_config(Config().from_str(cfg_text, interpolate=False))
    rng = random.Random(5)
    train = [Example.from_doc(_review(rng, False)) for _ in range(300)]
    dev = [Example.from_doc(_review(rng, False)) for _ in range(60)]
    nlp.initialize(lambda: train)
    opt = nlp.create_optimizer()
    hist = []
    for step in range(300):                      # the sparse rows see few updates each: slower than the CNN
        losses = {}
        lo = (step * 16) % (len(train) - 16)
        nlp.update(train[lo:lo + 16], drop=0.0, sgd=opt, losses=losses)
        hist.append(float(losses["textcat"]))
    assert sum(hist[-5:]) < 0.6 * sum(hist[:5]), (hist[:5], hist[-5:])
    scores = nlp.evaluate(dev)
    assert scores["cats_score"] > 0.9, scores
    nlp.to_disk(tmp_path / "m")
    d1 = next(iter(nlp.pipe([dev[0].reference.copy_unannotated()])))
    d2 = next(iter(load(tmp_path / "m").pipe([dev[0].reference.copy_unannotated()])))
    assert d2.cats == pytest.approx(d1.cats)


def test_bow_ngrams_stay_inside_their_doc():
    import numpy as np
    import torch

    from spacy_ray_b200.models.textcat import _ngram_features
    from spacy_ray_b200.pipeline.language import Language

    nlp = Language.from_config(Config().from_str(BOW_CFG, interpolate=False))
    docs = [Doc(["a", "b", "c"]), Doc(["d"]), Doc(["e", "f"])]
    batch = nlp.make_batch(docs)
    f, d = _ngram_features(batch, 2, 1 << 18)
    counts = np.bincount(d.numpy(), minlength=3).tolist()
    assert counts == [3 + 2, 1 + 0, 2 + 1]            # unigrams + bigrams per doc, none across a boundary
    assert int(f.min()) >= 0 and int(f.max()) < (1 << 18)


def test_engine_staging_serves_the_tagger_subclasses():
    """``engine.Trainer`` treats every token tagger as a "tagger": the packed staging buffer must carry exactly
    the gold ids the generic update would use for senter / morphologizer (host side only; no GPU needed)."""
    import numpy as np

    from spacy_ray_b200.engine.trainer import ExampleStore, Trainer, _kind, _Layout, fill_stage, make_stage
    from spacy_ray_b200.nn.layers import fix_random_seed
    from spacy_ray_b200.pipeline.language import Language

    fix_random_seed(0)
    nlp = Language.from_config(Config().from_str(CFG, interpolate=False))
    examples = [Example.from_doc(d) for d in _docs(40, 2)]
    nlp.initialize(lambda: examples)
    heads = [(n, c, _kind(c)) for n, c in nlp.pipeline if getattr(c, "is_trainable", False)]
    assert [k for _, _, k in heads] == ["tok2vec", "tagger", "tagger"]
    assert Trainer.unsupported_reason(nlp) in (None, "") or "CUDA" in str(Trainer.unsupported_reason(nlp))
    store = ExampleStore(examples, heads)
    lay = _Layout(rows=512, docs=8, lmax=64, slots=tuple(store.slots))
    stage = make_stage(lay, store, pin=False)
    ids = np.array([1, 5, 8, 13, 21, 30, 33, 39])
    fill_stage(store, lay, stage, ids)
    chosen = [examples[i] for i in ids]
    batch = nlp.make_batch([eg.predicted for eg in chosen])
    rows = stage["rows"]
    assert rows == batch.n_rows
    for name in ("senter", "morphologizer"):
        want = nlp.get_pipe(name)._gold_labels(chosen, batch).numpy()
        np.testing.assert_array_equal(stage["np"]["gold"][name][:rows], want)
        assert (want >= 0).sum() == sum(len(eg) for eg in chosen)


def test_trainable_lemmatizer_learns_suffix_rules(tmp_path):
    from spacy_ray_b200.nn.layers import fix_random_seed
    from spacy_ray_b200.pipeline import load
    from spacy_ray_b200.pipeline.components import TrainableLemmatizer as TL
    from spacy_ray_b200.pipeline.language import Language

    assert TL.rule_of("Cats", "cat") == "L|1|" and TL.apply_rule("Cats", "L|1|") == "cat"
    assert TL.rule_of("running", "run") == "K|4|" and TL.rule_of("was", "be") == "K|3|be"
    assert TL.apply_rule("is", "K|3|be") is None and TL.rule_of("Paris", "Paris") == "K|0|"
    cfg = CFG.replace('pipeline = ["tok2vec", "senter", "morphologizer"]', 'pipeline = ["tok2vec", "trainable_lemmatizer"]')
    cfg = cfg[: cfg.index("[components.senter]")] + """[components.trainable_lemmatizer]
factory = "trainable_lemmatizer"
min_tree_freq = 2
backoff = "orth"

[components.trainable_lemmatizer.model]
@architectures = "spacy.Tagger.v2"

[components.trainable_lemmatizer.model.tok2vec]
@architectures = "spacy.Tok2VecListener.v1"
width = 32
upstream = "*"
""" + cfg[cfg.index("[corpora]"):]
    fix_random_seed(0)
    nlp = Language.from_config(Config().from_str(cfg, interpolate=False))

    def lemma(w):
        lw = w.lower()
        if lw.endswith("s") and lw[:-1] in NOUNS:
            return lw[:-1]
        if lw in VERBS:
            return lw[:-1]                      # sees -> see, likes -> like ...
        return lw if lw in DETS else w

    docs = _docs(200, 0) + _docs(40, 1)
    for d in docs:
        d.lemmas = [lemma(w) for w in d.words]
    train = [Example.from_doc(d) for d in docs[:200]]
    dev = [Example.from_doc(d) for d in docs[200:]]
    nlp.initialize(lambda: train)
    pipe = nlp.get_pipe("trainable_lemmatizer")
    assert "K|1|" in pipe.labels and "K|0|" in pipe.labels and "L|0|" in pipe.labels
    opt = nlp.create_optimizer()
    for step in range(60):
        lo = (step * 16) % (len(train) - 16)
        nlp.update(train[lo:lo + 16], drop=0.0, sgd=opt, losses={})
    scores = nlp.evaluate(dev)
    assert scores["lemma_acc"] > 0.95, scores
    nlp.to_disk(tmp_path / "m")
    d1 = next(iter(nlp.pipe([dev[0].reference.copy_unannotated()])))
    d2 = next(iter(load(tmp_path / "m").pipe([dev[0].reference.copy_unannotated()])))
    assert d1.lemmas == d2.lemmas and len(d1.lemmas) == len(d1)


SPANCAT_CFG = """
[nlp]
lang = "en"
pipeline = ["spancat"]

[components]

[components.spancat]
factory = "spancat"
spans_key = "sc"
threshold = 0.5
max_positive = null

[components.spancat.suggester]
@misc = "spacy.ngram_suggester.v1"
sizes = [1, 2, 3]

[components.spancat.model]
@architectures = "spacy.SpanCategorizer.v1"

[components.spancat.model.reducer]
@layers = "spacy.mean_max_reducer.v1"
hidden_size = 64

[components.spancat.model.scorer]
@layers = "spacy.LinearLogistic.v1"
nO = null
nI = null

[components.spancat.model.tok2vec]
@architectures = "spacy.HashEmbedCNN.v2"
width = 32
depth = 2
embed_size = 300
window_size = 1
maxout_pieces = 3
subword_features = true
pretrained_vectors = null
"""


def _span_doc(rng):
    """"<det> <noun>" is an NP, a noun alone is also a THING (overlapping spans), "<verb> <det> <noun>" is a VP."""
    words, spans = [], []
    for _ in range(rng.randint(1, 2)):
        s0 = len(words)
        words += [rng.choice(DETS), rng.choice(NOUNS), rng.choice(VERBS), rng.choice(DETS), rng.choice(NOUNS), "."]
        spans += [(s0, s0 + 2, "NP"), (s0 + 1, s0 + 2, "THING"), (s0 + 3, s0 + 5, "NP"), (s0 + 4, s0 + 5, "THING"),
                  (s0 + 2, s0 + 5, "VP")]
    return Doc(words, spans={"sc": spans})


def test_spancat_learns_overlapping_spans(tmp_path):
    from spacy_ray_b200.models.spancat import ngram_suggester
    from spacy_ray_b200.nn.layers import fix_random_seed
    from spacy_ray_b200.pipeline import load
    from spacy_ray_b200.pipeline.language import Language
    from spacy_ray_b200.training.docbin import DocBin

    assert ngram_suggester([1, 2])([3, 1]) == [(0, 0, 1), (0, 1, 2), (0, 2, 3), (0, 0, 2), (0, 1, 3), (1, 0, 1)]
    fix_random_seed(0)
    nlp = Language.from_config(Config().from_str(SPANCAT_CFG, interpolate=False))
    rng = random.Random(7)
    train = [Example.from_doc(_span_doc(rng)) for _ in range(200)]
    dev = [Example.from_doc(_span_doc(rng)) for _ in range(40)]
    nlp.initialize(lambda: train)
    assert nlp.get_pipe("spancat").labels == ["NP", "THING", "VP"]
    opt = nlp.create_optimizer()
    hist = []
    for step in range(120):
        losses = {}
        lo = (step * 16) % (len(train) - 16)
        nlp.update(train[lo:lo + 16], drop=0.0, sgd=opt, losses=losses)
        hist.append(float(losses["spancat"]))
    assert sum(hist[-5:]) < 0.3 * sum(hist[:5]), (hist[:5], hist[-5:])
    scores = nlp.evaluate(dev)
    assert scores["spans_sc_f"] > 0.9, scores
    doc = next(iter(nlp.pipe([dev[0].reference.copy_unannotated()])))
    assert any(l == "THING" for _s, _e, l in doc.spans["sc"]) and any(l == "VP" for _s, _e, l in doc.spans["sc"])
    # span groups survive DocBin and the checkpoint
    DocBin(docs=[dev[0].reference]).to_disk(tmp_path / "s.spacy")
    back = next(iter(DocBin().from_disk(tmp_path / "s.spacy").get_docs()))
    assert back.spans == dev[0].reference.spans
    nlp.to_disk(tmp_path / "m")
    doc2 = next(iter(load(tmp_path / "m").pipe([dev[0].reference.copy_unannotated()])))
    assert doc2.spans == doc.spans


def test_sentencizer_is_a_non_trainable_pipe(tmp_path):
    from spacy_ray_b200.pipeline import load
    from spacy_ray_b200.pipeline.language import Language

    cfg = """
[nlp]
lang = "en"
pipeline = ["sentencizer"]

[components]

[components.sentencizer]
factory = "sentencizer"
"""
    nlp = Language.from_config(Config().from_str(cfg, interpolate=False))
    nlp.initialize()
    doc = nlp(Doc(["Hello", "world", ".", "How", "are", "you", "?", "!", "Fine"]))
    assert doc.sent_starts == [True, False, False, True, False, False, False, False, True]
    egs = [Example.from_doc(d) for d in _docs(10, 4)]
    assert nlp.update(egs, losses={}) == {}                      # nothing to train
    assert nlp.evaluate(egs)["sents_f"] == pytest.approx(1.0)    # the toy docs end every sentence with "."
    nlp.to_disk(tmp_path / "m")
    assert load(tmp_path / "m")(Doc(["A", ".", "B"])).sent_starts == [True, False, True]
