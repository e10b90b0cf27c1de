"""Registered model architectures (the spaCy names the reference's configs use).

Shapes follow SURVEY.md 2.6.  Registered under the upstream registry names so
an upstream ``.cfg`` resolves unchanged:

``spacy.Tok2Vec.v2``, ``spacy.MultiHashEmbed.v2``, ``spacy.MaxoutWindowEncoder.v2``,
``spacy.HashEmbedCNN.v2``, ``spacy.Tok2VecListener.v1``, ``spacy.Tagger.v1/v2``,
``spacy.TransitionBasedParser.v2``, ``spacy.TextCatCNN.v1/v2``, ``spacy.TextCatReduce.v1``, ``spacy.TextCatBOW.v1-3``,
``spacy.TextCatEnsemble.v2``, ``spacy.SpanCategorizer.v1`` (+ ``spacy.ngram_suggester.v1``).
"""
from ..config import registry
from ..nn import layers as L
from .spancat import build_spancat_model
from .tagger import build_tagger_model
from .textcat import build_textcat_bow, build_textcat_ensemble, build_textcat_model
from .transition_model import build_transition_model, TransitionModelOutput

for _v in ("v1", "v2"):
    registry.architectures.register(f"spacy.Tok2Vec.{_v}", L.Tok2Vec)
    registry.architectures.register(f"spacy.MultiHashEmbed.{_v}", L.MultiHashEmbed)
    registry.architectures.register(f"spacy.MaxoutWindowEncoder.{_v}", L.MaxoutWindowEncoder)
    registry.architectures.register(f"spacy.HashEmbedCNN.{_v}", L.HashEmbedCNN)
    registry.architectures.register(f"spacy.Tagger.{_v}", build_tagger_model)
registry.architectures.register("spacy.Tok2VecListener.v1", L.Tok2VecListener)
for _n in ("spacy.TextCatCNN.v1", "spacy.TextCatCNN.v2", "spacy.TextCatReduce.v1"):
    registry.architectures.register(_n, build_textcat_model)
for _v in ("v1", "v2", "v3"):
    registry.architectures.register(f"spacy.TextCatBOW.{_v}", build_textcat_bow)
registry.architectures.register("spacy.TextCatEnsemble.v2", build_textcat_ensemble)
registry.architectures.register("spacy.SpanCategorizer.v1", build_spancat_model)
for _v in ("v1", "v2", "v3"):
    registry.architectures.register(f"spacy.TransitionBasedParser.{_v}", build_transition_model)

__all__ = ["build_tagger_model", "build_textcat_model", "build_transition_model", "TransitionModelOutput"]
