from .doc import Doc, Example, hash_string, lex_attrs, word_shape, featurize_words, featurize_words_py
from .language import Language, blank, load, default_config
from .components import Tok2VecComponent, Tagger, EntityRecognizer, DependencyParser, TrainablePipe
from .transitions import BiluoSystem, ArcEagerSystem

__all__ = [
    "Doc", "Example", "hash_string", "lex_attrs", "word_shape", "featurize_words", "featurize_words_py",
    "Language", "blank", "load", "default_config", "Tok2VecComponent", "Tagger", "EntityRecognizer",
    "DependencyParser", "TrainablePipe", "BiluoSystem", "ArcEagerSystem",
]
