__title__ = "spacy_ray_b200"
__version__ = "0.1.0"
__summary__ = "B200-native parallel training for spaCy-style pipelines (spacy-ray capabilities, no Ray)"
__reference__ = "explosion/spacy-ray @ 09ffba5 (v0.1.4)"
