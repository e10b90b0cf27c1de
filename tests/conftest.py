import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on a B200 via gpurun)")
    config.addinivalue_line("markers", "multigpu: needs >= 2 CUDA devices")
    config.addinivalue_line("markers", "slow: multi-process CPU test")


def pytest_collection_modifyitems(config, items):
    import torch

    have_gpu = torch.cuda.is_available()
    n_gpu = torch.cuda.device_count() if have_gpu else 0
    for item in items:
        if "gpu" in item.keywords and not have_gpu:
            item.add_marker(pytest.mark.skip(reason="no CUDA device"))
        if "multigpu" in item.keywords and n_gpu < 2:
            item.add_marker(pytest.mark.skip(reason="needs >= 2 GPUs"))


@pytest.fixture(autouse=True)
def _cpu_ops_by_default():
    """Every test starts on the CPU reference backend; GPU tests select theirs."""
    from spacy_ray_b200.ops import require_cpu

    require_cpu()
    yield
    require_cpu()


TAGGER_CFG = """
[nlp]
lang = "en"
pipeline = ["tagger"]

[components]

[components.tagger]
factory = "tagger"

[components.tagger.model]
@architectures = "spacy.Tagger.v2"

[components.tagger.model.tok2vec]
@architectures = "spacy.Tok2Vec.v2"

[components.tagger.model.tok2vec.embed]
@architectures = "spacy.MultiHashEmbed.v2"
width = 32
attrs = ["NORM","PREFIX","SUFFIX","SHAPE"]
rows = [500,250,250,250]
include_static_vectors = false

[components.tagger.model.tok2vec.encode]
@architectures = "spacy.MaxoutWindowEncoder.v2"
width = 32
depth = 2
window_size = 1
maxout_pieces = 3

[corpora]

[corpora.train]
@readers = "spacy_ray_b200.SyntheticCorpus.v1"
n_docs = 120
seed = 1
min_len = 4
max_len = 12
vocab_size = 300
tasks = ["tagger"]

[corpora.dev]
@readers = "spacy_ray_b200.SyntheticCorpus.v1"
n_docs = 40
seed = 2
min_len = 4
max_len = 12
vocab_size = 300
tasks = ["tagger"]

[training]
max_steps = 12
eval_frequency = 6
dropout = 0.0

[training.logger]
@loggers = "spacy-ray.ConsoleLogger.v1"

[training.batcher]
@batchers = "spacy.batch_by_words.v1"
size = 200
tolerance = 0.2
"""


@pytest.fixture
def tagger_config():
    from spacy_ray_b200.config import Config

    return Config().from_str(TAGGER_CFG, interpolate=False)


def multi_cfg(pipeline, width=32, depth=2, n_docs=80, max_len=10, hidden=32):
    comps = []
    for name in pipeline:
        if name == "tagger":
            comps.append(f"""
[components.tagger]
factory = "tagger"

[components.tagger.model]
@architectures = "spacy.Tagger.v2"

[components.tagger.model.tok2vec]
@architectures = "spacy.HashEmbedCNN.v2"
width = {width}
depth = {depth}
embed_size = 300
window_size = 1
maxout_pieces = 3
subword_features = true
pretrained_vectors = null
""")
        else:
            comps.append(f"""
[components.{name}]
factory = "{name}"

[components.{name}.model]
@architectures = "spacy.TransitionBasedParser.v2"
state_type = "{name}"
extra_state_tokens = false
hidden_width = {hidden}
maxout_pieces = 2
use_upper = true

[components.{name}.model.tok2vec]
@architectures = "spacy.HashEmbedCNN.v2"
width = {width}
depth = {depth}
embed_size = 300
window_size = 1
maxout_pieces = 3
subword_features = true
pretrained_vectors = null
""")
    tasks = '["tagger","ner","parser"]'
    return f"""
[nlp]
lang = "en"
pipeline = {list(pipeline)!r}

[components]
{''.join(comps)}
[corpora]

[corpora.train]
@readers = "spacy_ray_b200.SyntheticCorpus.v1"
n_docs = {n_docs}
seed = 1
min_len = 4
max_len = {max_len}
vocab_size = 300
tasks = {tasks}

[corpora.dev]
@readers = "spacy_ray_b200.SyntheticCorpus.v1"
n_docs = 30
seed = 2
min_len = 4
max_len = {max_len}
vocab_size = 300
tasks = {tasks}

[training]
max_steps = 10
eval_frequency = 5
dropout = 0.0
""".replace("'", '"')
