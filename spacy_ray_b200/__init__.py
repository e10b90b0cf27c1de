"""spacy_ray_b200: B200-native parallel training for spaCy-style pipelines.

Capabilities of explosion/spacy-ray (reference @ 09ffba5) rebuilt from scratch:
the ``spacy ray train`` CLI, the ``Worker`` / peer-proxy API with
parameter-ownership sharding, the console logger - with the Ray RPC transport
replaced by sm_100a kernels over NVLink peer memory, and with its own model /
config / optimizer / training-loop layers (spaCy, thinc and Ray are not
required).  See DESIGN.md.
"""
from .about import __version__  # noqa: F401
from . import config as _config  # noqa: F401
from .config import Config, registry, load_config  # noqa: F401
from . import models as _models  # noqa: F401  (registers architectures)
from . import training as _training  # noqa: F401  (registers optimizers, batchers, loggers, readers)
from .pipeline import Language, Doc, Example, blank, load  # noqa: F401
from .worker import Worker, Evaluator, FakeOptimizer, thread_training  # noqa: F401,E402
from .train_cli import ray_train, ray_cli  # noqa: F401,E402
from .parallel.proxies import RayPeerProxy, RayOptimizer, PeerProxy  # noqa: F401,E402
