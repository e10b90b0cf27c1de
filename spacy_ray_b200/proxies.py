"""Reference-compatible module path (``spacy_ray.proxies``)."""
from .parallel.proxies import PeerProxy, RayPeerProxy, RayOptimizer  # noqa: F401
from .parallel.sync_proxy import ShardedSyncProxy, FlatLayout  # noqa: F401
