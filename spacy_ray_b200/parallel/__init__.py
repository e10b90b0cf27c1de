"""Parallelism: ownership partitioning, the async peer-proxy protocol, the
synchronous flat-bucket sharded proxy, comm backends, the actor runtime."""
from .util import KeyT, make_key, divide_params, divide_params_balanced, set_params_proxy, Timer, ManyTimer
from .proxies import PeerProxy, RayPeerProxy, RayOptimizer
from .sync_proxy import FlatLayout, ShardedSyncProxy, LocalComm, TorchDistComm

__all__ = [
    "KeyT", "make_key", "divide_params", "divide_params_balanced", "set_params_proxy", "Timer", "ManyTimer",
    "PeerProxy", "RayPeerProxy", "RayOptimizer", "FlatLayout", "ShardedSyncProxy", "LocalComm", "TorchDistComm",
]
