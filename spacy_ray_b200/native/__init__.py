"""Native (C++) host runtime: featurisation and batch collation.

Built in-tree by ``spacy_ray_b200.build`` (``g++ -O3 -shared -fPIC``) into
``_host_runtime.so`` and loaded with ctypes.  Every entry point has a
bit-identical pure-Python fallback so CPU-only environments work unbuilt.
"""
from . import featurize  # noqa: F401
