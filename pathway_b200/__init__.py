"""pathway_b200 -- B200-native drop-in for Pathway's brute-force KNN index path.

Only what the hot path needs: csrc/ (sm_100a CUDA kernels + the
C ABI of include/pwknn.h) and the host-side mirror of the reference interface.

Import paths mirror the reference's public ones (python/pathway/__init__.py:79,99): `pathway_b200.indexing` is
`pathway_b200.stdlib.indexing` (as `pw.indexing` is `pw.stdlib.indexing`), `pathway_b200.engine` holds the pyo3
surface (`ExternalIndexFactory`, `BruteForceKnnMetricKind`, ...), `pathway_b200.debug` the table scaffolding.
"""

from . import debug, engine  # noqa: F401
from .engine import ERROR, BruteForceKnnMetricKind, ExternalIndexFactory, USearchMetricKind  # noqa: F401
from .stdlib import indexing  # noqa: F401

__all__ = ["ERROR", "BruteForceKnnMetricKind", "ExternalIndexFactory", "USearchMetricKind", "debug", "engine",
           "indexing"]
