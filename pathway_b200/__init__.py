"""pathway_b200 -- B200-native drop-in for Pathway's brute-force KNN index path.

Only what the hot path needs: csrc/ (sm_100a CUDA kernels + the
C ABI of include/pwknn.h) and the host-side mirror of the reference interface.
"""

from .engine import ERROR, BruteForceKnnMetricKind, ExternalIndexFactory, USearchMetricKind  # noqa: F401

__all__ = ["ERROR", "BruteForceKnnMetricKind", "ExternalIndexFactory", "USearchMetricKind"]
