from ...engine import BruteForceKnnMetricKind  # noqa: F401  (pw.stdlib.indexing.BruteForceKnnMetricKind)
from .nearest_neighbors import BruteForceKnn, BruteForceKnnFactory, DefaultKnnFactory  # noqa: F401
