from ...engine import BruteForceKnnMetricKind, USearchMetricKind  # noqa: F401  (pw.stdlib.indexing.*MetricKind)
from .nearest_neighbors import (BruteForceKnn, BruteForceKnnFactory, DefaultKnnFactory, USearchKnn,  # noqa: F401
                                UsearchKnnFactory)
from .data_index import DataIndex  # noqa: F401
from .hybrid_index import HybridIndex, HybridIndexFactory  # noqa: F401
from .retrievers import InnerIndexFactory  # noqa: F401
