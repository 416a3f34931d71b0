"""Mirror of the slice of ``pathway.engine`` (pyo3) that the KNN path uses.

Reference surface (paths relative to /root/reference):
  * ``BruteForceKnnMetricKind``          src/python_api/external_index_wrappers.rs:224-233
  * ``ExternalIndexFactory.brute_force_knn_factory(*, dimensions, reserved_space,
    auxiliary_space, metric)``           src/python_api/external_index_wrappers.rs:56-71,
                                         python/pathway/engine.pyi:1080-1086
  * ``ExternalIndexFactory.make_instance`` -> ``Box<dyn ExternalIndex>``
                                         src/external_integration/mod.rs:47-49,
                                         brute_force_knn_integration.rs:263-273
``Pointer`` (Key, u128) values are plain Python ints here.
"""

from __future__ import annotations

import enum
from dataclasses import dataclass


class BruteForceKnnMetricKind(enum.Enum):
    L2SQ = 0
    COS = 1


class USearchMetricKind(enum.Enum):
    """src/python_api/external_index_wrappers.rs:164-202.  The B200 index answers IP / L2SQ / COS exactly
    (recall 1.0 where usearch's HNSW is approximate); the other kinds have no dense-GEMM form and raise."""

    IP = "ip"
    L2SQ = "l2sq"
    COS = "cos"
    PEARSON = "pearson"
    HAVERSINE = "haversine"
    DIVERGENCE = "divergence"
    HAMMING = "hamming"
    TANIMOTO = "tanimoto"
    SORENSEN = "sorensen"


class _Error:
    """``Value::Error`` sentinel (src/engine/value.rs)."""

    _inst = None

    def __new__(cls):
        if cls._inst is None:
            cls._inst = super().__new__(cls)
        return cls._inst

    def __repr__(self):
        return "Error"


ERROR = _Error()


@dataclass(frozen=True)
class ExternalIndexFactory:
    """``PyExternalIndexFactory``: an ``Arc<dyn ExternalIndexFactory>``; one
    ``make_instance()`` per worker (src/python_api.rs:3378)."""

    kind: str
    params: dict

    @staticmethod
    def brute_force_knn_factory(*, dimensions: int, reserved_space: int, auxiliary_space: int,
                                metric: BruteForceKnnMetricKind) -> "ExternalIndexFactory":
        # Same signature as the reference; in this build the instance it makes IS the B200 index.
        return ExternalIndexFactory("b200_brute_force", dict(
            dimensions=dimensions, reserved_space=reserved_space, auxiliary_space=auxiliary_space,
            metric=metric, device=None))

    @staticmethod
    def usearch_knn_factory(*, dimensions: int, reserved_space: int, metric: USearchMetricKind, connectivity: int,
                            expansion_add: int, expansion_search: int) -> "ExternalIndexFactory":
        """Signature of engine.pyi:1063-1072.  HNSW parameters are accepted and ignored: the B200 index is exact."""
        table = {USearchMetricKind.IP: 2, USearchMetricKind.L2SQ: 0, USearchMetricKind.COS: 1}
        if metric not in table:
            raise NotImplementedError(f"USearchMetricKind.{metric.name} is not supported by the B200 exact index")
        return ExternalIndexFactory("b200_brute_force", dict(
            dimensions=dimensions, reserved_space=reserved_space, auxiliary_space=1024 * 128,
            metric=table[metric], device=None))

    @staticmethod
    def b200_knn_factory(*, dimensions: int, reserved_space: int, metric: BruteForceKnnMetricKind,
                         auxiliary_space: int = 1024 * 128, device: int | None = None,
                         devices: list[int] | None = None) -> "ExternalIndexFactory":
        """The one new staticmethod a maintainer adds to PyExternalIndexFactory.  `devices`: one
        instance drives several GPUs (pwknn_create_multi)."""
        return ExternalIndexFactory("b200_brute_force", dict(
            dimensions=dimensions, reserved_space=reserved_space, auxiliary_space=auxiliary_space,
            metric=metric, device=device, devices=devices))

    def make_instance(self, inner_override=None):
        """-> ExternalIndex (DerivedFilteredSearchIndex over the non-filtering index),
        brute_force_knn_integration.rs:263-273.  ``inner_override`` lets tests wrap another
        NonFilteringExternalIndex (e.g. a CPU checker) in the same host logic."""
        from .external_index import B200BruteForceKNNIndex, DerivedFilteredSearchIndex

        if inner_override is not None:
            return DerivedFilteredSearchIndex(inner_override)
        p = self.params
        return DerivedFilteredSearchIndex(B200BruteForceKNNIndex(
            p["dimensions"], p["reserved_space"], p["auxiliary_space"], p["metric"], device=p["device"],
            devices=p.get("devices")))
