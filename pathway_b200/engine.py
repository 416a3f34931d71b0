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


@dataclass(frozen=True)
class ExternalIndexData:
    """engine.pyi:1095-1099 / src/python_api/external_index_wrappers.rs:92-121.  `table` is the index stream;
    column paths are column names here."""

    table: object
    data_column: str
    filter_data_column: str | None


@dataclass(frozen=True)
class ExternalIndexQuery:
    """engine.pyi:1101-1106 / external_index_wrappers.rs:123-162"""

    table: object
    query_column: str
    limit_column: str | None
    filter_column: str | None


class Scope:
    """The one `Scope` method on this path (engine.pyi:688-694, src/python_api.rs:3367-3381): make one index instance
    from the factory (one per worker in the engine) and run the as-of-now operator over the two streams
    (src/engine/dataflow/operators/external_index.rs:76-158)."""

    def use_external_index_as_of_now(self, index: ExternalIndexData, queries: ExternalIndexQuery,
                                     table_properties=None, external_index_factory: ExternalIndexFactory = None,
                                     *, _inner_override=None):
        from .external_index import IndexDerivedImpl, use_external_index_as_of_now

        def acc(name):
            return None if name is None else (lambda row: row[name])

        errors: list = []
        if callable(_inner_override):  # a zero-argument maker: one fresh checker per retriever of a HybridIndex
            _inner_override = _inner_override()
        impl = IndexDerivedImpl(external_index_factory.make_instance(inner_override=_inner_override), errors,
                                acc(index.data_column), acc(index.filter_data_column), acc(queries.query_column),
                                acc(queries.limit_column), acc(queries.filter_column))
        out = use_external_index_as_of_now(
            [(r.key, r.values, r.time, r.diff) for r in index.table.rows],
            [(r.key, r.values, r.time, r.diff) for r in queries.table.rows], impl)
        return out, errors
