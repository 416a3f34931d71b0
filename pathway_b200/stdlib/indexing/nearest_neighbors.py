"""Drop-in mirror of the brute-force classes of
/root/reference/python/pathway/stdlib/indexing/nearest_neighbors.py
(`BruteForceKnn` :169-258, `BruteForceKnnFactory` :485-528, `DefaultKnnFactory` :573-586):
same class names, kwargs, defaults and error behaviour; the factory they build now makes the
B200 index.  `pw.Table` is replaced by the tiny row container in pathway_b200.debug."""

from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable

from ...debug import _INDEX_REPLY, ColumnReference, Table  # noqa: F401
from ...engine import BruteForceKnnMetricKind, ExternalIndexFactory, USearchMetricKind
from .retrievers import InnerIndexFactory

_NO_OF_MATCHES = "_pw_number_of_matches"  # colnames.py


def _calculate_embeddings(column: ColumnReference, embedder: Callable | None) -> ColumnReference:
    """nearest_neighbors.py:51-61"""
    if embedder is None:
        return column
    table = column.table.with_columns(_pw_embedded_column=lambda row: embedder(row[column.name]))
    return table._pw_embedded_column


def _as_of_now(inner, index_factory, query_column, number_of_matches, metadata_filter, _inner_override):
    """shared tail of USearchKnn.query_as_of_now / BruteForceKnn.query_as_of_now (nearest_neighbors.py:127-166,217-258)"""
    index = inner._data_column.table
    query_column = _calculate_embeddings(query_column, inner.embedder)
    queries = query_column.table
    number_of_matches_ref = number_of_matches
    if isinstance(number_of_matches, int):
        queries = queries.with_columns(**{_NO_OF_MATCHES: number_of_matches})
        number_of_matches_ref = queries[_NO_OF_MATCHES]
        query_column = queries[query_column.name]
    metadata_name = inner.metadata_column.name if inner.metadata_column is not None else None
    return index._external_index_as_of_now(
        queries, index_column=inner._data_column, query_column=query_column, index_factory=index_factory,
        query_responses_limit_column=number_of_matches_ref, index_filter_data_column=metadata_name,
        query_filter_column=metadata_filter, _inner_override=_inner_override)


@dataclass(frozen=True, kw_only=True)
class USearchKnn:
    """nearest_neighbors.py:64-166: same kwargs; answered exactly by the B200 index (IP / L2SQ / COS)."""

    data_column: ColumnReference
    metadata_column: ColumnReference | None = None
    dimensions: int
    reserved_space: int
    metric: USearchMetricKind
    connectivity: int = 0
    expansion_add: int = 0
    expansion_search: int = 0
    embedder: Callable | None = None
    _data_column: ColumnReference = field(init=False)

    def __post_init__(self):
        object.__setattr__(self, "_data_column", _calculate_embeddings(self.data_column, self.embedder))

    def query(self, query_column, number_of_matches=3, metadata_filter=None):
        raise NotImplementedError(
            "Currently, usearch knn index is supported only in the as-of-now variant"
        )

    def query_as_of_now(self, query_column: ColumnReference, number_of_matches: ColumnReference | int = 3,
                        metadata_filter: ColumnReference | None = None, _inner_override=None) -> Table:
        names = ("dimensions", "reserved_space", "metric", "connectivity", "expansion_add", "expansion_search")
        index_factory = ExternalIndexFactory.usearch_knn_factory(**{n: getattr(self, n) for n in names})
        return _as_of_now(self, index_factory, query_column, number_of_matches, metadata_filter, _inner_override)


@dataclass(frozen=True, kw_only=True)
class BruteForceKnn:
    """nearest_neighbors.py:169-258"""

    data_column: ColumnReference
    metadata_column: ColumnReference | None = None
    dimensions: int
    reserved_space: int
    auxiliary_space: int = 1024 * 128
    metric: BruteForceKnnMetricKind
    embedder: Callable | None = None
    _data_column: ColumnReference = field(init=False)

    def __post_init__(self):
        object.__setattr__(self, "_data_column", _calculate_embeddings(self.data_column, self.embedder))

    def query(self, query_column, number_of_matches=3, metadata_filter=None):
        """Currently, brute force knn index is supported only in the as-of-now variant"""
        raise NotImplementedError(
            "Currently, brute force knn index is supported only in the as-of-now variant"
        )

    def query_as_of_now(self, query_column: ColumnReference, number_of_matches: ColumnReference | int = 3,
                        metadata_filter: ColumnReference | None = None, _inner_override=None) -> Table:
        names = ("dimensions", "reserved_space", "auxiliary_space", "metric")
        index_factory = ExternalIndexFactory.brute_force_knn_factory(**{n: getattr(self, n) for n in names})
        return _as_of_now(self, index_factory, query_column, number_of_matches, metadata_filter, _inner_override)


@dataclass(kw_only=True)
class BruteForceKnnFactory(InnerIndexFactory):
    """nearest_neighbors.py:485-528 (KnnIndexFactory fields: dimensions, embedder)"""

    dimensions: int | None = None
    embedder: Callable | None = None
    reserved_space: int = 400
    auxiliary_space: int = 1024 * 128
    metric: BruteForceKnnMetricKind = BruteForceKnnMetricKind.COS

    def build_inner_index(self, data_column: ColumnReference, metadata_column: ColumnReference | None = None):
        dims = self.dimensions
        if dims is None and self.embedder is not None:
            dims = len(self.embedder("."))  # KnnIndexFactory probes the embedder (nearest_neighbors.py:452-460)
        assert isinstance(dims, int), \
            "`dimensions` is not set, this may indicate something is wrong with embedder."
        passed = {n: getattr(self, n) for n in ("reserved_space", "auxiliary_space", "metric", "embedder")}
        return BruteForceKnn(data_column=data_column, metadata_column=metadata_column, dimensions=dims, **passed)


@dataclass(kw_only=True)
class UsearchKnnFactory(InnerIndexFactory):
    """nearest_neighbors.py:430-483"""

    dimensions: int | None = None
    embedder: Callable | None = None
    reserved_space: int = 400
    metric: USearchMetricKind = USearchMetricKind.COS
    connectivity: int = 0
    expansion_add: int = 0
    expansion_search: int = 0

    def build_inner_index(self, data_column: ColumnReference, metadata_column: ColumnReference | None = None):
        dims = self.dimensions
        if dims is None and self.embedder is not None:
            dims = len(self.embedder("."))
        assert isinstance(dims, int), \
            "`dimensions` is not set, this may indicate something is wrong with embedder."
        passed = {n: getattr(self, n) for n in ("reserved_space", "metric", "connectivity", "expansion_add",
                                                 "expansion_search", "embedder")}
        return USearchKnn(data_column=data_column, metadata_column=metadata_column, dimensions=dims, **passed)


@dataclass(kw_only=True)
class DefaultKnnFactory(BruteForceKnnFactory):
    """nearest_neighbors.py:573-586: the default KNN index is the brute-force one."""
