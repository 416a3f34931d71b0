"""Drop-in mirror of /root/reference/python/pathway/stdlib/indexing/hybrid_index.py
(`HybridIndex` :14-158, `HybridIndexFactory` :161-190): same constructor arguments, defaults and errors.

The reference fuses the retrievers' replies with dataflow operators -- per retriever a udf + flatten + select
(:43-66), then concat_reindex, groupby(query, match).sum, groupby(query, sort_by).tuple and a limit udf (:72-120) --
all on Q*k*R rows per query batch.  Here the whole fusion is ONE call into libpwknn (`pwknn_rrf_fuse`,
include/pwknn.h): the retrievers' replies go in as key arrays, the fused top-n comes back."""

from __future__ import annotations

import ctypes as C

import numpy as np

from ... import _native
from ...debug import _INDEX_REPLY, ColumnReference, Row, Table
from ...engine import ERROR
from ...external_index import array_to_keys, keys_to_array
from .retrievers import InnerIndexFactory


def rrf_fuse(replies_per_retriever: list[list[list[int]]], limits: list[int], k: float):
    """replies_per_retriever[l][q] = matched keys of retriever l for query q in rank order
    -> per query [(key, fused score)] (descending score, at most limits[q])."""
    nq = len(limits)
    n_lists = len(replies_per_retriever)
    keys_arrs, cnt_arrs = [], []
    for replies in replies_per_retriever:
        assert len(replies) == nq
        cnt_arrs.append(np.asarray([len(r) for r in replies], dtype=np.uint32))
        flat = [key for r in replies for key in r]
        keys_arrs.append(keys_to_array(flat) if flat else np.zeros((1, 2), dtype=np.uint64))
    lim = np.asarray([max(0, int(x)) for x in limits], dtype=np.uint32)
    total = max(1, int(lim.sum()))
    out_keys = np.zeros((total, 2), dtype=np.uint64)
    out_scores = np.zeros(total, dtype=np.float64)
    out_counts = np.zeros(max(1, nq), dtype=np.uint32)
    vp = C.c_void_p * n_lists
    rc = _native.lib().pwknn_rrf_fuse(
        n_lists, nq, vp(*[a.ctypes.data for a in keys_arrs]), vp(*[a.ctypes.data for a in cnt_arrs]), None,
        float(k), lim.ctypes.data, out_keys.ctypes.data, out_scores.ctypes.data, out_counts.ctypes.data)
    _native.check(None, rc)
    out, off = [], 0
    for q in range(nq):
        c = int(out_counts[q])
        ks = array_to_keys(out_keys[off:off + c])
        out.append(list(zip(ks, out_scores[off:off + c].tolist())))
        off += int(lim[q])
    return out


class HybridIndex:
    """hybrid_index.py:14-158.  Composes any number of indices with Reciprocal Rank Fusion: a row ranked r-th by a
    retriever gets 1/(k + r) from it, summed over the retrievers."""

    def __init__(self, retrievers: list, k: float = 60):
        if len(retrievers) < 2:
            raise ValueError(
                "HybridIndex requires at least two indices to be provided during initialization"
            )
        self.retrievers = retrievers
        self.k = k

    def _combine_results(self, tables: list[Table], query_table: Table, number_of_matches) -> Table:
        """hybrid_index.py:35-125, as-of-now variant.  Each retriever table holds one reply row per query row
        (same key / time / diff, external_index.rs:145-151); a query nobody answered gets no row (:92-103)."""
        seen: dict = {}
        per_list: list[dict] = []
        errors: list = []
        for t in tables:
            errors.extend(t.error_log)
            d = {}
            for r in t.rows:
                ident = (r.key, r.time, r.diff)
                d[ident] = r
                seen.setdefault(ident)
            per_list.append(d)
        order = list(seen)
        limit_of = {}
        if isinstance(number_of_matches, ColumnReference):
            for r in query_table.rows:
                limit_of[(r.key, r.time, r.diff)] = r.values[number_of_matches.name]
        replies, limits = [[] for _ in tables], []
        for ident in order:
            for l, d in enumerate(per_list):
                row = d.get(ident)
                reply = row.values[_INDEX_REPLY] if row is not None else ()
                # an Error reply contributes no candidates (its error is already in the retriever's log)
                replies[l].append([] if reply is ERROR else [m for m, _ in reply])
            limits.append(number_of_matches if isinstance(number_of_matches, int) else limit_of[ident])
        fused = rrf_fuse(replies, limits, self.k) if order else []
        rows = [Row(ident[0], {_INDEX_REPLY: tuple(f)}, ident[1], ident[2])
                for ident, f in zip(order, fused) if f]
        return Table(rows, errors)

    def query(self, query_column: ColumnReference, *, number_of_matches=3, metadata_filter=None) -> Table:
        tables = [r.query(query_column, number_of_matches=number_of_matches, metadata_filter=metadata_filter)
                  for r in self.retrievers]
        return self._combine_results(tables, query_column.table, number_of_matches)

    def query_as_of_now(self, query_column: ColumnReference, *, number_of_matches=3, metadata_filter=None,
                        **retriever_kwargs) -> Table:
        tables = [r.query_as_of_now(query_column, number_of_matches=number_of_matches,
                                    metadata_filter=metadata_filter, **retriever_kwargs)
                  for r in self.retrievers]
        return self._combine_results(tables, query_column.table, number_of_matches)


class HybridIndexFactory(InnerIndexFactory):
    """hybrid_index.py:161-190"""

    def __init__(self, retriever_factories: list, k: float = 60):
        if len(retriever_factories) < 2:
            raise ValueError(
                "HybridIndexFactory requires at least two retriever factories to be provided during initialization"
            )
        self.retriever_factories = retriever_factories
        self.k = k

    def build_inner_index(self, data_column: ColumnReference, metadata_column: ColumnReference | None = None):
        retrievers = [f.build_inner_index(data_column, metadata_column) for f in self.retriever_factories]
        return HybridIndex(retrievers, self.k)
