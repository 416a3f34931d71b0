"""reciprocal-rank fusion (`HybridIndex`) and reply repacking (`DataIndex`) behind the operator.

Known answers are transcribed from /root/reference/python/pathway/tests/ml/test_index.py:954-1058
(`test_hybrid_index`, `test_hybrid_index_ignores_duplicates`); the checker is oracle/hybrid_oracle.py."""

import ctypes as C

import numpy as np
import pytest

from oracle import hybrid_oracle as ho
from oracle_adapter import OracleInner
from pathway_b200 import _native
from pathway_b200.debug import table_from_rows
from pathway_b200.engine import BruteForceKnnMetricKind
from pathway_b200.external_index import keys_to_array
from pathway_b200.stdlib.indexing import (BruteForceKnn, BruteForceKnnFactory, DataIndex, HybridIndex,
                                          HybridIndexFactory)
from pathway_b200.stdlib.indexing.data_index import _SCORE
from pathway_b200.stdlib.indexing.hybrid_index import rrf_fuse


def _random_replies(rng, nq, n_lists, universe, kmax):
    per_list = []
    for _ in range(n_lists):
        replies = []
        for _q in range(nq):
            cnt = int(rng.integers(0, kmax + 1))
            ids = rng.choice(universe, size=min(cnt, universe), replace=False)
            # 128-bit keys: spread the high half too
            replies.append([int(i) | (int(i) * 7919 % 5) << 64 for i in ids])
        per_list.append(replies)
    return per_list


@pytest.mark.parametrize("n_lists,k_rrf", [(2, 60.0), (3, 2.0), (5, 0.5)])
def test_rrf_fuse_matches_oracle(n_lists, k_rrf):
    rng = np.random.default_rng(n_lists)
    nq = 200
    per_list = _random_replies(rng, nq, n_lists, universe=40, kmax=25)
    limits = [int(x) for x in rng.integers(0, 30, size=nq)]
    got = rrf_fuse(per_list, limits, k_rrf)
    as_dicts = [{q: [(m, 0.0) for m in replies[q]] for q in range(nq)} for replies in per_list]
    want = ho.combine_results(as_dicts, dict(enumerate(limits)), k_rrf)
    for q in range(nq):
        w = list(want.get(q, ()))
        assert [m for m, _ in got[q]] == [m for m, _ in w], q
        assert [s for _, s in got[q]] == [s for _, s in w], q  # same summation order => same bits


def test_rrf_fuse_reads_search_layout():
    """A list laid out like pwknn_search's output (stride = limit, count <= limit) is consumed in place."""
    lib = _native.lib()
    nq = 3
    lim_a = np.array([4, 2, 3], dtype=np.uint32)
    cnt_a = np.array([2, 2, 0], dtype=np.uint32)
    keys_a = np.zeros((int(lim_a.sum()), 2), dtype=np.uint64)
    keys_a[0:2] = keys_to_array([11, 12])
    keys_a[4:6] = keys_to_array([21, 22])
    cnt_b = np.array([1, 2, 1], dtype=np.uint32)  # packed
    keys_b = keys_to_array([12, 22, 23, 31])
    limits = np.array([3, 3, 3], dtype=np.uint32)
    out_keys = np.zeros((9, 2), dtype=np.uint64)
    out_scores = np.zeros(9)
    out_counts = np.zeros(3, dtype=np.uint32)
    vp = C.c_void_p * 2
    rc = lib.pwknn_rrf_fuse(2, nq, vp(keys_a.ctypes.data, keys_b.ctypes.data), vp(cnt_a.ctypes.data, cnt_b.ctypes.data),
                            vp(lim_a.ctypes.data, None), 60.0, limits.ctypes.data, out_keys.ctypes.data,
                            out_scores.ctypes.data, out_counts.ctypes.data)
    assert rc == 0
    assert out_counts.tolist() == [2, 3, 1]
    assert out_keys[0:2, 0].tolist() == [12, 11]  # 12: 1/62 + 1/61, 11: 1/61
    assert out_scores[0] == 1 / 62 + 1 / 61 and out_scores[1] == 1 / 61
    assert out_keys[3:6, 0].tolist() == [22, 21, 23]  # 22: 1/62+1/61; 21: 1/61; 23: 1/62
    assert out_keys[6, 0] == 31
    # count > limit is rejected, not read out of bounds
    bad = np.array([5, 2, 0], dtype=np.uint32)
    rc = lib.pwknn_rrf_fuse(2, nq, vp(keys_a.ctypes.data, keys_b.ctypes.data), vp(bad.ctypes.data, cnt_b.ctypes.data),
                            vp(lim_a.ctypes.data, None), 60.0, limits.ctypes.data, out_keys.ctypes.data,
                            out_scores.ctypes.data, out_counts.ctypes.data)
    assert rc == _native.E_INVALID and b"count > limit" in lib.pwknn_last_global_error()


def test_needs_two_retrievers():
    with pytest.raises(ValueError, match="at least two indices"):
        HybridIndex([object()])
    with pytest.raises(ValueError, match="at least two retriever factories"):
        HybridIndexFactory([BruteForceKnnFactory(dimensions=3)])


def _embedder(close: str, near: str):
    def embed(x: str):
        if x == "query" or x == close:
            return [1.0, 2.0, 3.0]
        if x == near:
            return [1.0, 2.0, 4.0]
        return [4.0, 5.0, 6.0]
    return embed


def _hybrid_kat(inner_maker, second_embedder, k=None):
    query = table_from_rows(["query"], [("query",)])
    docs = table_from_rows(["doc"], [("doc1",), ("doc2",), ("doc3",)])
    kw = dict(dimensions=3, reserved_space=3, metric=BruteForceKnnMetricKind.COS)
    index1 = BruteForceKnn(data_column=docs.doc, metadata_column=None, embedder=_embedder("doc1", "doc2"), **kw)
    index2 = BruteForceKnn(data_column=docs.doc, metadata_column=None, embedder=second_embedder, **kw)
    hybrid = HybridIndex([index1, index2], k=k) if k is not None else HybridIndex([index1, index2])
    index = DataIndex(docs, hybrid)
    extra = {"_inner_override": inner_maker} if inner_maker else {}
    res = index.query_as_of_now(query.query, collapse_rows=True, number_of_matches=2, **extra)
    assert res.error_log == []
    (row,) = res.rows
    return row.values


def _oracle_maker():
    return OracleInner(3, 3, 1024 * 128, BruteForceKnnMetricKind.COS)


def _check_hybrid_index(inner_maker):
    """test_index.py:954-1011: retriever 1 ranks doc1, doc2; retriever 2 ranks doc2, doc3; k = 2."""
    v = _hybrid_kat(inner_maker, _embedder("doc2", "doc3"), k=2)
    assert v["query"] == "query"
    assert v["doc"] == ("doc2", "doc1")
    assert v[_SCORE] == (1 / 3 + 1 / 4, 1 / 3)


def _check_ignores_duplicates(inner_maker):
    """test_index.py:1014-1058: both retrievers rank doc1, doc3 -> the pair once."""
    emb = lambda x: [1.0, 2.0, 3.0] if x in ("query", "doc1") else ([1.0, 2.0, 4.0] if x == "doc3" else [4.0, 5.0, 6.0])  # noqa: E731
    query = table_from_rows(["query"], [("query",)])
    docs = table_from_rows(["doc"], [("doc1",), ("doc2",), ("doc3",)])
    kw = dict(dimensions=3, reserved_space=3, metric=BruteForceKnnMetricKind.COS, embedder=emb)
    hybrid = HybridIndex([BruteForceKnn(data_column=docs.doc, **kw), BruteForceKnn(data_column=docs.doc, **kw)])
    extra = {"_inner_override": inner_maker} if inner_maker else {}
    res = DataIndex(docs, hybrid).query_as_of_now(query.query, collapse_rows=True, number_of_matches=2, **extra)
    (row,) = res.rows
    assert sorted(row.values["doc"]) == ["doc1", "doc3"]
    assert row.values[_SCORE] == (2 / 61, 2 / 62)


def test_hybrid_index_kat_host_logic():
    _check_hybrid_index(_oracle_maker)


def test_hybrid_index_ignores_duplicates_host_logic():
    _check_ignores_duplicates(_oracle_maker)


@pytest.mark.gpu
def test_hybrid_index_kat_gpu():
    _check_hybrid_index(None)


@pytest.mark.gpu
def test_hybrid_index_ignores_duplicates_gpu():
    _check_ignores_duplicates(None)


def _stream_case(inner_maker):
    """Data table changing over time + per-query limits: DataIndex over a single KNN retriever and over a hybrid of
    two (second one sees the vectors reversed), checked against the restated dataflow."""
    rng = np.random.default_rng(5)
    d, n = 8, 60
    vecs = rng.standard_normal((n, d))
    rows = [(f"doc{i}", vecs[i].tolist(), 2 * (i // 20), 1) for i in range(n)]
    rows += [(f"doc{i}", vecs[i].tolist(), 3, -1) for i in range(0, 10)]  # retract ten docs at time 3
    docs = table_from_rows(["doc", "vec"], rows, id_from=["doc"], is_stream=True)
    qv = rng.standard_normal((12, d))
    queries = table_from_rows(["q", "vec", "lim"], [(f"q{j}", qv[j].tolist(), 1 + j % 5, 1 + (j % 6), 1)
                                                     for j in range(12)], id_from=["q"], is_stream=True)
    kw = dict(dimensions=d, reserved_space=4, metric=BruteForceKnnMetricKind.L2SQ)
    knn_a = BruteForceKnn(data_column=docs.vec, **kw)
    knn_b = BruteForceKnn(data_column=docs.vec, embedder=lambda v: list(reversed(v))[:d // 2] + [0.0] * (d - d // 2), **kw)
    extra = {"_inner_override": inner_maker} if inner_maker else {}
    columns = ["doc", "vec"]
    raw_a = knn_a.query_as_of_now(queries.vec, number_of_matches=queries.lim, **extra)
    raw_b = knn_b.query_as_of_now(queries.vec, number_of_matches=queries.lim, **extra)
    replies_a = {r.key: r.values["_pw_index_reply"] for r in raw_a.rows}
    replies_b = {r.key: r.values["_pw_index_reply"] for r in raw_b.rows}
    lim_of = {r.key: r.values["lim"] for r in queries.rows}

    def data_as_of(t):
        live = {}
        for r in sorted(docs.rows, key=lambda r: (r.time, r.diff > 0)):
            if r.time <= t:
                live.pop(r.key, None) if r.diff < 0 else live.__setitem__(r.key, r.values)
        return live

    for index, raw_want in ((knn_a, replies_a), (HybridIndex([knn_a, knn_b], k=3), None)):
        got = DataIndex(docs, index).query_as_of_now(queries.vec, number_of_matches=queries.lim, **extra)
        assert len(got.rows) == len(queries.rows)
        if raw_want is None:
            raw_want = ho.combine_results([replies_a, replies_b], lim_of, 3)
        for q, g in zip(queries.rows, got.rows):
            want = ho.repack_collapsed({q.key: raw_want.get(q.key, ())}, [q.key], data_as_of(q.time), columns)[q.key]
            assert g.key == q.key and g.values["q"] == q.values["q"]
            assert g.values["doc"] == want["doc"], q.values["q"]
            assert g.values[_SCORE] == want["_pw_index_reply_score"]
            assert len(g.values["doc"]) <= q.values["lim"]
    # flat output: one row per match, unanswered queries kept with None
    flat = DataIndex(docs, knn_a).query_as_of_now(queries.vec, number_of_matches=queries.lim, collapse_rows=False,
                                                   **extra)
    assert sum(1 for r in flat.rows if r.values["doc"] is not None) == sum(len(v) for v in replies_a.values())


def test_data_index_stream_host_logic():
    _stream_case(lambda: OracleInner(8, 4, 1024 * 128, BruteForceKnnMetricKind.L2SQ))


@pytest.mark.gpu
def test_data_index_stream_gpu():
    _stream_case(None)
