"""Pins the CPU oracle (numpy and C restatements) against the reference's own known-answer tests
and, through them, the product's host logic (operator emulation, IndexDerivedImpl,
DerivedFilteredSearchIndex).  No GPU."""

import numpy as np
import pytest

from golden import kats
from kat_harness import run_case
from oracle_adapter import OracleInner
from pathway_b200.debug import ref_scalar, table_from_rows
from pathway_b200.engine import BruteForceKnnMetricKind
from pathway_b200.stdlib.indexing import BruteForceKnn, BruteForceKnnFactory, DefaultKnnFactory


@pytest.mark.parametrize("impl", ["numpy", "c"])
@pytest.mark.parametrize("name", sorted(kats.CASES))
def test_reference_kat(name, impl):
    inner = run_case(name, lambda **kw: OracleInner(impl=impl, **kw))
    if name == "space_shenanigans_resize":  # reserved 1 -> 6 -> max(12, 9)
        assert inner.current_allocated == 12
    if name == "resize_after_delete":  # 1 -> 6 -> 12, delete 3 (no shrink: 4*6 >= 12), add 3
        assert inner.current_size == 9 and inner.current_allocated == 12


def _stream_tables(as_numpy):
    conv = (lambda c: np.array(c)) if as_numpy else (lambda c: tuple(float(x) for x in c))
    points = table_from_rows(["coords"], [(conv(c), t, 1) for c, t in kats.STREAM_POINTS], is_stream=True)
    queries = table_from_rows(["coords"], [(conv(c), t, 1) for c, t in kats.STREAM_QUERIES], is_stream=True)
    return points, queries


@pytest.mark.parametrize("impl", ["numpy", "c"])
@pytest.mark.parametrize("as_numpy", [True, False])
def test_asof_now_stream(impl, as_numpy):
    """test_index.py:1104-1147 (BruteForce, L2SQ): int numpy arrays and float tuples both unpack."""
    points, queries = _stream_tables(as_numpy)
    index = BruteForceKnn(data_column=points.coords, dimensions=2, reserved_space=3,
                          metric=BruteForceKnnMetricKind.L2SQ)
    res = index.query_as_of_now(queries.coords, number_of_matches=2,
                                _inner_override=OracleInner(2, 3, 1024 * 128, BruteForceKnnMetricKind.L2SQ, impl))
    assert res.error_log == []
    coords_of = {r.key: tuple(int(x) for x in r.values["coords"]) for r in points.rows}
    got = {tuple(int(x) for x in r.values["coords"]): {coords_of[k] for k, _ in r.values["_pw_index_reply"]}
           for r in res.rows}
    assert got == kats.STREAM_EXPECTED_K2


@pytest.mark.parametrize("impl", ["numpy", "c"])
def test_multi_removal(impl):
    """test_index.py:1150-1190: -1/+1 of the same key inside one timestamp, reserved_space=0."""
    data = table_from_rows(["a", "v"], [(a, np.array(v), t, d) for a, v, t, d in kats.MULTI_REMOVAL_DATA],
                           id_from=["a"], is_stream=True)
    queries = table_from_rows(["a", "v"], [(a, np.array(v), t, d) for a, v, t, d in kats.MULTI_REMOVAL_QUERIES],
                              id_from=["a"], is_stream=True)
    index = BruteForceKnn(data_column=data.v, dimensions=3, reserved_space=0, metric=BruteForceKnnMetricKind.COS)
    res = index.query_as_of_now(queries.v, number_of_matches=1,
                                _inner_override=OracleInner(3, 0, 1024 * 128, BruteForceKnnMetricKind.COS, impl))
    assert res.error_log == []
    got = {r.values["a"]: [k for k, _ in r.values["_pw_index_reply"]] for r in res.rows}
    assert got == {q: [ref_scalar(d)] for q, d in kats.MULTI_REMOVAL_EXPECTED.items()}


def fake_embedder(x: str) -> list[float]:  # test_index.py:790-793
    return [0.0, 1.0, float(ord(x[0])) / 5.0]


@pytest.mark.parametrize("factory", [
    BruteForceKnnFactory(dimensions=3, reserved_space=3, metric=BruteForceKnnMetricKind.COS, embedder=fake_embedder),
    BruteForceKnnFactory(reserved_space=3, metric=BruteForceKnnMetricKind.COS, embedder=fake_embedder),
    DefaultKnnFactory(dimensions=3, embedder=fake_embedder),
])
def test_index_factory_and_empty_index(factory):
    """test_index.py:797-855 (query 'a' -> doc 'a') and :1061-1101 (empty index -> empty reply)."""
    query = table_from_rows(["query"], [("a",)])
    docs = table_from_rows(["doc"], [("a",), ("b",), ("c",)])
    inner = factory.build_inner_index(docs.doc)
    res = inner.query_as_of_now(query.query, number_of_matches=1,
                                _inner_override=OracleInner(3, factory.reserved_space, 1, factory.metric))
    doc_of = {r.key: r.values["doc"] for r in docs.rows}
    assert [[doc_of[k] for k, _ in r.values["_pw_index_reply"]] for r in res.rows] == [["a"]]
    empty = table_from_rows(["doc"], [])
    inner = factory.build_inner_index(empty.doc)
    res = inner.query_as_of_now(query.query, number_of_matches=1,
                                _inner_override=OracleInner(3, factory.reserved_space, 1, factory.metric))
    assert [r.values["_pw_index_reply"] for r in res.rows] == [()]
    with pytest.raises(NotImplementedError):
        inner.query(query.query)


# ----------------------------------------------------------------------------------------------- usearch API mirror
def _usearch_distances(inner_for):
    from pathway_b200.stdlib.indexing import UsearchKnnFactory

    emb = lambda x: [float(v) for v in kats.USEARCH_EMBED.get(x[0], [1, 1, 1])]  # noqa: E731
    factory = UsearchKnnFactory(embedder=emb)  # dimensions probed from the embedder, metric COS (reference default)
    query = table_from_rows(["query"], [("a",)])
    docs = table_from_rows(["doc"], [("b",), ("c",), ("d",)])
    inner = factory.build_inner_index(docs.doc)
    assert inner.dimensions == 3
    res = inner.query_as_of_now(query.query, number_of_matches=3, _inner_override=inner_for(3))
    doc_of = {r.key: r.values["doc"] for r in docs.rows}
    got = {doc_of[k]: -s for k, s in res.rows[0].values["_pw_index_reply"]}
    assert set(got) == {"b", "c", "d"}
    for doc, want in kats.USEARCH_DISTANCES.items():
        assert np.isclose(got[doc], want, rtol=1e-5, atol=0.0), (doc, got[doc], want)
    assert [doc_of[k] for k, _ in res.rows[0].values["_pw_index_reply"]] == ["c", "b", "d"]


def test_usearch_api_mirror_distances_oracle():
    """test_index.py:858-900 through the USearchKnn mirror (exact index behind the usearch API)."""
    _usearch_distances(lambda d: OracleInner(d, 400, 1024 * 128, 1))


def test_usearch_metric_kinds():
    from oracle import brute_force_oracle as bfo
    from pathway_b200.engine import ExternalIndexFactory, USearchMetricKind

    f = ExternalIndexFactory.usearch_knn_factory(dimensions=3, reserved_space=10, metric=USearchMetricKind.L2SQ,
                                                 connectivity=0, expansion_add=0, expansion_search=0)
    assert f.params["metric"] == 0
    with pytest.raises(NotImplementedError):
        ExternalIndexFactory.usearch_knn_factory(dimensions=3, reserved_space=10, metric=USearchMetricKind.HAMMING,
                                                 connectivity=0, expansion_add=0, expansion_search=0)
    ip = bfo.BruteForceKNNIndex(2, 4, 100, bfo.IP)
    ip.add([(1, [1.0, 0.0]), (2, [3.0, 0.0]), (3, [-1.0, 0.0])])
    (_k, res), = ip.search([(0, [2.0, 5.0], 3)])
    assert res == [(2, 5.0), (1, 1.0), (3, -3.0)]  # score = -(1 - x.q)
