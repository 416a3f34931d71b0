"""`KNNIndex` facade (BASELINE.json configs[0]) over the brute-force index: the reference's own known answers
(stdlib/ml/index.py:222-247 doctest, tests/ml/test_index.py:352-389 `test_asof_now_with_variable_k`), CPU (host logic
over the oracle) and GPU."""

import numpy as np
import pytest

from oracle_adapter import OracleInner
from pathway_b200.debug import table_from_markdown
from pathway_b200.engine import BruteForceKnnMetricKind
from harness.knn_index_harness import KNNIndex


def _coords(t):
    return t.with_columns(coords=lambda r: (float(r["x"]), float(r["y"])))


def _maker(metric=BruteForceKnnMetricKind.L2SQ):
    return lambda: OracleInner(2, 400, 1024 * 128, metric)


def _doctest_case(extra):
    data = _coords(table_from_markdown("""
         x | y | __time__
         2 | 3 |     2
         0 | 0 |     2
         2 | 2 |     6
        -3 | 3 |    10
    """))
    data = type(data)([type(r)(r.key, {"coords": r.values["coords"]}, r.time, r.diff) for r in data.rows])
    queries = _coords(table_from_markdown("""
         x | y | __time__ | __diff__
         1 | 1 |     4    |     1
        -3 | 1 |     8    |     1
    """))
    index = KNNIndex(data.coords, data, n_dimensions=2)
    ans = index.get_nearest_items_asof_now(queries.coords, k=2, **extra)
    got = {q.values["coords"]: a.values["coords"] for q, a in zip(queries.rows, ans.rows)}
    assert got == {(1.0, 1.0): ((0.0, 0.0), (2.0, 3.0)), (-3.0, 1.0): ((0.0, 0.0), (2.0, 2.0))}
    assert [(a.time, a.diff) for a in ans.rows] == [(4, 1), (8, 1)]


def _variable_k_case(extra):
    points = _coords(table_from_markdown("""
         x |  y | __time__
         2 |  2 |     2
         3 | -2 |     4
        -1 |  0 |     8
         1 |  2 |    12
        -3 |  1 |    16
         1 | -4 |    20
    """))
    points = type(points)([type(r)(r.key, {"coords": r.values["coords"]}, r.time, r.diff) for r in points.rows])
    queries = _coords(table_from_markdown("""
         x |  y | k | __time__
         0 |  0 | 1 |     6
         2 | -2 | 2 |    10
        -1 |  1 | 3 |    14
        -2 | -3 | 0 |    18
    """))
    index = KNNIndex(points.coords, points, n_dimensions=2, n_and=5)
    ans = index.get_nearest_items_asof_now(queries.coords, queries.k, with_distances=True, **extra)
    got = {q.values["coords"]: tuple(sorted(a.values["coords"])) for q, a in zip(queries.rows, ans.rows)}
    assert got == {
        (0.0, 0.0): ((2.0, 2.0),),
        (2.0, -2.0): ((-1.0, 0.0), (3.0, -2.0)),
        (-1.0, 1.0): ((-1.0, 0.0), (1.0, 2.0), (2.0, 2.0)),
        (-2.0, -3.0): (),
    }
    by_q = {q.values["coords"]: a.values for q, a in zip(queries.rows, ans.rows)}
    assert by_q[(2.0, -2.0)]["coords"] == ((3.0, -2.0), (-1.0, 0.0))  # increasing squared distance: 1, 13
    np.testing.assert_allclose(by_q[(2.0, -2.0)]["dist"], (1.0, 13.0), rtol=1e-12)
    flat = index.get_nearest_items_asof_now(queries.coords, queries.k, collapse_rows=False, **extra)
    assert len(flat.rows) == 1 + 2 + 3 and all("query_id" in r.values for r in flat.rows)


def test_doctest_kat_host_logic():
    _doctest_case({"_inner_override": _maker()})


def test_asof_now_with_variable_k_host_logic():
    _variable_k_case({"_inner_override": _maker()})


def test_facade_errors():
    t = _coords(table_from_markdown("""
         x | y
         1 | 2
    """))
    with pytest.raises(ValueError, match="allowed values are 'euclidean' and 'cosine'"):
        KNNIndex(t.coords, t, n_dimensions=2, distance_type="manhattan")
    with pytest.raises(NotImplementedError, match="as-of-now"):
        KNNIndex(t.coords, t, n_dimensions=2).get_nearest_items(t.coords)


@pytest.mark.gpu
def test_doctest_kat_gpu():
    _doctest_case({})


@pytest.mark.gpu
def test_asof_now_with_variable_k_gpu():
    _variable_k_case({})


@pytest.mark.gpu
def test_config0_shape_1k_x_64_k3_gpu():
    """BASELINE.json configs[0]: KNNIndex k = 3 on a 1 k x 64 random static table, against the oracle."""
    from oracle import brute_force_oracle as bfo
    from pathway_b200.debug import table_from_rows

    rng = np.random.default_rng(0)
    emb = bfo.bf16_round(rng.standard_normal((1000, 64)))
    qv = bfo.bf16_round(rng.standard_normal((100, 64)))
    data = table_from_rows(["doc", "emb"], [(i, emb[i].tolist()) for i in range(1000)], id_from=["doc"])
    queries = table_from_rows(["q", "emb"], [(j, qv[j].tolist()) for j in range(100)], id_from=["q"])
    got = KNNIndex(data.emb, data, n_dimensions=64, distance_type="cosine").get_nearest_items_asof_now(
        queries.emb, k=3, with_distances=True)
    ora = bfo.BruteForceKNNIndex(64, 400, 1024 * 128, bfo.COS)
    ora.add([(i, emb[i]) for i in range(1000)])
    want = ora.search([(j, qv[j], 3) for j in range(100)])
    for row, (_, w) in zip(got.rows, want):
        assert list(row.values["doc"]) == [m for m, _ in w]
        np.testing.assert_allclose(row.values["dist"], [-s for _, s in w], rtol=1e-9, atol=1e-12)
