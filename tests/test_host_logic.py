"""Host-side mirror of the reference's index framework (no GPU): IndexDerivedImpl,
DerivedFilteredSearchIndex (JMESPath + globmatch + k-doubling), batch_by_time / operator ordering,
error-row conventions.  Innermost index = CPU oracle.  Reference: src/external_integration/mod.rs,
src/engine/dataflow/operators/external_index.rs, tests/ml/test_index.py:608-723,
tests/external_index/test_usearch_knn.py:29-97."""

import numpy as np
import pytest

from golden import kats
from oracle_adapter import OracleInner
from pathway_b200.debug import ref_scalar
from pathway_b200.engine import ERROR, BruteForceKnnMetricKind, ExternalIndexFactory
from pathway_b200.external_index import (AddDataEntry, DerivedFilteredSearchIndex, ErrorInIndexSearch,
                                         ErrorInIndexUpdate, IndexDerivedImpl, MissingKey, QueryEntry, TypeMismatch,
                                         batch_by_time, unpack_vec_f64, use_external_index_as_of_now)

COS, L2SQ = BruteForceKnnMetricKind.COS, BruteForceKnnMetricKind.L2SQ


def _filtered(metric=L2SQ, dims=3):
    return DerivedFilteredSearchIndex(OracleInner(dims, 4, 1024 * 128, metric))


def test_filter_globmatch_and_equality_counts():
    """test_usearch_knn.py:29-97 semantics: glob on a path, JMESPath equality; k-doubling finds matches
    that are not among the first k nearest."""
    ix = _filtered()
    ix.add([AddDataEntry(ref_scalar(pk), vec, meta) for pk, vec, meta in kats.FILTER_INDEX])
    q = (0.0, 0.1, 0.1)
    res = dict(ix.search([
        QueryEntry(1, q, 2, "globmatch('foo/bar/*', path)"),
        QueryEntry(2, q, 2, "k == `2`"),
        QueryEntry(3, q, 1, "globmatch('**/*.txt', path)"),
        QueryEntry(4, q, 3, "contains(path, 'nothing')"),
        QueryEntry(5, q, 2, None),
    ]))
    ids = lambda r: [k for k, _ in r]  # noqa: E731
    assert ids(res[1]) == [ref_scalar(1), ref_scalar(3)]
    assert ids(res[2]) == [ref_scalar(2), ref_scalar(4)]
    assert ids(res[3]) == [ref_scalar(3)]
    assert res[4] == ()
    assert ids(res[5]) == [ref_scalar(1), ref_scalar(2)]


def test_filter_errors_are_per_query():
    ix = _filtered()
    ix.add([AddDataEntry(1, (1.0, 0.0, 0.0), {"a": 1}), AddDataEntry(2, (0.0, 1.0, 0.0), {"a": 2})])
    res = dict(ix.search([
        QueryEntry(1, (1.0, 0.0, 0.0), 1, "a =="),          # bad JMESPath -> error reply (test_index.py:666-723)
        QueryEntry(2, (1.0, 0.0, 0.0), 1, "a"),             # not a boolean (mod.rs:447-453)
        QueryEntry(3, (1.0, 0.0, 0.0), -1, None),           # negative limit: usize::try_from fails
        QueryEntry(4, (1, 0.0, 0.0), 1, None),              # Int inside a tuple: TypeMismatch (mod.rs:345-358)
        QueryEntry(5, (1.0, 0.0, 0.0), None, None),         # missing limit column -> 1 (mod.rs:417-421)
        QueryEntry(6, (1.0, 0.0, 0.0), 0, None),
    ]))
    assert all(isinstance(res[i], Exception) for i in (1, 2, 3, 4))
    assert isinstance(res[4], TypeMismatch)
    assert [k for k, _ in res[5]] == [1] and res[6] == ()


@pytest.mark.parametrize("pushdown", [False, True])
@pytest.mark.parametrize("bad", ["a**b", "[abc", "[z-a]"])
def test_malformed_glob_fails_only_its_query(bad, pushdown):
    """A glob::PatternError becomes a JmespathError of that query (mod.rs:270-277) -> its reply is an error
    (mod.rs:432-455, :480-484); the other queries of the same timestamp are answered."""
    from oracle_adapter import OracleInnerWithFilters

    ix = DerivedFilteredSearchIndex(OracleInnerWithFilters(3, 4, 1024 * 128, L2SQ), pushdown=True) if pushdown else _filtered()
    ix.add([AddDataEntry(1, (1.0, 0.0, 0.0), {"path": "a/b.txt"}), AddDataEntry(2, (0.0, 1.0, 0.0), {"path": "c/d.md"})])
    res = dict(ix.search([
        QueryEntry(1, (1.0, 0.0, 0.0), 1, f"globmatch('{bad}', path)"),
        QueryEntry(2, (1.0, 0.0, 0.0), 1, "globmatch('**/*.md', path)"),
        QueryEntry(3, (1.0, 0.0, 0.0), 1, None),
    ]))
    assert isinstance(res[1], Exception)
    assert [k for k, _ in res[2]] == [2] and [k for k, _ in res[3]] == [1]


def test_unpack_encodings():
    assert list(unpack_vec_f64(np.array([1, 2, 3]))) == [1.0, 2.0, 3.0]          # Value::IntArray
    assert list(unpack_vec_f64(np.array([1.5, 2], dtype=np.float32))) == [1.5, 2.0]  # FloatArray (widened)
    assert unpack_vec_f64((1.0, 2.5)) == [1.0, 2.5]
    with pytest.raises(TypeMismatch):
        unpack_vec_f64((1.0, 2))
    with pytest.raises(TypeMismatch):
        unpack_vec_f64("1,2")


def test_take_updates_remove_then_add_and_error_rows():
    """mod.rs:110-151: diff==0 dropped, Error cells logged and skipped, removes before adds."""
    log = []
    inner = _filtered(COS)
    ix = IndexDerivedImpl(inner, log, lambda r: r["v"], None, lambda r: r["v"], lambda r: r["k"], None)
    ix.take_updates([(1, {"v": (1.0, 0.0, 0.0)}, 1), (2, {"v": (0.0, 1.0, 0.0)}, 1), (3, {"v": ERROR}, 1),
                     (4, {"v": (0.0, 0.0, 1.0)}, 0)])
    assert len(log) == 1 and isinstance(log[0], ErrorInIndexUpdate)
    assert inner.inner.current_size == 2
    # same key retracted and re-added in one batch, listed add-first: the remove still happens first
    ix.take_updates([(1, {"v": (0.0, 0.0, 1.0)}, 1), (1, {"v": (1.0, 0.0, 0.0)}, -1), (9, {"v": (1.0, 1.0, 1.0)}, -1)])
    assert inner.inner.current_size == 2
    assert isinstance(log[-1], MissingKey)  # key 9 never existed: logged, nothing else changes
    out = ix.search([(10, {"v": (0.0, 0.0, 1.0), "k": 1}, 1), (11, {"v": ERROR, "k": 1}, 1)])
    assert out[0][0] == 10 and [k for k, _ in out[0][1][1]] == [1] and out[0][2] == 1
    assert out[1][1][1] is ERROR and isinstance(log[-1], ErrorInIndexSearch)


def test_operator_times_are_ascending_and_as_of_now():
    """external_index.rs:119-153: updates of time t are visible to queries of time t, never later ones;
    one reply per query row, same key/time/diff; consolidation cancels +1/-1 of the same row."""
    inner = _filtered(L2SQ, dims=1)
    ix = IndexDerivedImpl(inner, [], lambda r: r["v"], None, lambda r: r["v"], lambda r: r["k"], None)
    index_stream = [(5, {"v": (5.0,)}, 4, 1), (1, {"v": (1.0,)}, 2, 1), (1, {"v": (1.0,)}, 6, -1),
                    (7, {"v": (7.0,)}, 2, 1), (7, {"v": (7.0,)}, 2, -1)]
    queries = [(100, {"v": (0.0,), "k": 5}, 6, 1), (101, {"v": (0.0,), "k": 5}, 2, 1),
               (102, {"v": (0.0,), "k": 5}, 1, 1), (103, {"v": (0.0,), "k": 5}, 4, 1)]
    out = {k: ([m for m, _ in reply], t, d) for k, (_row, reply), t, d in
           use_external_index_as_of_now(index_stream, queries, ix)}
    assert out == {102: ([], 1, 1), 101: ([1], 2, 1), 103: ([1, 5], 4, 1), 100: ([5], 6, 1)}
    bt = batch_by_time(index_stream)
    assert list(bt) == [2, 4, 6] and [k for k, _, _ in bt[2]] == [1]


def test_factory_signature_matches_reference():
    """engine.pyi:1080-1086: keyword-only dimensions/reserved_space/auxiliary_space/metric."""
    f = ExternalIndexFactory.brute_force_knn_factory(dimensions=3, reserved_space=10, auxiliary_space=1000,
                                                     metric=COS)
    with pytest.raises(TypeError):
        ExternalIndexFactory.brute_force_knn_factory(3, 10, 1000, COS)
    ix = f.make_instance(inner_override=OracleInner(3, 10, 1000, COS))
    assert isinstance(ix, DerivedFilteredSearchIndex)


def test_pushdown_host_logic_matches_the_lazy_loop():
    """DerivedFilteredSearchIndex(pushdown=True) over a CPU stand-in with the filter surface: same replies as the
    reference's k-doubling loop; masks of cached expressions are extended when rows arrive; the cache is LRU-capped
    and never evicts a filter the current batch uses; non-boolean expressions fall back to the lazy loop."""
    from oracle_adapter import OracleInnerWithFilters

    rng = np.random.default_rng(3)
    inner = OracleInnerWithFilters(4, 4, 1024 * 128, L2SQ)
    push = DerivedFilteredSearchIndex(inner, pushdown=True)
    push.MAX_PUSHED_FILTERS = 2
    loop = DerivedFilteredSearchIndex(OracleInner(4, 4, 1024 * 128, L2SQ), pushdown=False)
    exprs = ["k == `1`", "globmatch('a/*', p)", "k", "k >= `2`", None]
    nxt = 0
    for step in range(4):
        batch = [AddDataEntry(nxt + i, tuple(rng.standard_normal(4)), {"k": (nxt + i) % 3, "p": f"{'ab'[(nxt + i) % 2]}/x{i}"})
                 for i in range(60)]
        rm = list(range(max(0, nxt - 25), nxt - 5)) if step else []
        for ix in (push, loop):
            ix.remove(rm)
            ix.add([AddDataEntry(e.key, e.data, e.filter_data) for e in batch])
        nxt += 60
        use = exprs if step % 2 == 0 else ["k >= `2`", "k == `0`", None]  # a different working set -> eviction
        qs = [QueryEntry(100 * step + i, tuple(rng.standard_normal(4)), int(rng.integers(0, 6)), use[i % len(use)])
              for i in range(25)]
        got, want = dict(push.search(qs)), dict(loop.search(qs))
        for k in want:
            if isinstance(want[k], Exception):
                assert isinstance(got[k], Exception)
            else:
                assert [m for m, _ in got[k]] == [m for m, _ in want[k]], (step, k)
                np.testing.assert_allclose([x for _, x in got[k]], [x for _, x in want[k]], rtol=1e-12, atol=1e-14)
        assert len(push._pushed) <= 2 and all(fid in inner.filters for fid, _ in push._pushed.values())
    assert inner.created >= 3  # eviction + re-creation happened


def test_public_import_paths():
    """python/pathway/__init__.py:79,99 and stdlib/indexing/__init__.py:7,39: `pw.indexing` is `pw.stdlib.indexing`,
    the metric enum is reachable from both `pw.indexing` and `pw.engine`."""
    import pathway_b200 as pw

    assert pw.indexing is pw.stdlib.indexing
    assert pw.indexing.BruteForceKnnMetricKind is pw.engine.BruteForceKnnMetricKind
    for name in ("BruteForceKnn", "BruteForceKnnFactory", "DefaultKnnFactory", "USearchKnn", "UsearchKnnFactory",
                 "DataIndex", "HybridIndex", "HybridIndexFactory", "InnerIndexFactory", "USearchMetricKind"):
        assert hasattr(pw.indexing, name), name
    for name in ("ExternalIndexFactory", "ExternalIndexData", "ExternalIndexQuery", "Scope"):
        assert hasattr(pw.engine, name), name


def test_placement_weights_stay_inside_the_band():
    """clamp_weights: mean 1.0 and every weight within +-25 % of it, also when one clip + renormalise would leave the
    band again ([0.5, 1.4, 1.4, 1.4] -> 0.69 in round 1); unknown rates (no profiled launch) get the mean speed."""
    from pathway_b200.placement import clamp_weights

    for rates in ([0.5, 1.4, 1.4, 1.4], [1.0, 100.0], [3.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0], [0.0, 2.0, 2.0], [0.0, 0.0]):
        w = np.array(clamp_weights(rates, damping=1.0))
        assert abs(w.mean() - 1.0) < 1e-9 and w.min() >= 0.75 - 1e-9 and w.max() <= 1.25 + 1e-9, (rates, w)
    assert clamp_weights([0.0, 2.0, 2.0]) == [1.0, 1.0, 1.0]


def test_operator_coalescing_of_update_batches_keeps_every_reply():
    """SURVEY.md 8(f-4) emulated: update batches of consecutive timestamps without queries are applied as ONE netted
    batch (a key added and retracted in between never reaches the index, upserts collapse to their last value); every
    as-of-now reply equals the uncoalesced operator's (external_index.rs:119-153), with far fewer index calls."""
    rng = np.random.default_rng(17)
    d = 6

    def make():
        inner = _filtered(L2SQ, dims=d)
        return IndexDerivedImpl(inner, [], lambda r: r["v"], None, lambda r: r["v"], lambda r: r["k"], None)

    index_stream, queries, live, nxt = [], [], {}, 0
    for t in range(2, 400, 2):
        for _ in range(int(rng.integers(0, 6))):
            roll = rng.random()
            if roll < 0.5 or not live:  # new key
                row = {"v": tuple(float(x) for x in rng.standard_normal(d))}
                index_stream.append((nxt, row, t, 1)); live[nxt] = row; nxt += 1
            elif roll < 0.75:  # upsert of a live key: retract the old row, insert the new one, same timestamp
                k = int(rng.choice(list(live)))
                row = {"v": tuple(float(x) for x in rng.standard_normal(d))}
                index_stream.append((k, live[k], t, -1)); index_stream.append((k, row, t, 1)); live[k] = row
            else:  # removal
                k = int(rng.choice(list(live)))
                index_stream.append((k, live.pop(k), t, -1))
        if rng.random() < 0.08:  # queries are rare: long runs of update-only timestamps
            for j in range(3):
                queries.append((10_000 + len(queries), {"v": tuple(float(x) for x in rng.standard_normal(d)), "k": 4}, t, 1))
    s_plain, s_merged = {}, {}
    plain = use_external_index_as_of_now(index_stream, queries, make(), stats=s_plain)
    merged = use_external_index_as_of_now(index_stream, queries, make(), coalesce=True, stats=s_merged)
    assert len(plain) == len(merged) == len(queries) > 10
    for (k1, (_r1, rep1), t1, d1), (k2, (_r2, rep2), t2, d2) in zip(plain, merged):
        assert (k1, t1, d1) == (k2, t2, d2)
        assert [m for m, _ in rep1] == [m for m, _ in rep2]
        np.testing.assert_allclose([x for _, x in rep1], [x for _, x in rep2], rtol=1e-12)
    assert s_plain["update_calls"] == s_plain["timestamps_with_updates"]
    assert s_merged["update_calls"] * 4 < s_plain["update_calls"]           # far fewer add / remove round trips
    assert s_merged["update_rows"] < s_merged["update_rows_uncoalesced"]     # and rows that never reach the index
