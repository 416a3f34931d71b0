"""Replays the transcribed reference KATs (tests/golden/kats.py) through the product's host logic
(pathway_b200.debug / external_index) with a pluggable innermost index: the CPU oracle in the
`-m "not gpu"` suite, the B200 index in the `-m gpu` suite.  Mirrors `get_ret` of
/root/reference/python/pathway/tests/external_index/test_brute_force_knn.py:37-60."""

from pathway_b200.debug import table_from_markdown
from pathway_b200.engine import BruteForceKnnMetricKind, ExternalIndexFactory

from golden import kats


def make_list(vector_as_str: str) -> list[float]:
    return [float(x) for x in vector_as_str.split(",")]


def get_ret(queries, index, index_factory, inner_override=None):
    raw = index._external_index_as_of_now(
        queries,
        index_column=index.data,
        query_column=queries.data,
        index_factory=index_factory,
        query_responses_limit_column=queries.limit,
        _inner_override=inner_override,
    )
    by_id = {r.key: r.values["pk_source"] for r in index.rows}
    ret = []
    for r in raw.rows:
        for matched_item_id, matched_item_score in r.values["_pw_index_reply"]:
            ret.append((r.values["pk_source"], by_id[matched_item_id], round(-matched_item_score, 2)))
    return sorted(ret), raw


def run_case(name: str, make_inner=None):
    case = kats.CASES[name]
    index = table_from_markdown(case["index"], id_from=["pk_source"]).with_columns(
        data=lambda row: make_list(row["data"]))
    queries = table_from_markdown(case["queries"], id_from=["pk_source"]).with_columns(
        data=lambda row: make_list(row["data"]))
    f = dict(case["factory"])
    metric = BruteForceKnnMetricKind[f.pop("metric")]
    factory = ExternalIndexFactory.brute_force_knn_factory(metric=metric, **f)
    inner = make_inner(metric=metric, **f) if make_inner else None
    ret, raw = get_ret(queries, index, factory, inner)
    assert raw.error_log == []
    assert ret == sorted(case["expected"]), (name, ret)
    return inner
