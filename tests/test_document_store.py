"""`DocumentStore.retrieve_query` over `BruteForceKnnFactory` (BASELINE.json configs[2]'s caller): the glob / metadata
filter known answers of /root/reference/python/pathway/xpacks/llm/tests/test_document_store.py:140-190
(`test_vectorstore_glob_filtering`), :384-430 (`test_vs_filtering_negatives`), :433-482 (`test_vs_filtering_metadata`)
and :485-530 (`test_vs_filtering_edge_cases`), CPU (host logic over the oracle, with and without filter push-down) and
GPU."""

import pytest

from oracle_adapter import OracleInner, OracleInnerWithFilters
from pathway_b200.debug import table_from_rows
from pathway_b200.engine import BruteForceKnnMetricKind
from pathway_b200.stdlib.indexing import BruteForceKnnFactory, HybridIndexFactory
from harness.document_store_harness import DocumentStore, _get_jmespath_filter

PATH = "pathway/xpacks/llm/tests/test_vector_store.py"
COS = BruteForceKnnMetricKind.COS


def fake_embeddings_model(x: str) -> list[float]:
    return [1.0, 1.0, 0.0]


def _store(path=PATH, factory=None):
    docs = table_from_rows(["data", "_metadata"], [("test".encode("utf-8"), {"path": path})])
    factory = factory or BruteForceKnnFactory(dimensions=3, reserved_space=10, embedder=fake_embeddings_model, metric=COS)
    return DocumentStore(docs, retriever_factory=factory)


def _retrieve(store, query, k, metadata_filter, glob, extra):
    q = table_from_rows(["query", "k", "metadata_filter", "filepath_globpattern"], [(query, k, metadata_filter, glob)])
    out = store.retrieve_query(q, **extra)
    assert out.error_log == []
    (row,) = out.rows
    return row.values["result"]


MAKERS = {
    "loop": lambda: OracleInner(3, 10, 1024 * 128, COS),
    "pushdown": lambda: OracleInnerWithFilters(3, 10, 1024 * 128, COS),
}

GLOB_HITS = ["", "**/*.py", PATH]
GLOB_MISSES = ["somefile.pdf", "**/*.txt", "pathway/test_vector_store.py", "src.py", "`" + PATH + "`"]
META_HITS = [
    "",
    "contains(path, `test_vector_store`)",
    'contains(path, `"test_vector_store"`)',
    "contains(path, `pathway/xpacks/llm/tests/test_vector_store.py`)",
    "path == `pathway/xpacks/llm/tests/test_vector_store.py`",
    "globmatch(`pathway/xpacks/llm/tests/test_vector_store.py`, path)",
]
EDGE_PATH = "Document Enregistrement Universel 2023 publié à l'XYZ le 28 février 2024.pdf"
EDGE_META = [
    "",
    "contains(path, `Document Enregistrement Universel 2023 publié à l'XYZ le 28 février 2024.pdf`)",
    "path == `Document Enregistrement Universel 2023 publié à l'XYZ le 28 février 2024.pdf`",
    'path == "`Document Enregistrement Universel 2023 publié à l\'XYZ le 28 février 2024.pdf"`',
    "contains(path, `Document Enregistrement`)",
]


def _check_all(extra):
    store = _store()
    for glob in GLOB_HITS:  # test_vectorstore_glob_filtering
        (hit,) = _retrieve(store, "test", 1, None, glob, extra)
        assert hit["dist"] < 1.0e-6 and hit["text"] == "test" and hit["metadata"]["path"] == PATH
    for glob in GLOB_MISSES:  # test_vs_filtering_negatives
        assert _retrieve(store, "Foo", 1, None, glob, extra) == []
    for mf in META_HITS:  # test_vs_filtering_metadata
        (hit,) = _retrieve(store, "Foo", 1, mf, None, extra)
        assert hit["dist"] < 1.0e-6 and hit["text"]
    edge = _store(EDGE_PATH)
    for mf in EDGE_META:  # test_vs_filtering_edge_cases
        for glob in (None, "*.pdf"):
            (hit,) = _retrieve(edge, "Foo", 1, mf, glob, extra)
            assert hit["dist"] < 1.0e-6 and hit["text"]


@pytest.mark.parametrize("mode", sorted(MAKERS))
def test_reference_filter_kats_host_logic(mode):
    _check_all({"_inner_override": MAKERS[mode]})


@pytest.mark.gpu
def test_reference_filter_kats_gpu():
    _check_all({})


def test_jmespath_filter_strings():
    """document_store.py:34-46"""
    assert _get_jmespath_filter("", "") is None and _get_jmespath_filter(None, None) is None
    assert _get_jmespath_filter("", "**/*.py") == "globmatch('**/*.py', path)"
    assert _get_jmespath_filter("path == `a.py`", None) == "(path == 'a.py')"
    assert _get_jmespath_filter('contains(path, `"x"`)', "*.pdf") == "(contains(path, 'x')) && globmatch('*.pdf', path)"
    assert _get_jmespath_filter("owner == `o'neil`", None) == r"(owner == 'o\'neil')"


def _splitting_store(extra_factory=None):
    docs = table_from_rows(["data", "_metadata"], [
        (b"alpha beta", {"path": "docs/a.txt", "owner": "ann"}),
        (b"gamma", {"path": "docs/b.md", "owner": "bob"}),
    ])
    vocab = {"alpha": [1.0, 0.0, 0.0], "beta": [0.0, 1.0, 0.0], "gamma": [0.0, 0.0, 1.0], "query": [0.9, 0.1, 0.0]}
    factory = extra_factory or BruteForceKnnFactory(dimensions=3, reserved_space=4, embedder=lambda x: vocab[x],
                                                    metric=BruteForceKnnMetricKind.L2SQ)
    splitter = lambda text: [(w, {"word": w}) for w in text.split()]  # noqa: E731
    return DocumentStore(docs, retriever_factory=factory, splitter=splitter), vocab


def _check_splitter(extra):
    store, _ = _splitting_store()
    assert sorted(r.values["text"] for r in store.chunked_docs.rows) == ["alpha", "beta", "gamma"]
    hits = _retrieve(store, "query", 3, None, None, extra)
    assert [h["text"] for h in hits] == ["alpha", "beta", "gamma"]  # increasing distance
    assert hits[0]["metadata"]["word"] == "alpha" and hits[0]["metadata"]["owner"] == "ann"
    assert [h["dist"] for h in hits] == sorted(h["dist"] for h in hits)
    hits = _retrieve(store, "query", 3, "owner == `bob`", None, extra)
    assert [h["text"] for h in hits] == ["gamma"]
    hits = _retrieve(store, "query", 1, "owner == `ann`", "docs/*.txt", extra)
    assert [h["text"] for h in hits] == ["alpha"]


def test_parser_splitter_and_merged_filters_host_logic():
    _check_splitter({"_inner_override": lambda: OracleInnerWithFilters(3, 4, 1024 * 128, BruteForceKnnMetricKind.L2SQ)})


@pytest.mark.gpu
def test_parser_splitter_and_merged_filters_gpu():
    _check_splitter({})


def test_hybrid_retriever_factory_host_logic():
    """test_hybrid_docstore_glob_filtering (:243-330) with two KNN retrievers: fused reply, one hit, RRF score -> dist."""
    f = lambda: BruteForceKnnFactory(dimensions=3, reserved_space=10, embedder=fake_embeddings_model, metric=COS)  # noqa: E731
    store = _store(factory=HybridIndexFactory([f(), f()]))
    for glob in GLOB_HITS:
        (hit,) = _retrieve(store, "test", 1, None, glob, {"_inner_override": MAKERS["loop"]})
        assert hit["text"] == "test" and hit["dist"] == -(2 / 61)
    assert _retrieve(store, "test", 1, None, "**/*.txt", {"_inner_override": MAKERS["loop"]}) == []
