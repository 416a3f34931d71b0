"""The retrieval path of `DocumentStore`
(/root/reference/python/pathway/xpacks/llm/document_store.py:53-573), the caller that puts `BruteForceKnnFactory`
behind `VectorStoreServer` / the RAG templates (BASELINE.json configs[2]): documents -> parser -> splitter ->
`retriever_factory.build_index(text, chunks, metadata_column=metadata)`, and `retrieve_query`, which turns the REST
fields `metadata_filter` / `filepath_globpattern` into ONE JMESPath expression per query (`_get_jmespath_filter`,
:34-46) -- exactly the strings the filter push-down of the B200 index caches masks for.

Only that path: no MCP / REST server, statistics or input listings (control plane).  Parsers
and splitters are plain callables `text -> [(text, metadata_update)]`; the default parser decodes UTF-8 like the
reference's `Utf8Parser`, the default splitter keeps the document whole."""

from __future__ import annotations

from typing import Callable, Iterable

from pathway_b200.debug import Row, Table
from pathway_b200.stdlib.indexing.data_index import _SCORE


def _get_jmespath_filter(metadata_filter: str | None, filepath_globpattern: str | None) -> str | None:
    """document_store.py:34-46, character for character."""
    ret_parts = []
    if metadata_filter:
        metadata_filter = (
            metadata_filter.replace("'", r"\'").replace("`", "'").replace('"', "")
        )
        ret_parts.append(f"({metadata_filter})")
    if filepath_globpattern:
        ret_parts.append(f"globmatch('{filepath_globpattern}', path)")
    if ret_parts:
        return " && ".join(ret_parts)
    return None


def _utf8_parser(data) -> list[tuple[str, dict]]:
    return [(data.decode("utf-8") if isinstance(data, (bytes, bytearray)) else str(data), {})]


class DocumentStore:
    """document_store.py:53-101.  docs: a table (or several) with columns `data` and `_metadata` (a dict with at least
    `path` when glob patterns are used)."""

    def __init__(self, docs: Table | Iterable[Table], retriever_factory,
                 parser: Callable | None = None, splitter: Callable | None = None,
                 doc_post_processors: list[Callable] | None = None):
        self.docs = [docs] if isinstance(docs, Table) else list(docs)
        self.retriever_factory = retriever_factory
        self.parser = parser or _utf8_parser
        self.splitter = splitter or (lambda text: [(text, {})])
        self.doc_post_processors = doc_post_processors or []
        self.build_pipeline()

    @staticmethod
    def _apply_processor(table: Table, processor: Callable) -> Table:
        """document_store.py:285-303: a processor may split a document; metadata is merged into every part."""
        rows = []
        for r in table.rows:
            for n, (text, extra) in enumerate(processor(r.values["text"])):
                meta = dict(r.values["metadata"])
                meta.update(extra or {})
                rows.append(Row((r.key << 8 ^ n) & ((1 << 128) - 1) if n else r.key,
                                {"text": text, "metadata": meta}, r.time, r.diff))
        return Table(rows)

    def build_pipeline(self):
        """document_store.py:319-355 (the statistics tables of :357-407 are not built)."""
        if not self.docs:
            raise ValueError(
                """Please provide at least one data source, e.g. read files from disk:
pw.io.fs.read('./sample_docs', format='binary', mode='static', with_metadata=True)
"""
            )
        rows = []
        for t_no, t in enumerate(self.docs):  # concat_reindex
            for r in t.rows:
                key = r.key if len(self.docs) == 1 else (r.key ^ (t_no * 0x9E3779B97F4A7C15)) & ((1 << 128) - 1)
                meta = dict(r.values.get("_metadata") or {})
                meta["_file_id"] = str(key)
                rows.append(Row(key, {"text": r.values["data"], "metadata": meta}, r.time, r.diff))
        self.input_docs = Table(rows)
        self.parsed_docs = self._apply_processor(self.input_docs, self.parser)
        post = self.parsed_docs
        for p in self.doc_post_processors:  # :305-317: (text, metadata) -> (text, metadata)
            post = Table([Row(r.key, dict(zip(("text", "metadata"), p(r.values["text"], r.values["metadata"]))),
                              r.time, r.diff) for r in post.rows])
        self.post_processed_docs = post
        self.chunked_docs = self._apply_processor(post, self.splitter)
        self._retriever = self.retriever_factory.build_index(
            self.chunked_docs.text, self.chunked_docs, metadata_column=self.chunked_docs.metadata)

    @staticmethod
    def merge_filters(queries: Table) -> Table:
        """document_store.py:441-450"""
        rows = []
        for r in queries.rows:
            v = {k: x for k, x in r.values.items() if k not in ("metadata_filter", "filepath_globpattern")}
            v["metadata_filter"] = _get_jmespath_filter(r.values.get("metadata_filter"),
                                                        r.values.get("filepath_globpattern"))
            rows.append(Row(r.key, v, r.time, r.diff))
        return Table(rows)

    def retrieve_query(self, retrieval_queries: Table, **inner_kwargs) -> Table:
        """document_store.py:529-568.  queries: columns query, k, metadata_filter, filepath_globpattern ->
        column `result`: [{"text", "metadata", "dist"}] ordered by increasing dist (= -score)."""
        queries = self.merge_filters(retrieval_queries)
        found = self._retriever.query_as_of_now(queries.query, number_of_matches=queries.k,
                                                metadata_filter=queries.metadata_filter, **inner_kwargs)
        rows = []
        for r in found.rows:
            hits = [{"text": t, "metadata": m, "dist": -s}
                    for t, m, s in zip(r.values["text"] or (), r.values["metadata"] or (), r.values[_SCORE] or ())]
            rows.append(Row(r.key, {"result": sorted(hits, key=lambda h: h["dist"])}, r.time, r.diff))
        return Table(rows, list(found.error_log))

    @property
    def index(self):
        return self._retriever
