"""Pure-Python restatement of the reference's reply post-processing: `HybridIndex._combine_results`
(reciprocal-rank fusion) and `DataIndex._repack_results` (collapsed rows).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Paths are relative to ``/root/reference/python/pathway/stdlib/indexing``.  The reference expresses both as
dataflow (flatten / groupby / join); each step below is that operator applied to plain Python rows, in the
reference's order, so it can be read side by side with the source.  Pinned by the reference's own known-answer
tests ``tests/ml/test_index.py:954-1011`` (`test_hybrid_index`: fused scores 1/3+1/4 and 1/3) and ``:1014-1058``
(`test_hybrid_index_ignores_duplicates`), replayed in ``tests/test_hybrid_rrf.py``.

Tie order: the reference sorts a group by ``(sort key, row Key, value)`` (src/engine/reduce.rs:890-923) where the
row Key of the deduplicated table is an opaque hash of (query id, matched id); exact score ties are therefore
unpinned -- this restatement breaks them by ascending matched id, like the product.
"""

from __future__ import annotations


def enumerate_results(results):
    """hybrid_index.py:43-47"""
    return [(i, x[0]) for i, x in enumerate(results, start=1)]


def query_single_retriever(replies: dict, k: float) -> list[dict]:
    """hybrid_index.py:49-70: reply -> rows {query id, matched id, 1/(k+rank)} (select + flatten + select)."""
    rows = []
    for query_id, reply in replies.items():
        for rank, matched in enumerate_results(reply):
            rows.append({"query": query_id, "matched": matched, "score": 1 / (k + rank)})
    return rows


def combine_results(replies_per_retriever: list[dict], number_of_matches, k: float = 60) -> dict:
    """hybrid_index.py:72-125.  replies_per_retriever[l] = {query id: [(matched id, score), ...]};
    number_of_matches: int or {query id: int}.  -> {query id: ((matched id, fused score), ...)}; queries without
    any candidate are absent (the groupby has no row for them)."""
    results = []
    for replies in replies_per_retriever:  # concat_reindex, :72-75
        results.extend(query_single_retriever(replies, k))
    removed_duplicates: dict = {}  # groupby(query, matched).reduce(sum(score)), :76-84
    for r in results:
        key = (r["query"], r["matched"])
        removed_duplicates[key] = removed_duplicates.get(key, 0.0) + r["score"]
    grouped_by_query: dict = {}  # groupby(query, sort_by=-sum).reduce(tuple((matched, score))), :90-101
    for (query, matched), score in removed_duplicates.items():
        grouped_by_query.setdefault(query, []).append((-score, matched, score))
    out = {}
    for query, items in grouped_by_query.items():
        items.sort(key=lambda t: (t[0], t[1]))
        n = number_of_matches if isinstance(number_of_matches, int) else number_of_matches[query]
        out[query] = tuple((matched, score) for _, matched, score in items)[:max(0, n)]  # limit_results, :86-88,118-120
    return out


def repack_collapsed(raw_result: dict, query_ids: list, data_rows: dict, columns: list[str]) -> dict:
    """data_index.py:328-347 + :92-200 (`_extract_data_collapsed_rows`).  raw_result = {query id: reply},
    data_rows = {id: {column: value}} (the data table as of the query).  -> {query id: {column: tuple, score: tuple}}."""
    id_matching = []  # flatten + unpack_col, :328-335
    for query, reply in raw_result.items():
        for matched, score in reply:
            id_matching.append({"query": query, "matched": matched, "score": score})
    joined = []  # inner join on matched == data.id, packed data, sort key -score, :139-151
    for m in id_matching:
        if m["matched"] in data_rows:
            packed = tuple(data_rows[m["matched"]][c] for c in columns)
            joined.append({"query": m["query"], "score": m["score"], "packed": packed, "sort": -m["score"]})
    for query in query_ids:  # artificial rows so every query answers, :153-166
        joined.append({"query": query, "score": 0, "packed": None, "sort": 0})
    groups: dict = {}
    order = 0
    for row in joined:  # groupby(id=query, sort_by=sort key).reduce(tuple(packed), tuple(score)), :172-183
        groups.setdefault(row["query"], []).append((row["sort"], order, row))
        order += 1
    out = {}
    for query, items in groups.items():
        items.sort(key=lambda t: (t[0], t[1]))
        topk = tuple(r["packed"] for _, _, r in items)
        score = tuple(r["score"] for _, _, r in items)
        keep = [x is not None for x in topk]  # remove_artificial_entries, :168-170,184-189
        topk = tuple(x for x, kp in zip(topk, keep) if kp)
        score = tuple(x for x, kp in zip(score, keep) if kp)
        transposed = tuple(zip(*topk)) if topk else tuple(() for _ in columns)  # transpose, :192-197
        out[query] = dict(zip(columns, transposed))
        out[query]["_pw_index_reply_score"] = score
    return out
