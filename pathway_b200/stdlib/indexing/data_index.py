"""Drop-in mirror of `DataIndex` (/root/reference/python/pathway/stdlib/indexing/data_index.py:273-470):
augments an inner index's `(id, score)` replies with the columns of `data_table`.

The reference repacks with flatten -> unpack -> as-of-now join -> concat(artificial rows) -> groupby(sort_by) ->
two tuple reducers -> transpose (data_index.py:136-200, 294-347).  Behind the operator the same answer is one
pass per query over its reply: look the matched ids up in the data table as it stands at the
query's time, order by descending score, emit the columns as tuples."""

from __future__ import annotations

from dataclasses import dataclass
from typing import Any

from ...debug import _INDEX_REPLY, ColumnReference, Row, Table, ref_scalar
from ...engine import ERROR

_QUERY_ID = "_pw_query_id"  # colnames.py
_MATCHED_ID = "_pw_index_reply_id"
_SCORE = "_pw_index_reply_score"


def _state_by_time(table: Table):
    """[(time, {key: values})]: the table as it stands after each of its timestamps (retractions of a timestamp
    before its insertions, as in IndexDerivedImpl, mod.rs:144-150)."""
    state: dict[int, dict] = {}
    out = []
    rows = sorted(table.rows, key=lambda r: (r.time, r.diff > 0))
    i = 0
    while i < len(rows):
        t = rows[i].time
        while i < len(rows) and rows[i].time == t:
            r = rows[i]
            if r.diff < 0:
                state.pop(r.key, None)
            elif r.diff > 0:
                state[r.key] = r.values
            i += 1
        out.append((t, dict(state)))
    return out


def _state_at(states, time: int) -> dict:
    cur: dict = {}
    for t, st in states:
        if t > time:
            break
        cur = st
    return cur


@dataclass
class DataIndex:
    """data_index.py:273-291: `data_table` holds the supplementary data (matched by id), `inner_index` answers with
    the matched ids."""

    data_table: Table
    inner_index: Any

    def _repack_results(self, raw_result: Table, query_table: Table, collapse_rows: bool) -> Table:
        """data_index.py:294-347 with the as-of-now extractors (:63-89 flat, :92-200 collapsed)."""
        columns = self.data_table.column_names()
        states = _state_by_time(self.data_table)
        reply_of = {(r.key, r.time, r.diff): r.values[_INDEX_REPLY] for r in raw_result.rows}
        rows = []
        for q in query_table.rows:
            reply = reply_of.get((q.key, q.time, q.diff), ())
            data = _state_at(states, q.time)
            matched = [(m, s, data[m]) for m, s in (() if reply is ERROR else reply) if m in data]  # inner join :78-83
            matched.sort(key=lambda t: -t[1])  # _pw_groupby_sort_key = -score (:156); stable => reply order on ties
            if collapse_rows:
                v = dict(q.values)
                for c in columns:
                    v[c] = tuple(d.get(c) for _, _, d in matched)
                v[_SCORE] = tuple(s for _, s, _ in matched)
                rows.append(Row(q.key, v, q.time, q.diff))
            elif matched:
                for m, s, d in matched:
                    v = dict(q.values)
                    v.update({c: d.get(c) for c in columns})
                    v[_SCORE] = s
                    rows.append(Row(ref_scalar(q.key, m), v, q.time, q.diff))
            else:  # left join keeps the unanswered query with None on the right (:84-89)
                v = dict(q.values)
                v.update({c: None for c in columns})
                v[_SCORE] = None
                rows.append(Row(ref_scalar(q.key, None), v, q.time, q.diff))
        return Table(rows, list(raw_result.error_log))

    def query(self, query_column: ColumnReference, *, number_of_matches=3, collapse_rows: bool = True,
              metadata_filter=None) -> Table:
        raw = self.inner_index.query(query_column, number_of_matches=number_of_matches,
                                     metadata_filter=metadata_filter)
        return self._repack_results(raw, query_column.table, collapse_rows)

    def query_as_of_now(self, query_column: ColumnReference, number_of_matches=3, collapse_rows: bool = True,
                        metadata_filter=None, **inner_kwargs) -> Table:
        raw = self.inner_index.query_as_of_now(query_column, number_of_matches=number_of_matches,
                                               metadata_filter=metadata_filter, **inner_kwargs)
        return self._repack_results(raw, query_column.table, collapse_rows)
