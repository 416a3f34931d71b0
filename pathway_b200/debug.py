"""Minimal stand-in for the `pw.debug` / `pw.Table` scaffolding the reference's index tests use
(python/pathway/debug/__init__.py:222-446 `table_from_markdown`, `table_from_rows`;
python/pathway/internals/table.py:591-634 `Table._external_index_as_of_now`).

Just enough of a row container to replay the reference's known-answer tests through the same
operator semantics without the Rust engine: rows carry (key, values, time, diff)."""

from __future__ import annotations

import hashlib
from dataclasses import dataclass, field
from typing import Any, Callable

from .engine import ExternalIndexData, ExternalIndexFactory, ExternalIndexQuery, Scope

_INDEX_REPLY = "_pw_index_reply"  # python/pathway/stdlib/indexing/colnames.py


def ref_scalar(*values) -> int:
    """Deterministic 128-bit Pointer from primary-key values (stand-in for Key::for_values)."""
    h = hashlib.blake2b(repr(values).encode(), digest_size=16).digest()
    return int.from_bytes(h, "little")


@dataclass
class Row:
    key: int
    values: dict
    time: int = 0
    diff: int = 1


@dataclass(frozen=True)
class ColumnReference:
    table: "Table"
    name: str


@dataclass
class Table:
    rows: list = field(default_factory=list)
    error_log: list = field(default_factory=list)  # what pw.global_error_log() would hold for this op

    def __getattr__(self, name):
        if name.startswith("__") or name in ("rows", "error_log"):
            raise AttributeError(name)
        return ColumnReference(self, name)

    def __getitem__(self, name):
        return ColumnReference(self, name)

    def column_names(self) -> list[str]:
        names: dict = {}
        for r in self.rows:
            for c in r.values:
                names.setdefault(c)
        return list(names)

    def with_columns(self, **fns: Callable[[dict], Any] | Any) -> "Table":
        out = []
        for r in self.rows:
            v = dict(r.values)
            for name, fn in fns.items():
                v[name] = fn(r.values) if callable(fn) else fn
            out.append(Row(r.key, v, r.time, r.diff))
        return Table(out)

    def _external_index_as_of_now(self, queries: "Table", *, index_column, query_column,
                                  index_factory: ExternalIndexFactory, res_type=None,
                                  query_responses_limit_column=None, index_filter_data_column=None,
                                  query_filter_column=None, _inner_override=None) -> "Table":
        """table.py:591-634 + expression_evaluator.py:902-947 + python_api.rs:3367-3381."""
        name = lambda c: c.name if isinstance(c, ColumnReference) else c  # noqa: E731
        out, errors = Scope().use_external_index_as_of_now(
            ExternalIndexData(self, name(index_column), name(index_filter_data_column)),
            ExternalIndexQuery(queries, name(query_column), name(query_responses_limit_column), name(query_filter_column)),
            None, index_factory, _inner_override=_inner_override)
        rows = []
        for key, (qrow, reply), time, diff in out:
            v = dict(qrow)
            v[_INDEX_REPLY] = reply
            rows.append(Row(key, v, time, diff))
        return Table(rows, errors)


def _parse(tok: str):
    tok = tok.strip()
    for cast in (int, float):
        try:
            return cast(tok)
        except ValueError:
            pass
    return tok


def table_from_markdown(txt: str, id_from: list[str] | None = None) -> Table:
    lines = [ln for ln in txt.strip().splitlines() if ln.strip()]
    header = [h.strip() for h in lines[0].split("|")]
    rows = []
    for n, ln in enumerate(lines[1:]):
        cells = [_parse(c) for c in ln.split("|")]
        d = dict(zip(header, cells))
        time = d.pop("__time__", 0)
        diff = d.pop("__diff__", 1)
        key = ref_scalar(*[d[c] for c in id_from]) if id_from else ref_scalar("row", n)
        rows.append(Row(key, d, time, diff))
    return Table(rows)


def table_from_rows(columns: list[str], rows: list[tuple], id_from: list[str] | None = None,
                    is_stream: bool = False) -> Table:
    out = []
    for n, r in enumerate(rows):
        if is_stream:
            *vals, time, diff = r
        else:
            vals, time, diff = list(r), 0, 1
        d = dict(zip(columns, vals))
        key = ref_scalar(*[d[c] for c in id_from]) if id_from else ref_scalar("row", n)
        out.append(Row(key, d, time, diff))
    return Table(out)
