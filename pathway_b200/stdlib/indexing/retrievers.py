"""/root/reference/python/pathway/stdlib/indexing/retrievers.py:7-36: a factory of inner indices also builds the
`DataIndex` over them."""

from __future__ import annotations

from ...debug import ColumnReference, Table
from .data_index import DataIndex


class InnerIndexFactory:
    def build_inner_index(self, data_column: ColumnReference, metadata_column: ColumnReference | None = None):
        raise NotImplementedError

    def build_index(self, data_column: ColumnReference, data_table: Table,
                    metadata_column: ColumnReference | None = None) -> DataIndex:
        return DataIndex(data_table, self.build_inner_index(data_column, metadata_column))
