"""`KNNIndex`, the legacy facade of /root/reference/python/pathway/stdlib/ml/index.py:9-301 (BASELINE.json
configs[0] names it): same constructor and `get_nearest_items_asof_now` signature.

The reference's `KNNIndex` is an APPROXIMATE index (locality-sensitive hashing expressed as dataflow,
stdlib/ml/classifiers/_knn_lsh.py:135-290): it "might return less than k records per query or skip some closer
points" (index.py:14).  Here the same facade is answered EXACTLY by the B200 brute-force index -- "euclidean" is the
squared Euclidean distance (_knn_lsh.py:50-51 = BruteForceKnnMetricKind.L2SQ), "cosine" is 1 - cos (:54-57 = COS) --
so the LSH knobs (`n_or`, `n_and`, `bucket_length`) are accepted and ignored.  Wherever the reference's tests pin
results (they use point sets on which LSH finds every neighbour) the answers coincide; see tests/test_knn_index.py."""

from __future__ import annotations

from pathway_b200.debug import ColumnReference, Row, Table
from pathway_b200.engine import BruteForceKnnMetricKind
from pathway_b200.stdlib.indexing.data_index import _SCORE, DataIndex
from pathway_b200.stdlib.indexing.nearest_neighbors import BruteForceKnn


class KNNIndex:
    def __init__(self, data_embedding: ColumnReference, data: Table, n_dimensions: int, n_or: int = 20,
                 n_and: int = 10, bucket_length: float = 10.0, distance_type: str = "euclidean",
                 metadata: ColumnReference | None = None):
        if distance_type == "euclidean":
            metric = BruteForceKnnMetricKind.L2SQ
        elif distance_type == "cosine":
            metric = BruteForceKnnMetricKind.COS
        else:  # _knn_lsh.py:94-97
            raise ValueError(
                f"Not supported `type` {distance_type} in knn_lsh_classifier_train. "
                "The allowed values are 'euclidean' and 'cosine'."
            )
        self.data = data
        self._columns = data.column_names()
        inner = BruteForceKnn(data_column=data_embedding, metadata_column=metadata, dimensions=n_dimensions,
                              reserved_space=400, metric=metric)
        self._index = DataIndex(data, inner)

    def get_nearest_items(self, query_embedding, k=3, collapse_rows=True, with_distances=False, metadata_filter=None):
        """index.py:54-191 keeps old answers up to date when documents arrive later; the external-index operator is
        as-of-now only (nearest_neighbors.py:206-215), so is this facade."""
        raise NotImplementedError("KNNIndex over the B200 index is supported only in the as-of-now variant: "
                                  "use get_nearest_items_asof_now")

    def get_nearest_items_asof_now(self, query_embedding: ColumnReference, k: ColumnReference | int = 3,
                                   collapse_rows: bool = True, with_distances: bool = False,
                                   metadata_filter: ColumnReference | None = None, **inner_kwargs) -> Table:
        """index.py:193-261.  Result rows carry the data table's columns (tuples of the matched rows' values when
        `collapse_rows`, ordered by increasing distance) and, with `with_distances`, `dist`; one row per query, keyed
        like the query (queries without a match keep empty tuples, index.py:279-285)."""
        res = self._index.query_as_of_now(query_embedding, number_of_matches=k, collapse_rows=collapse_rows,
                                          metadata_filter=metadata_filter, **inner_kwargs)
        rows = []
        for r in res.rows:
            v = {c: r.values[c] for c in self._columns}
            if with_distances:  # score = -distance (brute_force_knn_integration.rs:230)
                s = r.values[_SCORE]
                v["dist"] = tuple(-x for x in s) if collapse_rows else (None if s is None else -s)
            if not collapse_rows:
                if r.values[_SCORE] is None:
                    continue  # _extract_data_flat: queries without a match only keep their id (index.py:287-301)
                v["query_id"] = r.key
            rows.append(Row(r.key, v, r.time, r.diff))
        return Table(rows, list(res.error_log))
