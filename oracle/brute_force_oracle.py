"""numpy restatement (f64) of the reference's brute-force KNN index.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Every function cites the reference lines it restates; paths are relative to
``/root/reference``.  The reference is Rust and cannot be compiled in this
image (no cargo/rustc), so this restatement is pinned instead against the
reference's own known-answer tests (``tests/test_oracle_golden.py`` replays
``python/pathway/tests/external_index/test_brute_force_knn.py`` and the
brute-force cases of ``python/pathway/tests/ml/test_index.py``).

Arithmetic is plain f64: ``ndarray::dot`` is sum(a*b) in f64 with an unpinned
accumulation order, so distances are compared with a tolerance and ids exactly.
"""

from __future__ import annotations

import math

import numpy as np

L2SQ = 0
COS = 1
IP = 2  # NOT part of the brute-force reference: usearch's MetricKind::IP (distance = 1 - x.q), used only to check the
        # USearchKnn API mirror against the reference's usearch KATs (usearch 2.15.3 is third-party: parity unpinned)


class MissingKey(KeyError):
    """`DataError::MissingKey(key)` -- src/external_integration/mod.rs:248-252."""


class KeyToU64IdMapper:
    """src/external_integration/mod.rs:209-256 (dense slot ids, two hash maps)."""

    def __init__(self) -> None:
        self.next_id = 0
        self.id_to_key: dict[int, int] = {}
        self.key_to_id: dict[int, int] = {}

    def get_next_free_u64_id(self, key: int) -> int:  # mod.rs:224-233
        got = self.key_to_id.get(key)
        if got is not None:
            return got
        idx = self.next_id
        self.next_id += 1
        self.id_to_key[idx] = key
        self.key_to_id[key] = idx
        return idx

    def assign_key(self, key: int, idx: int) -> None:  # mod.rs:235-238
        self.id_to_key[idx] = key
        self.key_to_id[key] = idx

    def decrement_next_free_id(self) -> None:  # mod.rs:240-242
        self.next_id -= 1

    def get_key_for_id(self, idx: int):  # mod.rs:244-246
        return self.id_to_key.get(idx)

    def remove_key(self, key: int) -> int:  # mod.rs:248-255
        if key not in self.key_to_id:
            raise MissingKey(key)
        idx = self.key_to_id.pop(key)
        self.id_to_key.pop(idx, None)
        return idx


def _ordered_float_key(d: float, slot: int):
    """`(OrderedFloat(distance), idx)` ordering, NaN greater than everything
    (ordered-float semantics) -- brute_force_knn_integration.rs:220-224."""
    return (1, 0.0, slot) if math.isnan(d) else (0, d, slot)


class BruteForceKNNIndex:
    """src/external_integration/brute_force_knn_integration.rs:22-239."""

    def __init__(self, dimensions: int, reserved_space: int, auxiliary_space: int, metric: int):
        # :34-51
        self.index_array = np.zeros((reserved_space, dimensions), dtype=np.float64)
        self.current_size = 0
        self.current_allocated = reserved_space
        self.minimum_allocated = reserved_space
        self.auxiliary_space = auxiliary_space
        self.dimensions = dimensions
        self.metric = metric
        self.key_to_id_mapper = KeyToU64IdMapper()

    # ---- :114-144
    def add(self, add_data):
        """add_data: list of (key, sequence of float). Returns [(key, None|Exception)]."""
        n = len(add_data)
        if n + self.current_size > self.current_allocated:
            new_allocated = max(2 * self.current_allocated, n + self.current_size)  # :116-119
            new_arr = np.zeros((new_allocated, self.dimensions), dtype=np.float64)
            flat_old = self.index_array.reshape(-1)
            new_arr.reshape(-1)[: flat_old.size] = flat_old  # :121-124 (zip of flat iters)
            self.index_array = new_arr
            self.current_allocated = new_allocated
        self.current_size += n  # :130
        ret = []
        for key, data in add_data:
            idx = self.key_to_id_mapper.get_next_free_u64_id(key)  # :134
            m = min(len(data), self.dimensions)  # zip semantics, :137-139
            self.index_array[idx, :m] = np.asarray(data, dtype=np.float64)[:m]
            ret.append((key, None))
        return ret

    # ---- :146-187
    def remove(self, keys):
        ret = []
        for key in keys:
            last_row_key = self.key_to_id_mapper.get_key_for_id(self.current_size - 1)  # :150-153
            if last_row_key is None:
                # the reference unwrap()s here and panics; out of contract.
                raise RuntimeError("remove from empty/inconsistent index (reference panics)")
            try:
                removed_key_id = self.key_to_id_mapper.remove_key(key)  # :154
            except MissingKey as e:
                ret.append((key, e))  # :169
                continue
            self.current_size -= 1  # :156
            self.key_to_id_mapper.decrement_next_free_id()  # :157
            if last_row_key != key:  # :158-166
                self.index_array[removed_key_id, :] = self.index_array[self.current_size, :]
                self.key_to_id_mapper.assign_key(last_row_key, removed_key_id)
            ret.append((key, None))
        # shrink rule :174-185
        if (4 * self.current_size < self.current_allocated
                and self.current_allocated // 2 >= self.minimum_allocated):
            new_allocated = self.current_allocated // 2
            new_arr = np.zeros((new_allocated, self.dimensions), dtype=np.float64)
            flat_old = self.index_array.reshape(-1)
            new_arr.reshape(-1)[:] = flat_old[: new_arr.size]
            self.index_array = new_arr
            self.current_allocated = new_allocated
        return ret

    # ---- :70-111
    def _fill_distances(self, index_arr, query_arr, dot_p):
        index_sq_norms = np.einsum("ij,ij->i", index_arr, index_arr)  # :75-78 / :97-100
        query_sq_norms = np.einsum("ij,ij->j", query_arr, query_arr)  # :80-83 / :101-104
        with np.errstate(divide="ignore", invalid="ignore"):
            if self.metric == COS:  # :85-89
                return 1.0 - dot_p / np.sqrt(index_sq_norms[:, None] * query_sq_norms[None, :])
            if self.metric == IP:
                return 1.0 - dot_p
            # :106-110
            return index_sq_norms[:, None] + query_sq_norms[None, :] - 2.0 * dot_p

    # ---- :189-238
    def search(self, queries):
        """queries: list of (key, vector, limit) -> [(key, [(match_key, score)])]."""
        if self.current_size == 0:  # :193-198
            return [(key, []) for key, _, _ in queries]
        index_arr = self.index_array[: self.current_size, : self.dimensions]
        max_queries = max(1, self.auxiliary_space // self.current_size)  # :205
        ret = []
        query_arr = np.zeros((self.dimensions, max_queries), dtype=np.float64)  # :207
        for b in range(0, len(queries), max_queries):
            batch = queries[b:b + max_queries]
            for j, (_key, data, _k) in enumerate(batch):  # :209-213 (zip: stale tail kept)
                m = min(len(data), self.dimensions)
                query_arr[:m, j] = np.asarray(data, dtype=np.float64)[:m]
            sl = query_arr[:, : len(batch)]
            dot_p = index_arr @ sl  # :215
            dist = self._fill_distances(index_arr, sl, dot_p)  # :216
            for j, (key, _data, limit) in enumerate(batch):  # :218-232
                col = dist[:, j]
                order = k_smallest(col, limit)
                res = [(self.key_to_id_mapper.get_key_for_id(int(i)), -float(col[i])) for i in order]
                ret.append((key, res))
        return ret


def k_smallest(col: np.ndarray, limit: int) -> list[int]:
    """`itertools::k_smallest(limit)` over `(OrderedFloat(d), idx)`: the `limit`
    smallest in ascending order -- brute_force_knn_integration.rs:220-224."""
    n = col.shape[0]
    if limit <= 0:
        return []
    if limit >= n:
        cand = np.arange(n)
    else:
        # argpartition is only a pre-filter; ties at the boundary are resolved
        # by including every element <= the boundary value.
        key = np.where(np.isnan(col), np.inf, col)
        kth = np.partition(key, limit - 1)[limit - 1]
        cand = np.nonzero(key <= kth)[0]
    cand = sorted((int(i) for i in cand), key=lambda i: _ordered_float_key(float(col[i]), i))
    return cand[:limit]


def bf16_round(x: np.ndarray) -> np.ndarray:
    """Round-to-nearest-even f32->bf16->f64, the quantisation the B200 corpus
    applies at upsert time (not part of the reference; used to prepare
    'identical inputs' for parity runs)."""
    a = np.ascontiguousarray(x, dtype=np.float32)
    u = a.view(np.uint32).astype(np.uint64)
    nan = np.isnan(a)
    rounded = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    out = rounded.astype(np.uint32).view(np.float32).astype(np.float64)
    out[nan] = np.nan
    return out
