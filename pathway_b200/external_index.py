"""Host-side mirror of the reference's external-index framework for the KNN path.

The reference's host code for this path is Rust; there is no Rust toolchain in
this image, so the layers ABOVE the C ABI are mirrored here with the same names,
argument meaning and error behaviour (paths relative to /root/reference):

  B200BruteForceKNNIndex      <-> BruteForceKNNIndex as NonFilteringExternalIndex<Vec<f64>,Vec<f64>>
                                  src/external_integration/brute_force_knn_integration.rs:113-239
                                  (all arithmetic happens in libpwknn.so on the GPU)
  DerivedFilteredSearchIndex  <-> src/external_integration/mod.rs:384-604 (JMESPath + globmatch, k-doubling)
  IndexDerivedImpl            <-> src/external_integration/mod.rs:51-205   (accessors, Error rows, remove-then-add)
  use_external_index_as_of_now<-> src/engine/dataflow/operators/external_index.rs:76-158 + operators/utils.rs:111-137

A maintainer integrating into the real engine keeps the Rust versions of the
last three untouched and only swaps the innermost index.
"""

from __future__ import annotations

import ctypes as C
import functools
import re
import json
from collections import defaultdict
from typing import Any, Callable, Iterable, Optional

import jmespath
import jmespath.exceptions
import numpy as np
from jmespath import functions as _jp_functions

from . import _native
from .engine import ERROR, BruteForceKnnMetricKind


# ----------------------------------------------------------------------------- errors (src/engine/error.rs)
class DataError(Exception):
    pass


class MissingKey(DataError):  # DataError::MissingKey(key)  error.rs:275-276
    def __init__(self, key):
        super().__init__(f"key missing in output table: {key}")
        self.key = key


class TypeMismatch(DataError):  # DataError::TypeMismatch{expected, value}
    def __init__(self, expected, value):
        super().__init__(f"type mismatch: expected {expected}, got {value!r}")


class ValueError_(DataError):  # DataError::ValueError(String)
    pass


class ErrorInIndexUpdate(DataError):  # error.rs:329-333
    def __init__(self):
        super().__init__("error in index update")


class ErrorInIndexSearch(DataError):
    def __init__(self):
        super().__init__("error in index search")


def keys_to_array(keys) -> np.ndarray:
    a = np.empty((len(keys), 2), dtype=np.uint64)
    for i, k in enumerate(keys):
        k = int(k)
        a[i, 0] = k & 0xFFFFFFFFFFFFFFFF
        a[i, 1] = (k >> 64) & 0xFFFFFFFFFFFFFFFF
    return a


def array_to_keys(a: np.ndarray) -> list[int]:
    return [(int(hi) << 64) | int(lo) for lo, hi in a]


# ----------------------------------------------------------------------------- innermost index (GPU)
class B200BruteForceKNNIndex:
    """`impl NonFilteringExternalIndex<Vec<f64>, Vec<f64>> for BruteForceKNNIndex`, on a B200.

    add / remove / search keep the trait's shapes: lists of tuples in, lists of
    ``(key, None | Exception)`` / ``(key, [(match_key, score), ...] | Exception)`` out.
    """

    def __init__(self, dimensions: int, reserved_space: int, auxiliary_space: int = 1024 * 128,
                 metric: BruteForceKnnMetricKind | int = BruteForceKnnMetricKind.COS, device: int | None = None,
                 exact_rerank: str | None = None, devices: list[int] | None = None,
                 weights: list[float] | None = None):
        """exact_rerank: None (re-rank on the stored bf16 values), "f32" or "f64" (keep an exact copy of the corpus in
        HBM for the f64 re-rank: results are the reference's on the raw inputs); default from PWKNN_EXACT_RERANK.
        devices: several CUDA ordinals -> one handle whose rows are spread over them (pwknn_create_multi); default
        from PWKNN_DEVICES="0,1,...".  weights: placement share per device (pwknn_multi_set_weights; see
        pathway_b200.placement.calibrate_device_weights)."""
        self._L = _native.lib()
        self.dimensions = int(dimensions)
        metric_i = metric.value if isinstance(metric, BruteForceKnnMetricKind) else int(metric)
        if devices is None and device is None:
            import os
            if os.environ.get("PWKNN_DEVICES"):
                devices = [int(x) for x in os.environ["PWKNN_DEVICES"].split(",")]
        if devices is not None:
            import os
            if exact_rerank is None:
                exact_rerank = os.environ.get("PWKNN_EXACT_RERANK", "none")
            flags = {"none": 0, "bf16": 0, "f32": 1, "f64": 2}[exact_rerank]
            arr = (C.c_int * len(devices))(*[int(d) for d in devices])
            h = C.c_void_p()
            _native.check(None, self._L.pwknn_create_multi(self.dimensions, int(reserved_space), int(auxiliary_space),
                                                           metric_i, arr, len(devices), flags, C.byref(h)))
            self._h, self.metric, self.device, self.devices = h, metric_i, int(devices[0]), list(devices)
            if weights is not None:
                self.set_weights(weights)
            return
        self.devices = None
        if device is None:
            import os
            device = int(os.environ.get("LOCAL_RANK", "0")) if os.environ.get("PWKNN_DEVICE") is None \
                else int(os.environ["PWKNN_DEVICE"])
        h = C.c_void_p()
        if exact_rerank is None:
            rc = self._L.pwknn_create(self.dimensions, int(reserved_space), int(auxiliary_space), metric_i, int(device),
                                      C.byref(h))
        else:
            flags = {"none": 0, "bf16": 0, "f32": 1, "f64": 2}[exact_rerank]
            rc = self._L.pwknn_create_ex(self.dimensions, int(reserved_space), int(auxiliary_space), metric_i,
                                         int(device), flags, C.byref(h))
        _native.check(None, rc)
        self._h = h
        self.metric = metric_i
        self.device = int(device)

    def snapshot_save(self, path: str):
        _native.check(self._h, self._L.pwknn_snapshot_save(self._h, str(path).encode()))

    @classmethod
    def snapshot_load(cls, path: str, device: int = 0, devices: list[int] | None = None) -> "B200BruteForceKNNIndex":
        """devices: restore a multi-device snapshot (same number of shards as it was saved with)."""
        self = cls.__new__(cls)
        self._L = _native.lib()
        h = C.c_void_p()
        if devices is not None:
            arr = (C.c_int * len(devices))(*[int(d) for d in devices])
            _native.check(None, self._L.pwknn_snapshot_load_multi(str(path).encode(), arr, len(devices), C.byref(h)))
            device = devices[0]
        else:
            _native.check(None, self._L.pwknn_snapshot_load(str(path).encode(), int(device), C.byref(h)))
        self._h = h
        self.dimensions = int(self._L.pwknn_dimensions(h))
        self.device = int(device)
        self.devices = list(devices) if devices is not None else None
        self.metric = None
        return self

    def close(self):
        if getattr(self, "_h", None):
            if not getattr(self, "_borrowed", False):
                self._L.pwknn_destroy(self._h)
            self._h = None

    def set_weights(self, weights: list[float]):
        """Placement share per device of a multi-device handle; only while the index is empty."""
        arr = (C.c_double * len(weights))(*[float(w) for w in weights])
        _native.check(self._h, self._L.pwknn_multi_set_weights(self._h, len(weights), arr))

    def shard(self, i: int) -> "B200BruteForceKNNIndex":
        """The i-th per-device index of a multi-device handle (borrowed: closing it does nothing); the per-device
        entry points (add_device, read_rows, search_device, ...) work on it."""
        h = self._L.pwknn_shard(self._h, int(i))
        if not h:
            raise IndexError(f"no shard {i} (single-device handle or out of range)")
        sh = type(self).__new__(type(self))
        sh._L, sh._h, sh._borrowed = self._L, C.c_void_p(h), True
        sh.dimensions, sh.metric, sh.devices = self.dimensions, self.metric, None
        sh.device = int(self._L.pwknn_device(sh._h))
        return sh

    def shards_of_keys(self, keys_u64x2: np.ndarray) -> np.ndarray:
        """Which shard the router puts each key on (all 0 for a single-device handle)."""
        keys_u64x2 = np.ascontiguousarray(keys_u64x2, dtype=np.uint64)
        out = np.zeros(len(keys_u64x2), dtype=np.uint32)
        _native.check(self._h, self._L.pwknn_shards_of_keys(self._h, len(keys_u64x2), keys_u64x2.ctypes.data,
                                                             out.ctypes.data))
        return out

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- introspection
    @property
    def current_size(self) -> int:
        return int(self._L.pwknn_size(self._h))

    @property
    def current_allocated(self) -> int:
        return int(self._L.pwknn_allocated(self._h))

    def set_option(self, name: str, value: int):
        _native.check(self._h, self._L.pwknn_set_option(self._h, name.encode(), int(value)))

    def stat(self, name: str) -> int:
        return int(self._L.pwknn_get_stat(self._h, name.encode()))

    @property
    def stream_ptr(self) -> int:
        return int(self._L.pwknn_stream(self._h) or 0)

    def read_rows(self, slot0: int, n: int) -> np.ndarray:
        out = np.empty((n, self.dimensions), dtype=np.float64)
        _native.check(self._h, self._L.pwknn_read_rows(self._h, slot0, n, out.ctypes.data))
        return out

    # -- array tier (what a Rust shim would call)
    def add_arrays(self, keys_u64x2: np.ndarray, rows: np.ndarray) -> np.ndarray:
        keys_u64x2 = np.ascontiguousarray(keys_u64x2, dtype=np.uint64)
        n = len(keys_u64x2)
        status = np.zeros(n, dtype=np.int32)
        if rows.dtype == np.float32:
            rows = np.ascontiguousarray(rows)
            fn = self._L.pwknn_add_f32
        else:
            rows = np.ascontiguousarray(rows, dtype=np.float64)
            fn = self._L.pwknn_add
        assert rows.shape == (n, self.dimensions), (rows.shape, n, self.dimensions)
        _native.check(self._h, fn(self._h, n, keys_u64x2.ctypes.data, rows.ctypes.data, status.ctypes.data))
        return status

    def remove_arrays(self, keys_u64x2: np.ndarray) -> np.ndarray:
        keys_u64x2 = np.ascontiguousarray(keys_u64x2, dtype=np.uint64)
        status = np.zeros(len(keys_u64x2), dtype=np.int32)
        _native.check(self._h, self._L.pwknn_remove(self._h, len(keys_u64x2), keys_u64x2.ctypes.data,
                                                    status.ctypes.data))
        return status

    def search_arrays(self, queries: np.ndarray, limits: np.ndarray):
        """-> (out_keys [T,2] u64, out_scores [T] f64, counts [nq] u32, status [nq] i32) with
        clamped = min(limits, current_size), T = sum(clamped); query q's matches start at offset sum(clamped[:q])."""
        if queries.dtype == np.float32:
            queries = np.ascontiguousarray(queries)
            fn = self._L.pwknn_search_f32
        else:
            queries = np.ascontiguousarray(queries, dtype=np.float64)
            fn = self._L.pwknn_search
        # a limit beyond the index size cannot return more than `size` rows: clamp before sizing the outputs
        # (a "give me everything" limit of 1e9 must not allocate 16 GB); the reply layout follows the clamped limits
        limits = np.minimum(np.asarray(limits, dtype=np.uint64), np.uint64(self.current_size)).astype(np.uint32)
        limits = np.ascontiguousarray(limits)
        nq = len(limits)
        assert queries.shape == (nq, self.dimensions)
        tot = int(limits.astype(np.uint64).sum())
        out_keys = np.zeros((max(tot, 1), 2), dtype=np.uint64)
        out_scores = np.zeros(max(tot, 1), dtype=np.float64)
        counts = np.zeros(nq, dtype=np.uint32)
        status = np.zeros(nq, dtype=np.int32)
        _native.check(self._h, fn(self._h, nq, queries.ctypes.data, limits.ctypes.data, out_keys.ctypes.data,
                                  out_scores.ctypes.data, counts.ctypes.data, status.ctypes.data))
        return out_keys, out_scores, counts, status

    # -- filter push-down (pwknn_filter_*): per-row booleans kept by the library next to the rows
    NO_FILTER = 0xFFFFFFFF

    def filter_create(self) -> int:
        fid = C.c_uint32()
        _native.check(self._h, self._L.pwknn_filter_create(self._h, C.byref(fid)))
        return int(fid.value)

    def filter_destroy(self, fid: int):
        _native.check(self._h, self._L.pwknn_filter_destroy(self._h, int(fid)))

    def filter_set(self, fid: int, keys: list[int], passes: list[bool]) -> np.ndarray:
        if not keys:
            return np.zeros(0, dtype=np.int32)
        ka = keys_to_array(keys)
        pa = np.ascontiguousarray(np.asarray(passes, dtype=np.uint8))
        st = np.zeros(len(keys), dtype=np.int32)
        _native.check(self._h, self._L.pwknn_filter_set(self._h, int(fid), len(keys), ka.ctypes.data, pa.ctypes.data,
                                                        st.ctypes.data))
        return st

    def search_filtered(self, queries: list[tuple[int, Any, int, Optional[int]]]):
        """queries: (key, vector, limit, filter id | None) -> [(key, [(match_key, score)] | Exception)], one pass."""
        if not queries:
            return []
        ret: list = [None] * len(queries)
        good = [i for i, (_k, v, lim, _f) in enumerate(queries) if self._vec_ok(v) and 0 <= lim <= 0xFFFFFFFF]
        good_set = set(good)
        for i, (k, v, lim, _f) in enumerate(queries):
            if i not in good_set:
                ret[i] = (k, ValueError_("bad query vector or limit"))
        if good:
            q = np.array([np.asarray(queries[i][1], dtype=np.float64) for i in good], dtype=np.float64)
            q = np.ascontiguousarray(q.reshape(len(good), self.dimensions))
            limits = np.minimum(np.array([queries[i][2] for i in good], dtype=np.uint64), self.current_size).astype(np.uint32)
            fids = np.array([self.NO_FILTER if queries[i][3] is None else queries[i][3] for i in good], dtype=np.uint32)
            tot = int(limits.astype(np.uint64).sum())
            ok = np.zeros((max(tot, 1), 2), dtype=np.uint64); sc = np.zeros(max(tot, 1), dtype=np.float64)
            cnt = np.zeros(len(good), dtype=np.uint32); st = np.zeros(len(good), dtype=np.int32)
            _native.check(self._h, self._L.pwknn_search_filtered(self._h, len(good), q.ctypes.data, limits.ctypes.data,
                                                                 fids.ctypes.data, ok.ctypes.data, sc.ctypes.data,
                                                                 cnt.ctypes.data, st.ctypes.data))
            off = 0
            for j, i in enumerate(good):
                k = queries[i][0]
                if st[j] != _native.ROW_OK:
                    ret[i] = (k, ValueError_("non-finite query component or unknown filter"))
                else:
                    ret[i] = (k, [((int(ok[off + r, 1]) << 64) | int(ok[off + r, 0]), float(sc[off + r]))
                                  for r in range(int(cnt[j]))])
                off += int(limits[j])
        return ret

    # -- device tier (torch tensors; bench `value` and the sharded host use this)
    def add_device(self, keys_u64x2: np.ndarray, rows_bf16, stream_ptr: int = 0) -> np.ndarray:
        """rows_bf16: torch.bfloat16 CUDA tensor [n, >=dims] on this index's device."""
        keys_u64x2 = np.ascontiguousarray(keys_u64x2, dtype=np.uint64)
        n = len(keys_u64x2)
        assert rows_bf16.is_cuda and rows_bf16.dim() == 2 and rows_bf16.shape[0] == n and rows_bf16.stride(1) == 1
        status = np.zeros(n, dtype=np.int32)
        _native.check(self._h, self._L.pwknn_add_device_bf16(self._h, n, keys_u64x2.ctypes.data, rows_bf16.data_ptr(),
                                                             rows_bf16.stride(0), stream_ptr or None,
                                                             status.ctypes.data))
        return status

    def search_device(self, queries_bf16, k: int, out_slots, out_dist, stream_ptr: int = 0):
        """queries_bf16 [nq, >=dims] bf16 CUDA; out_slots int32 [nq,k] (0xFFFFFFFF == -1 = none), out_dist f64 [nq,k].
        Asynchronous on `stream_ptr` (0 = the index's own stream)."""
        nq = queries_bf16.shape[0]
        assert queries_bf16.is_cuda and queries_bf16.stride(1) == 1
        assert out_slots.is_contiguous() and out_dist.is_contiguous()
        _native.check(self._h, self._L.pwknn_search_device(self._h, nq, queries_bf16.data_ptr(),
                                                           queries_bf16.stride(0), int(k), out_slots.data_ptr(),
                                                           out_dist.data_ptr(), stream_ptr or None))

    def search_device_packed(self, queries_bf16, k: int, id_offset: int, out_packed, stream_ptr: int = 0):
        """out_packed int64 [nq, k, 2] = {slot + id_offset | -1, distance bits}: this shard's all-gather payload."""
        nq = queries_bf16.shape[0]
        assert out_packed.is_contiguous() and out_packed.numel() == nq * k * 2
        _native.check(self._h, self._L.pwknn_search_device_packed(self._h, nq, queries_bf16.data_ptr(),
                                                                  queries_bf16.stride(0), int(k), int(id_offset),
                                                                  out_packed.data_ptr(), stream_ptr or None))

    def search_host_packed(self, queries_f64: np.ndarray, k: int, id_offset: int, out_packed, stream_ptr: int = 0):
        """Host f64 queries in, this shard's packed top-k left in HBM (out_packed int64 [nq, k, 2]); asynchronous."""
        queries_f64 = np.ascontiguousarray(queries_f64, dtype=np.float64)
        nq = queries_f64.shape[0]
        assert queries_f64.shape == (nq, self.dimensions) and out_packed.is_contiguous() and out_packed.numel() == nq * k * 2
        _native.check(self._h, self._L.pwknn_search_host_packed(self._h, nq, queries_f64.ctypes.data, int(k),
                                                                int(id_offset), out_packed.data_ptr(), stream_ptr or None))

    def merge_shards_device(self, gathered, out_ids, out_dist, stream_ptr: int = 0):
        """gathered int64 [world, nq, k, 2] (all-gathered packed results) -> out_ids int64 [nq,k], out_dist f64 [nq,k]."""
        world, nq, k, two = gathered.shape
        assert two == 2 and gathered.is_contiguous() and out_ids.is_contiguous() and out_dist.is_contiguous()
        _native.check(self._h, self._L.pwknn_merge_shards_device(self._h, world, nq, k, gathered.data_ptr(),
                                                                 out_ids.data_ptr(), out_dist.data_ptr(),
                                                                 stream_ptr or None))

    def debug_scores_device(self, queries_bf16, out_scores_f32, stream_ptr: int = 0):
        nq = queries_bf16.shape[0]
        _native.check(self._h, self._L.pwknn_debug_scores_device(self._h, nq, queries_bf16.data_ptr(),
                                                                 queries_bf16.stride(0), out_scores_f32.data_ptr(),
                                                                 out_scores_f32.stride(0), stream_ptr or None))

    def slots_to_keys(self, slots: np.ndarray) -> np.ndarray:
        slots = np.ascontiguousarray(slots, dtype=np.uint32).reshape(-1)
        out = np.zeros((len(slots), 2), dtype=np.uint64)
        _native.check(self._h, self._L.pwknn_slots_to_keys(self._h, len(slots), slots.ctypes.data, out.ctypes.data))
        return out

    # -- trait-shaped API
    def _vec_ok(self, v) -> bool:
        try:
            return len(v) == self.dimensions
        except TypeError:
            return False

    def add(self, add_data: list[tuple[int, Any]]) -> list[tuple[int, Optional[Exception]]]:
        if not add_data:
            return []
        good = [i for i, (_k, v) in enumerate(add_data) if self._vec_ok(v)]
        ret: list = [None] * len(add_data)
        for i, (k, v) in enumerate(add_data):
            if not self._vec_ok(v):
                ret[i] = (k, ValueError_(f"vector length {len(v) if hasattr(v, '__len__') else '?'} != "
                                         f"dimensions {self.dimensions}"))
        if good:
            keys = keys_to_array([add_data[i][0] for i in good])
            rows = np.array([np.asarray(add_data[i][1], dtype=np.float64) for i in good], dtype=np.float64)
            st = self.add_arrays(keys, rows.reshape(len(good), self.dimensions))
            for i, s in zip(good, st):
                k = add_data[i][0]
                ret[i] = (k, None if s == _native.ROW_OK else ValueError_("non-finite vector component"))
        return ret

    def remove(self, keys: list[int]) -> list[tuple[int, Optional[Exception]]]:
        if not keys:
            return []
        st = self.remove_arrays(keys_to_array(keys))
        return [(k, None if s == _native.ROW_OK else MissingKey(k)) for k, s in zip(keys, st)]

    def search(self, queries: list[tuple[int, Any, int]]):
        if not queries:
            return []
        ret: list = [None] * len(queries)
        good = []
        for i, (k, v, lim) in enumerate(queries):
            if not self._vec_ok(v):
                ret[i] = (k, ValueError_(f"query length != dimensions {self.dimensions}"))
            elif lim < 0 or lim > 0xFFFFFFFF:
                ret[i] = (k, ValueError_("limit out of range"))
            else:
                good.append(i)
        if good:
            q = np.array([np.asarray(queries[i][1], dtype=np.float64) for i in good], dtype=np.float64)
            limits = np.minimum(np.array([queries[i][2] for i in good], dtype=np.uint64), self.current_size).astype(np.uint32)
            ok, sc, cnt, st = self.search_arrays(q.reshape(len(good), self.dimensions), limits)
            off = 0
            for j, i in enumerate(good):
                k = queries[i][0]
                if st[j] != _native.ROW_OK:
                    ret[i] = (k, ValueError_("non-finite query component"))
                else:
                    c = int(cnt[j])
                    ret[i] = (k, [((int(ok[off + r, 1]) << 64) | int(ok[off + r, 0]), float(sc[off + r]))
                                  for r in range(c)])
                off += int(limits[j])
        return ret


# ----------------------------------------------------------------------------- Unpack (mod.rs:338-372)
def unpack_vec_f64(value) -> list[float] | np.ndarray:
    """`impl Unpack<Vec<f64>> for Value` mod.rs:345-358."""
    if isinstance(value, np.ndarray):
        if value.dtype.kind in "iu":  # Value::IntArray
            return value.astype(np.float64).reshape(-1)
        if value.dtype.kind == "f":  # Value::FloatArray
            return value.astype(np.float64).reshape(-1)
        raise TypeMismatch("vector of floats", value)
    if isinstance(value, (tuple, list)):  # Value::Tuple of Value::Float only
        out = []
        for x in value:
            if isinstance(x, bool) or not isinstance(x, (float, np.floating)):
                raise TypeMismatch("float", x)
            out.append(float(x))
        return out
    raise TypeMismatch("vector of floats", value)


def unpack_json(value):
    """`impl Unpack<Variable> for Value` mod.rs:361-366: Value::Json -> JMESPath variable."""
    if isinstance(value, (dict, list, int, float, bool)) or value is None:
        return value
    if isinstance(value, str):  # a JSON document passed as text
        try:
            return json.loads(value)
        except json.JSONDecodeError as e:
            raise TypeMismatch("Json", value) from e
    raise TypeMismatch("Json", value)


@functools.lru_cache(maxsize=1024)
def glob_pattern_regex(pattern: str):
    """`glob::Pattern::new(pattern)` with the default `MatchOptions` that `Pattern::matches` uses (glob 0.3: case
    sensitive, `*` and `?` also match `/`, no special leading dot) as a compiled regex: `?` one character, `*` any
    run, `[...]` / `[!...]` classes, and the recursive `**`, which must be a whole path component (`**/` = zero or
    more directories) -- anything else is a pattern error, as in the crate."""
    out, i, n = [], 0, len(pattern)
    while i < n:
        c = pattern[i]
        if c == "*":
            j = i
            while j < n and pattern[j] == "*":
                j += 1
            if j - i >= 2:
                before_ok = i == 0 or pattern[i - 1] == "/"
                after_ok = j == n or pattern[j] == "/"
                if not (before_ok and after_ok):
                    raise ValueError("glob parsing error: recursive wildcards must form a single path component")
                if j < n:  # "**/": zero or more directories
                    out.append("(?:.*/)?")
                    j += 1
                else:
                    out.append(".*")
            else:
                out.append(".*")
            i = j
        elif c == "?":
            out.append(".")
            i += 1
        elif c == "[":
            j = i + 1
            if j < n and pattern[j] == "!":
                j += 1
            if j < n and pattern[j] == "]":
                j += 1
            while j < n and pattern[j] != "]":
                j += 1
            if j >= n:
                raise ValueError("glob parsing error: invalid range pattern")
            body = pattern[i + 1:j]
            neg = body.startswith("!")
            body = body[1:] if neg else body
            out.append("[" + ("^" if neg else "") + body.replace("\\", "\\\\").replace("^", "\\^") + "]")
            i = j + 1
        else:
            out.append(re.escape(c))
            i += 1
    return re.compile("".join(out), re.DOTALL)


class GlobPatternError(jmespath.exceptions.JMESPathError):
    """glob::PatternError surfaced through the JMESPath runtime (mod.rs:270-277)."""


class _GlobFunctions(_jp_functions.Functions):
    """JMESPath runtime with the extra `globmatch(pattern, string)` function, mod.rs:259-297."""

    @_jp_functions.signature({"types": ["string"]}, {"types": ["string"]})
    def _func_globmatch(self, pattern, path):
        # a malformed pattern is a JmespathError of THIS query only (mod.rs:270-277 maps glob::PatternError into
        # one; :432-455 / :480-484 turn it into that query's Err), never an exception of the whole search call
        try:
            rx = glob_pattern_regex(pattern)
        except (ValueError, re.error) as e:
            raise GlobPatternError(f"globmatch: bad pattern {pattern!r}: {e}") from e
        return rx.fullmatch(path) is not None


_JP_OPTIONS = jmespath.Options(custom_functions=_GlobFunctions())


# ----------------------------------------------------------------------------- DerivedFilteredSearchIndex
class AddDataEntry:  # mod.rs:28-32
    __slots__ = ("key", "data", "filter_data")

    def __init__(self, key, data, filter_data=None):
        self.key, self.data, self.filter_data = key, data, filter_data


class QueryEntry:  # mod.rs:34-39
    __slots__ = ("key", "data", "limit", "filter")

    def __init__(self, key, data, limit=None, filter=None):
        self.key, self.data, self.limit, self.filter = key, data, limit, filter


class DerivedFilteredSearchIndex:
    """mod.rs:384-604: turns a NonFilteringExternalIndex into an ExternalIndex.

    pushdown=True (default when the inner index offers `filter_create`, i.e. the B200 index): instead of the
    reference's post-filter + k, 2k, 4k ... re-search loop, every distinct filter expression becomes a row bitmask
    kept by the library (evaluated once per row, for new rows at add time) and a filtered query is ONE search.
    Same replies as the loop; a filter that does not evaluate to a boolean on some row falls back to the loop for
    that query, which keeps the reference's lazy error behaviour."""

    MAX_PUSHED_FILTERS = 64

    def __init__(self, index, pushdown: Optional[bool] = None):
        self.filter_data_map: dict[int, Any] = {}
        self.inner = index
        self.pushdown = hasattr(index, "filter_create") if pushdown is None else pushdown
        self._pushed: dict[str, tuple[int, Any]] = {}  # expression text -> (filter id, compiled); insertion = LRU order

    # ---- push-down bookkeeping
    def _eval_bool(self, expr, meta):
        r = expr.search(meta, options=_JP_OPTIONS)
        if r is True or r is False:
            return r
        raise ValueError_("jmespath filter expression did not return a boolean value")

    def _push_filter(self, text: str, expr, in_use: set):
        """-> filter id, or None when some row makes the expression non-boolean / fail, or the cache is full of
        filters this very batch needs (use the reference's lazy loop for that query then)."""
        hit = self._pushed.pop(text, None)
        if hit is not None:
            self._pushed[text] = hit  # most recently used
            in_use.add(text)
            return hit[0]
        keys, passes = [], []
        try:
            for key, meta in self.filter_data_map.items():
                keys.append(key)
                passes.append(self._eval_bool(expr, meta))
        except (DataError, jmespath.exceptions.JMESPathError):
            return None
        if len(self._pushed) >= self.MAX_PUSHED_FILTERS:
            old_text = next((t for t in self._pushed if t not in in_use), None)  # least recently used, not needed now
            if old_text is None:
                return None
            self.inner.filter_destroy(self._pushed.pop(old_text)[0])
        fid = self.inner.filter_create()
        self.inner.filter_set(fid, keys, passes)
        self._pushed[text] = (fid, expr)
        in_use.add(text)
        return fid

    def _update_pushed_for_new_rows(self, keys: list[int]):
        for text, (fid, expr) in list(self._pushed.items()):
            try:
                passes = [self._eval_bool(expr, self.filter_data_map[k]) if k in self.filter_data_map else False
                          for k in keys]
            except (DataError, jmespath.exceptions.JMESPathError):
                self.inner.filter_destroy(fid)  # no longer a total boolean function of the rows: back to the lazy loop
                del self._pushed[text]
                continue
            self.inner.filter_set(fid, keys, passes)

    def _handle_filter_and_unpack_data(self, entry: AddDataEntry):  # :403-408
        if entry.filter_data is not None:
            self.filter_data_map[entry.key] = unpack_json(entry.filter_data)
        return unpack_vec_f64(entry.data)

    def _make_query_tuple(self, query: QueryEntry):  # :410-430
        if query.limit is not None:
            lim = query.limit
            if isinstance(lim, bool) or not isinstance(lim, (int, np.integer)):
                raise TypeMismatch("int", lim)
            if lim < 0:
                raise ValueError_("out of range integral type conversion attempted")  # usize::try_from
            limit = int(lim)
        else:
            limit = 1  # :417-421
        point = unpack_vec_f64(query.data)
        if query.filter is None:
            flt = None
        else:
            if not isinstance(query.filter, str):
                raise TypeMismatch("string", query.filter)
            try:
                flt = jmespath.compile(query.filter)
            except jmespath.exceptions.JMESPathError as e:
                raise ValueError_(f"jmespath: {e}") from e
        return point, limit, flt

    def _filter_results(self, results, expr):  # :432-455
        kept = []
        for key, score in results:
            try:
                r = expr.search(self.filter_data_map[key], options=_JP_OPTIONS)
            except jmespath.exceptions.JMESPathError as e:
                raise ValueError_(f"jmespath: {e}") from e
            if r is True:
                kept.append((key, score))
            elif r is False:
                continue
            else:
                raise ValueError_("jmespath filter expression did not return a boolean value")
        return kept

    def add(self, add_data: list[AddDataEntry]):  # :512-530
        batch, error_rows = [], []
        for entry in add_data:
            try:
                batch.append((entry.key, self._handle_filter_and_unpack_data(entry)))
            except DataError as e:
                error_rows.append((entry.key, e))
        ret = list(self.inner.add(batch))
        if self.pushdown and self._pushed:
            self._update_pushed_for_new_rows([k for k, r in ret if r is None])
        ret.extend(error_rows)
        return ret

    def remove(self, keys):  # :532-537
        for key in keys:
            self.filter_data_map.pop(key, None)
        return self.inner.remove(list(keys))

    def search(self, query_data: list[QueryEntry]):  # :539-603
        ok_queries, responses = [], []
        for q in query_data:
            try:
                ok_queries.append((q.key, self._make_query_tuple(q)))
            except DataError as e:
                responses.append((q.key, e))
        filtering = [(k, p, lim, f) for k, (p, lim, f) in ok_queries if f is not None]
        regular = [(k, p, lim) for k, (p, lim, f) in ok_queries if f is None]
        if self.pushdown and filtering:
            texts = {q.key: q.filter for q in query_data}
            pushed, lazy, in_use = [], [], set()
            for k, p, lim, f in filtering:
                fid = self._push_filter(texts[k], f, in_use)
                (pushed if fid is not None else lazy).append((k, p, lim, f) if fid is None else (k, p, lim, fid))
            for key, res in self.inner.search_filtered(pushed):  # one pass, no k-doubling
                responses.append((key, res if isinstance(res, Exception) else tuple(res)))
            filtering = lazy
        for key, res in self.inner.search(regular):  # :558
            responses.append((key, res if isinstance(res, Exception) else tuple(res)))
        pending = [(k, p, lim, lim, f) for k, p, lim, f in filtering]  # (key, query, limit, upper_bound, expr)
        while pending:  # :583-601
            answers = self.inner.search([(k, p, ub) for k, p, _lim, ub, _f in pending])
            nxt = []
            for (k, p, lim, ub, f), (k2, results) in zip(pending, answers):  # :457-504
                assert k == k2
                if isinstance(results, Exception):
                    responses.append((k, results))
                    continue
                res_len = len(results)
                try:
                    filtered = self._filter_results(results, f)
                except DataError as e:
                    responses.append((k, e))
                    continue
                if len(filtered) >= lim or res_len < ub:
                    responses.append((k, tuple(filtered[:lim])))
                else:
                    nxt.append((k, p, lim, 2 * ub, f))
            pending = nxt
        return responses


# ----------------------------------------------------------------------------- IndexDerivedImpl
class IndexDerivedImpl:
    """mod.rs:51-205.  Accessors are callables row -> value (None-returning option accessors allowed)."""

    def __init__(self, inner, error_logger: Optional[list], data_accessor: Callable,
                 filter_data_accessor: Optional[Callable], query_accessor: Callable,
                 query_limit_accessor: Optional[Callable], query_filter_accessor: Optional[Callable]):
        self.inner = inner
        self.error_log = error_logger if error_logger is not None else []
        self.data_accessor = data_accessor
        self.filter_data_accessor = filter_data_accessor
        self.query_accessor = query_accessor
        self.query_limit_accessor = query_limit_accessor
        self.query_filter_accessor = query_filter_accessor

    def _log(self, res):
        for _key, r in res:
            if isinstance(r, Exception):
                self.error_log.append(r)

    def take_updates(self, data: list[tuple[int, Any, int]]):  # :110-151
        to_insert, to_remove = [], []
        for k, v, diff in data:
            if diff == 0:  # :113
                continue
            d = self.data_accessor(v)
            fd = self.filter_data_accessor(v) if self.filter_data_accessor else None
            if d is ERROR or fd is ERROR:  # :117-124
                self.error_log.append(ErrorInIndexUpdate())
                continue
            if diff < 0:
                to_remove.append(k)
            else:
                to_insert.append(AddDataEntry(k, d, fd))
        self._log(self.inner.remove(to_remove))  # remove THEN add, :144-150
        self._log(self.inner.add(to_insert))

    def search(self, queries: list[tuple[int, Any, int]]):  # :153-204
        ok, errors = [], []
        for k, v, _diff in queries:
            vec = self.query_accessor(v)
            lim = self.query_limit_accessor(v) if self.query_limit_accessor else None
            flt = self.query_filter_accessor(v) if self.query_filter_accessor else None
            if vec is ERROR or lim is ERROR or flt is ERROR:
                self.error_log.append(ErrorInIndexSearch())
                errors.append(k)
            else:
                ok.append(QueryEntry(k, vec, lim, flt))
        responses = {}
        for key, res in self.inner.search(ok):
            if isinstance(res, Exception):
                self.error_log.append(res)
                responses[key] = ERROR
            else:
                responses[key] = res
        for k in errors:
            responses[k] = ERROR
        return [(k, (v, responses.pop(k, ERROR)), diff) for k, v, diff in queries]


# ----------------------------------------------------------------------------- the operator
def _freeze(v):
    if isinstance(v, np.ndarray):
        return ("nd", v.dtype.str, v.shape, v.tobytes())
    if isinstance(v, dict):
        return tuple(sorted((k, _freeze(x)) for k, x in v.items()))
    if isinstance(v, (list, tuple)):
        return tuple(_freeze(x) for x in v)
    return v


def batch_by_time(entries: Iterable[tuple[Any, Any, int, int]]):
    """operators/utils.rs:111-137 over an arranged (consolidated, key-sorted) batch:
    -> {time: [(key, val, diff)]} ascending time; inside a time, ascending key."""
    acc: dict = {}
    for key, val, time, diff in entries:
        ident = (time, key, _freeze(val))
        if ident in acc:
            acc[ident][3] += diff
        else:
            acc[ident] = [key, val, time, diff]
    by_time = defaultdict(list)
    for key, val, time, diff in acc.values():
        if diff != 0:
            by_time[time].append((key, val, diff))
    return {t: sorted(v, key=lambda e: e[0]) for t, v in sorted(by_time.items())}


def _net_updates(batches: list[list[tuple[int, Any, int]]], is_error: Callable[[Any], bool]):
    """Net effect of several consecutive per-timestamp update batches (oldest first) as ONE batch for `take_updates`
    (which removes first, then adds): per key, a removal is needed iff the key's first event in the window is a
    removal (it was live before the window), an insertion iff its last event is one (with that last value); whatever
    happened in between never reaches the index.  Inside one timestamp a key's removal precedes its insertion (the
    upsert pattern, mod.rs:144-150).  Error rows pass through untouched (take_updates logs and skips them)."""
    first_is_remove, last_add, order, passthrough = {}, {}, [], []
    for batch in batches:
        for k, v, diff in sorted(batch, key=lambda e: (e[0], e[2] > 0)):  # per key: -diff before +diff
            if diff == 0:
                continue
            if is_error(v):
                passthrough.append((k, v, diff))
                continue
            if k not in first_is_remove:
                first_is_remove[k] = (diff < 0, v)
                order.append(k)
            last_add[k] = v if diff > 0 else None
    merged = list(passthrough)
    for k in order:
        was_live, v_removed = first_is_remove[k]
        if was_live:
            merged.append((k, v_removed, -1))
        if last_add[k] is not None:
            merged.append((k, last_add[k], 1))
    return merged


def use_external_index_as_of_now(index_stream, query_stream, index: IndexDerivedImpl, coalesce: bool = False,
                                 stats: Optional[dict] = None):
    """external_index.rs:76-158: per timestamp (ascending) `take_updates(updates)` then
    `search(queries)`; one reply per query row with the query's key, time and diff.

    index_stream / query_stream: iterables of (key, row, time, diff).
    Returns [(key, (row, reply), time, diff)].

    coalesce=True emulates the operator-level batching of SURVEY.md section 8(f-4), which needs the engine to be built
    for real: the update batches of consecutive timestamps WITHOUT queries are applied as one netted batch right
    before the next timestamp that has queries (as-of-now replies cannot tell the difference: nobody looks at the index
    in between; only the slot layout, i.e. the order of exact distance ties, may differ).  `stats` (optional dict)
    receives the number of index calls and rows with / without coalescing."""
    upd = batch_by_time(index_stream)
    qry = batch_by_time(query_stream)
    out = []
    st = {"update_calls": 0, "update_rows": 0, "update_rows_uncoalesced": 0, "timestamps_with_updates": 0}

    def is_error(v):
        d = index.data_accessor(v)
        fd = index.filter_data_accessor(v) if index.filter_data_accessor else None
        return d is ERROR or fd is ERROR

    pending: list = []

    def flush():
        if not pending:
            return
        batch = _net_updates(pending, is_error) if len(pending) > 1 else pending[0]
        index.take_updates(batch)
        st["update_calls"] += 1
        st["update_rows"] += len(batch)
        pending.clear()

    for t in sorted(set(upd) | set(qry)):
        if t in upd:
            st["timestamps_with_updates"] += 1
            st["update_rows_uncoalesced"] += len(upd[t])
            pending.append(upd[t])
            if not coalesce:
                flush()
        if t in qry:
            flush()  # updates at time t are visible to the queries of time t (external_index.rs:140-151)
            for k, v, diff in index.search(qry[t]):
                out.append((k, v, t, diff))
    flush()
    if stats is not None:
        stats.update(st)
    return out
