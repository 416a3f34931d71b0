"""ctypes binding of oracle/brute_force_ref.c (TEST INFRASTRUCTURE ONLY).

Used by tests (C-vs-numpy restatement agreement) and by bench.py as the timed
CPU baseline (`cpu_baseline.kind = "port"`)."""

from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libpworacle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "brute_force_ref.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        L.pwo_create.restype = C.c_void_p
        L.pwo_create.argtypes = [C.c_size_t, C.c_size_t, C.c_size_t, C.c_int]
        L.pwo_destroy.argtypes = [C.c_void_p]
        L.pwo_size.restype = C.c_size_t
        L.pwo_size.argtypes = [C.c_void_p]
        L.pwo_allocated.restype = C.c_size_t
        L.pwo_allocated.argtypes = [C.c_void_p]
        L.pwo_add.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        L.pwo_remove.restype = C.c_int
        L.pwo_remove.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
        L.pwo_search.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.c_void_p, C.c_int]
        L.pwo_rows.restype = C.c_void_p
        L.pwo_rows.argtypes = [C.c_void_p]
        L.pwo_max_threads.restype = C.c_int
        _lib = L
    return _lib


def _keys_array(keys) -> np.ndarray:
    a = np.empty((len(keys), 2), dtype=np.uint64)
    for i, k in enumerate(keys):
        a[i, 0] = int(k) & 0xFFFFFFFFFFFFFFFF
        a[i, 1] = int(k) >> 64
    return a


class CBruteForceKNNIndex:
    """Same surface as oracle.brute_force_oracle.BruteForceKNNIndex, backed by C."""

    def __init__(self, dimensions, reserved_space, auxiliary_space, metric):
        self._L = lib()
        self.dimensions = dimensions
        self._h = self._L.pwo_create(dimensions, reserved_space, auxiliary_space, metric)

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.pwo_destroy(self._h)
            self._h = None

    @property
    def current_size(self):
        return self._L.pwo_size(self._h)

    @property
    def current_allocated(self):
        return self._L.pwo_allocated(self._h)

    def add_arrays(self, keys_u64x2: np.ndarray, rows: np.ndarray):
        rows = np.ascontiguousarray(rows, dtype=np.float64)
        keys_u64x2 = np.ascontiguousarray(keys_u64x2, dtype=np.uint64)
        st = np.zeros(len(rows), dtype=np.int32)
        self._L.pwo_add(self._h, len(rows), keys_u64x2.ctypes.data, rows.ctypes.data, rows.shape[1],
                        st.ctypes.data)
        return st

    def add(self, add_data):
        if not add_data:
            return []
        keys = [k for k, _ in add_data]
        lens = {len(v) for _, v in add_data}
        assert len(lens) == 1, "C oracle binding needs equal-length rows per call"
        rows = np.array([list(v) for _, v in add_data], dtype=np.float64).reshape(len(keys), -1)
        self.add_arrays(_keys_array(keys), rows)
        return [(k, None) for k in keys]

    def remove(self, keys):
        if not keys:
            return []
        ka = _keys_array(keys)
        st = np.zeros(len(keys), dtype=np.int32)
        rc = self._L.pwo_remove(self._h, len(keys), ka.ctypes.data, st.ctypes.data)
        if rc != 0:
            raise RuntimeError("remove from empty/inconsistent index (reference panics)")
        return [(k, None if s == 0 else KeyError(k)) for k, s in zip(keys, st)]

    def search_arrays(self, queries: np.ndarray, limits: np.ndarray, workers: int = 1):
        queries = np.ascontiguousarray(queries, dtype=np.float64)
        limits = np.ascontiguousarray(limits, dtype=np.uint32)
        nq = len(queries)
        tot = int(limits.sum())
        out_keys = np.zeros((max(tot, 1), 2), dtype=np.uint64)
        out_scores = np.zeros(max(tot, 1), dtype=np.float64)
        out_counts = np.zeros(nq, dtype=np.uint32)
        self._L.pwo_search(self._h, nq, queries.ctypes.data, limits.ctypes.data, out_keys.ctypes.data,
                           out_scores.ctypes.data, out_counts.ctypes.data, workers)
        return out_keys, out_scores, out_counts

    def search(self, queries, workers: int = 1):
        if not queries:
            return []
        q = np.array([list(v) for _, v, _ in queries], dtype=np.float64).reshape(len(queries), -1)
        assert q.shape[1] == self.dimensions
        limits = np.array([k for _, _, k in queries], dtype=np.uint32)
        ok, os_, oc = self.search_arrays(q, limits, workers)
        ret, off = [], 0
        for (key, _, lim), cnt in zip(queries, oc):
            res = [((int(ok[off + r, 1]) << 64) | int(ok[off + r, 0]), float(os_[off + r])) for r in range(cnt)]
            ret.append((key, res))
            off += int(lim)
        return ret
