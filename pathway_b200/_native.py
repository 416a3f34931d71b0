"""ctypes binding of libpwknn.so (the C ABI in include/pwknn.h).

There is no CPU fallback: importing works everywhere (so the symbol test can run
without a GPU), but creating an index fails loudly when the CUDA library or a
B200 is missing.
"""

from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpwknn.so")

OK, E_INVALID, E_CUDA, E_NOGPU, E_OOM = 0, 1, 2, 3, 4
ROW_OK, ROW_MISSING_KEY, ROW_BAD_VECTOR, ROW_DEVICE_ERROR = 0, 1, 2, 3
INVALID_SLOT = 0xFFFFFFFF

# name -> (restype, argtypes); must list every symbol include/pwknn.h declares
_u64, _u32, _i32, _i64, _vp, _cp = C.c_uint64, C.c_uint32, C.c_int32, C.c_int64, C.c_void_p, C.c_char_p
SIGNATURES = {
    "pwknn_abi_version": (_u32, []),
    "pwknn_last_global_error": (_cp, []),
    "pwknn_create": (C.c_int, [_u32, _u64, _u64, C.c_int, C.c_int, C.POINTER(_vp)]),
    "pwknn_create_ex": (C.c_int, [_u32, _u64, _u64, C.c_int, C.c_int, _u32, C.POINTER(_vp)]),
    "pwknn_create_multi": (C.c_int, [_u32, _u64, _u64, C.c_int, _vp, _u32, _u32, C.POINTER(_vp)]),
    "pwknn_multi_set_weights": (C.c_int, [_vp, _u32, _vp]),
    "pwknn_shard": (_vp, [_vp, _u32]),
    "pwknn_shards_of_keys": (C.c_int, [_vp, _u64, _vp, _vp]),
    "pwknn_destroy": (C.c_int, [_vp]),
    "pwknn_add": (C.c_int, [_vp, _u64, _vp, _vp, _vp]),
    "pwknn_add_f32": (C.c_int, [_vp, _u64, _vp, _vp, _vp]),
    "pwknn_add_device_bf16": (C.c_int, [_vp, _u64, _vp, _vp, _u64, _vp, _vp]),
    "pwknn_remove": (C.c_int, [_vp, _u64, _vp, _vp]),
    "pwknn_search": (C.c_int, [_vp, _u64, _vp, _vp, _vp, _vp, _vp, _vp]),
    "pwknn_search_f32": (C.c_int, [_vp, _u64, _vp, _vp, _vp, _vp, _vp, _vp]),
    "pwknn_filter_create": (C.c_int, [_vp, C.POINTER(_u32)]),
    "pwknn_filter_destroy": (C.c_int, [_vp, _u32]),
    "pwknn_filter_set": (C.c_int, [_vp, _u32, _u64, _vp, _vp, _vp]),
    "pwknn_search_filtered": (C.c_int, [_vp, _u64, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "pwknn_search_device": (C.c_int, [_vp, _u64, _vp, _u64, _u32, _vp, _vp, _vp]),
    "pwknn_search_device_packed": (C.c_int, [_vp, _u64, _vp, _u64, _u32, _i64, _vp, _vp]),
    "pwknn_merge_shards_device": (C.c_int, [_vp, _u32, _u64, _u32, _vp, _vp, _vp, _vp]),
    "pwknn_search_host_packed": (C.c_int, [_vp, _u64, _vp, _u32, _i64, _vp, _vp]),
    "pwknn_snapshot_save": (C.c_int, [_vp, _cp]),
    "pwknn_snapshot_load": (C.c_int, [_cp, C.c_int, C.POINTER(_vp)]),
    "pwknn_snapshot_load_multi": (C.c_int, [_cp, _vp, _u32, C.POINTER(_vp)]),
    "pwknn_slots_to_keys": (C.c_int, [_vp, _u64, _vp, _vp]),
    "pwknn_size": (_u64, [_vp]),
    "pwknn_allocated": (_u64, [_vp]),
    "pwknn_dimensions": (_u32, [_vp]),
    "pwknn_device": (C.c_int, [_vp]),
    "pwknn_stream": (_vp, [_vp]),
    "pwknn_last_error": (_cp, [_vp]),
    "pwknn_read_rows": (C.c_int, [_vp, _u64, _u64, _vp]),
    "pwknn_set_option": (C.c_int, [_vp, _cp, _i64]),
    "pwknn_get_stat": (_i64, [_vp, _cp]),
    "pwknn_rrf_fuse": (C.c_int, [_u32, _u64, _vp, _vp, _vp, C.c_double, _vp, _vp, _vp, _vp]),
    "pwknn_debug_scores_device": (C.c_int, [_vp, _u64, _vp, _u64, _vp, _u64, _vp]),
}

_lib = None


class PwknnError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libpwknn error {code}: {msg}")
        self.code = code


def lib():
    """Load libpwknn.so (raises if it was not built -- there is no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). pathway_b200 has no CPU fallback."
            )
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(h, rc: int):
    if rc != OK:
        L = lib()
        msg = (L.pwknn_last_error(h) if h else L.pwknn_last_global_error()) or b""
        raise PwknnError(rc, msg.decode("utf-8", "replace"))
