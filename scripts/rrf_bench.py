"""Times pwknn_rrf_fuse on the C4 reply shape -- Q=256 queries, R retrievers x k=100 ids -- next to
the restated reference dataflow (oracle/hybrid_oracle.py, pure Python; the reference runs these steps as
differential-dataflow operators, which cannot be built here).  Host only: no GPU needed."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import hybrid_oracle as ho  # noqa: E402
from pathway_b200 import _native  # noqa: E402


def main():
    lib = _native.lib()
    rng = np.random.default_rng(0)
    nq, k = 256, 100
    for n_lists in (2, 4):
        keys = [np.zeros((nq * k, 2), dtype=np.uint64) for _ in range(n_lists)]
        for a in keys:
            for q in range(nq):
                a[q * k:(q + 1) * k, 0] = rng.choice(400, size=k, replace=False)
        cnt = [np.full(nq, k, dtype=np.uint32) for _ in range(n_lists)]
        lim = np.full(nq, k, dtype=np.uint32)
        out_k = np.zeros((nq * k, 2), dtype=np.uint64)
        out_s = np.zeros(nq * k)
        out_c = np.zeros(nq, dtype=np.uint32)
        vp = C.c_void_p * n_lists
        a1, a2 = vp(*[a.ctypes.data for a in keys]), vp(*[a.ctypes.data for a in cnt])
        reps = 200
        for it in range(reps + 20):
            if it == 20:
                t0 = time.perf_counter()
            lib.pwknn_rrf_fuse(n_lists, nq, a1, a2, None, 60.0, lim.ctypes.data, out_k.ctypes.data,
                               out_s.ctypes.data, out_c.ctypes.data)
        fuse_ms = (time.perf_counter() - t0) / reps * 1e3
        d = [{q: [(int(x), 0.0) for x in a[q * k:(q + 1) * k, 0]] for q in range(nq)} for a in keys]
        t0 = time.perf_counter()
        ho.combine_results(d, k, 60.0)
        py_ms = (time.perf_counter() - t0) * 1e3
        print(json.dumps({"queries": nq, "k": k, "retrievers": n_lists, "pwknn_rrf_fuse_ms": round(fuse_ms, 3),
                          "ids_per_s": round(nq * k * n_lists / fuse_ms * 1e3),
                          "restated_dataflow_python_ms": round(py_ms, 1)}))


if __name__ == "__main__":
    main()
