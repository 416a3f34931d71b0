"""Filter push-down vs the reference's post-filter + k-doubling loop, both over the B200 index (one GPU):
200 k x 768 rows with JSON metadata, 256 queries, k = 10, filters of selectivity 50 % / 5 % / 1 %."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pathway_b200.engine import BruteForceKnnMetricKind  # noqa: E402
from pathway_b200.external_index import (AddDataEntry, B200BruteForceKNNIndex, DerivedFilteredSearchIndex,  # noqa: E402
                                         QueryEntry)

N, D, Q, K = 200_000, 768, 256, 10
rng = np.random.default_rng(0)
rows = rng.standard_normal((N, D)).astype(np.float32)
indices = {}
for name, pd_ in (("pushdown", True), ("k_doubling_loop", False)):
    gpu = B200BruteForceKNNIndex(D, N, metric=BruteForceKnnMetricKind.COS, device=0)
    ix = DerivedFilteredSearchIndex(gpu, pushdown=pd_)
    t0 = time.perf_counter()
    for c0 in range(0, N, 20000):
        ix.add([AddDataEntry(i, rows[i], {"b": i % 100}) for i in range(c0, c0 + 20000)])
    indices[name] = (ix, gpu, time.perf_counter() - t0)
qv = rng.standard_normal((Q, D)).astype(np.float32)
for sel, expr in ((0.5, "b < `50`"), (0.05, "b < `5`"), (0.01, "b == `7`")):
    out = {}
    for name, (ix, gpu, t_add) in indices.items():
        qs = [QueryEntry(i, qv[i], K, expr) for i in range(Q)]
        ix.search(qs)  # warm (builds the mask on the first use)
        t0 = time.perf_counter()
        res = ix.search(qs)
        out[name] = (time.perf_counter() - t0, dict(res))
    a, b = out["pushdown"][1], out["k_doubling_loop"][1]
    assert all([m for m, _ in a[k]] == [m for m, _ in b[k]] for k in a)
    print(json.dumps(dict(selectivity=sel, filter=expr, rows=N, dims=D, queries=Q, k=K,
                          pushdown_ms=round(out["pushdown"][0] * 1e3, 2),
                          k_doubling_loop_ms=round(out["k_doubling_loop"][0] * 1e3, 2),
                          speedup=round(out["k_doubling_loop"][0] / out["pushdown"][0], 1))), flush=True)
