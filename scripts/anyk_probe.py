"""Any-k cliff probe: host-tier pwknn_search on one GPU with mixed / large limits.
Round 1 ran the whole batch at the largest k (one k = 500 query sent 1023 k = 10 queries to f64 CUDA-core distances +
a full sort); per-query routing keeps the small ones on the tensor cores and answers the large one on the select path."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pathway_b200.engine import BruteForceKnnMetricKind  # noqa: E402
from pathway_b200.external_index import B200BruteForceKNNIndex  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=10_000_000)
ap.add_argument("--dims", type=int, default=768)
ap.add_argument("--reps", type=int, default=5)
args = ap.parse_args()
dev = torch.device("cuda", 0)
N, D = args.rows, args.dims
ix = B200BruteForceKNNIndex(D, N, metric=BruteForceKnnMetricKind.COS, device=0)
CH = 250_000
for c in range((N + CH - 1) // CH):
    m = min(CH, N - c * CH)
    g = torch.Generator(device=dev); g.manual_seed(1234 + c)
    rows = torch.randn((m, D), generator=g, device=dev, dtype=torch.float32).to(torch.bfloat16)
    keys = np.zeros((m, 2), dtype=np.uint64); keys[:, 0] = np.arange(c * CH, c * CH + m, dtype=np.uint64)
    torch.cuda.synchronize(dev)
    ix.add_device(keys, rows)
    del rows
rng = np.random.default_rng(5)
q = rng.standard_normal((1024, D)).astype(np.float32)


def run(name, limits, nq=1024):
    lim = np.asarray(limits, dtype=np.uint32)
    ix.search_arrays(q[:nq], lim)
    ts = []
    for _ in range(args.reps):
        t0 = time.perf_counter(); _, _, cnt, st = ix.search_arrays(q[:nq], lim); ts.append(time.perf_counter() - t0)
    assert (st == 0).all() and (cnt == np.minimum(lim, N)).all()
    print(json.dumps(dict(case=name, nq=nq, ms_median=round(sorted(ts)[len(ts) // 2] * 1e3, 3), ms_min=round(min(ts) * 1e3, 3),
                          last_tier=ix.stat("last_tier"), uncertain=ix.stat("last_uncertain"))), flush=True)


run("all k=10", [10] * 1024)
run("1023 x k=10 + 1 x k=500", [10] * 1023 + [500])
run("1016 x k=10 + 8 x k=200", [10] * 1016 + [200] * 8)
run("all k=100 (fused two-pass)", [100] * 1024)
run("64 x k=1000 (select)", [1000] * 64, nq=64)
run("256 x k=200 (select)", [200] * 256, nq=256)
run("1 x k=10000 (full sort)", [10000], nq=1)
ix.close()
