"""BASELINE.json configs[4] (incremental churn): 1 M x 768 corpus, rounds of {remove 1 % random keys, re-add them
with new vectors, answer a 1024-query batch}; 50 rounds = 50 % churn.  Reports the delta-path rate
(rows_written*D*2 + rows_moved*D*2*2 bytes of HBM traffic / time in remove+add) for host f64, host f32 and
device-resident bf16 deltas, plus the query rate in between.  One GPU (the 8-GPU variant shards keys)."""
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
ap.add_argument("--rows", type=int, default=1_000_000)
ap.add_argument("--dims", type=int, default=768)
ap.add_argument("--rounds", type=int, default=50)
ap.add_argument("--frac", type=float, default=0.01)
ap.add_argument("--qbatch", type=int, default=1024)
args = ap.parse_args()
dev = torch.device("cuda", 0)
N, D, Q = args.rows, args.dims, args.qbatch
rng = np.random.default_rng(0)
for mode in ("f64", "f32", "device_bf16"):
    ix = B200BruteForceKNNIndex(D, N, metric=BruteForceKnnMetricKind.COS, device=0)
    CH = 250_000
    for c in range(N // CH):
        g = torch.Generator(device=dev); g.manual_seed(7 + c)
        rows = torch.randn((CH, D), generator=g, device=dev, dtype=torch.float32).to(torch.bfloat16)
        keys = np.zeros((CH, 2), dtype=np.uint64); keys[:, 0] = np.arange(c * CH, (c + 1) * CH, dtype=np.uint64)
        torch.cuda.synchronize(dev); ix.add_device(keys, rows); del rows
    q = rng.standard_normal((Q, D)).astype(np.float32)
    limits = np.full(Q, 10, dtype=np.uint32)
    ix.search_arrays(q, limits)
    n_delta = int(N * args.frac)
    t_delta = t_query = 0.0
    moved_total = 0
    for r in range(args.rounds):
        victims = rng.choice(N, n_delta, replace=False).astype(np.uint64)
        keys = np.zeros((n_delta, 2), dtype=np.uint64); keys[:, 0] = victims
        new_f32 = rng.standard_normal((n_delta, D)).astype(np.float32)
        if mode == "f64":
            new = new_f32.astype(np.float64)
        elif mode == "f32":
            new = new_f32
        else:
            new = torch.from_numpy(new_f32).to(dev).to(torch.bfloat16); torch.cuda.synchronize(dev)
        # rows moved by swap-remove: every victim that is not among the trailing rows at its turn moves one row
        t0 = time.perf_counter()
        st = ix.remove_arrays(keys)
        if mode == "device_bf16":
            st2 = ix.add_device(keys, new)
        else:
            st2 = ix.add_arrays(keys, new)
        t1 = time.perf_counter()
        assert (st == 0).all() and (st2 == 0).all() and ix.current_size == N
        ix.search_arrays(q, limits)
        t2 = time.perf_counter()
        t_delta += t1 - t0; t_query += t2 - t1
        moved_total += n_delta  # upper bound: <= one moved row per removal
    written = args.rounds * n_delta
    bytes_hbm = written * D * 2 + moved_total * D * 2 * 2
    print(json.dumps(dict(mode=mode, rows=N, dims=D, rounds=args.rounds, deltas_per_round=n_delta,
                          delta_ms_per_round=round(t_delta / args.rounds * 1e3, 3),
                          delta_rows_per_s=round(written / t_delta, 1),
                          delta_path_gbs=round(bytes_hbm / t_delta / 1e9, 3),
                          host_input_gbs=round(written * D * (8 if mode == "f64" else 4 if mode == "f32" else 2) / t_delta / 1e9, 3),
                          query_ms_per_batch=round(t_query / args.rounds * 1e3, 3),
                          qps=round(Q * args.rounds / t_query, 1))), flush=True)
    ix.close()
