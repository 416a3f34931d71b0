"""BASELINE.json configs[2] (VectorStoreServer live RAG): 5 M x 1536 corpus row-sharded over the ranks, an insert
stream of 10 k rows/s (500 new f32 rows every 50 ms tick, routed to the owning shard by key) interleaved with
query batches (k = 8).  Run under torchrun; every rank applies the deltas it owns, answers the replicated query
batch on its shard and joins ONE NCCL all-gather + merge per batch.  Reports sustained queries/s while the
insert stream is being absorbed, and the time a tick spends in the delta path."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pathway_b200.engine import BruteForceKnnMetricKind  # noqa: E402
from pathway_b200.external_index import B200BruteForceKNNIndex  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=5_000_000)
ap.add_argument("--dims", type=int, default=1536)
ap.add_argument("--k", type=int, default=8)
ap.add_argument("--qbatch", type=int, default=1024)
ap.add_argument("--ticks", type=int, default=40)
ap.add_argument("--rows-per-tick", type=int, default=500)
args = ap.parse_args()
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
N, D, K, Q = args.rows, args.dims, args.k, args.qbatch
shard = N // world
ix = B200BruteForceKNNIndex(D, shard + args.ticks * args.rows_per_tick, metric=BruteForceKnnMetricKind.COS, device=local)
CH = 125_000
for c in range(shard // CH):
    g = torch.Generator(device=dev); g.manual_seed(1000 * rank + c)
    rows = torch.randn((CH, D), generator=g, device=dev).to(torch.bfloat16)
    ids = (np.arange(c * CH, (c + 1) * CH, dtype=np.uint64) * np.uint64(world) + np.uint64(rank))  # key % world == rank
    keys = np.zeros((CH, 2), dtype=np.uint64); keys[:, 0] = ids
    torch.cuda.synchronize(dev); ix.add_device(keys, rows); del rows
gq = torch.Generator(device=dev); gq.manual_seed(5)
q = torch.randn((Q, D), generator=gq, device=dev).to(torch.bfloat16)
packed = torch.empty((Q, K, 2), dtype=torch.int64, device=dev)
gathered = torch.empty((world, Q, K, 2), dtype=torch.int64, device=dev)
out_ids = torch.empty((Q, K), dtype=torch.int64, device=dev); out_d = torch.empty((Q, K), dtype=torch.float64, device=dev)
ext = torch.cuda.ExternalStream(ix.stream_ptr, device=dev)
rng = np.random.default_rng(99)  # same stream on every rank: every rank sees every delta (broadcast contract)
next_id = N


def query_batch():
    # slot -> global id needs the key map on the host only for final presentation; ids here are shard slots + offset
    ix.search_device_packed(q, K, rank * (1 << 40), packed)
    if world > 1:
        dist.all_gather_into_tensor(gathered, packed)
        ix.merge_shards_device(gathered, out_ids, out_d)


with torch.cuda.stream(ext):
    for _ in range(3):
        query_batch()
    ext.synchronize()
    if world > 1:
        dist.barrier()
    t_delta = 0.0; nq = 0; t0 = time.perf_counter()
    for tick in range(args.ticks):
        new = rng.standard_normal((args.rows_per_tick, D)).astype(np.float32)
        ids = np.arange(next_id, next_id + args.rows_per_tick, dtype=np.uint64); next_id += args.rows_per_tick
        mine = (ids % np.uint64(world)) == np.uint64(rank)
        td = time.perf_counter()
        if mine.any():
            keys = np.zeros((int(mine.sum()), 2), dtype=np.uint64); keys[:, 0] = ids[mine]
            ix.add_arrays(keys, new[mine])
        t_delta += time.perf_counter() - td
        tick_end = t0 + (tick + 1) * 0.050  # 50 ms autocommit window (vector_store.py:103)
        while True:  # answer query batches until the tick is over
            query_batch(); nq += Q
            ext.synchronize()
            if time.perf_counter() >= tick_end:
                break
    ext.synchronize()
    if world > 1:
        dist.barrier()
    wall = time.perf_counter() - t0
if rank == 0:
    print(json.dumps(dict(config="C3 live RAG", world=world, rows=N, dims=D, k=K, qbatch=Q, ticks=args.ticks,
                          insert_rows_per_s=args.rows_per_tick / 0.050, delta_ms_per_tick_rank0=round(t_delta / args.ticks * 1e3, 3),
                          qps_under_inserts=round(nq / wall, 1), ms_per_batch=round(wall / (nq / Q) * 1e3, 3),
                          final_rows_rank0=ix.current_size)))
ix.close()
if world > 1:
    dist.barrier(); dist.destroy_process_group()
