"""torchrun --nproc-per-node N scripts/sharded_nccl_check.py : ShardedKnnIndex over B200 shards + NCCL
all-gather/merge against a single CPU oracle (every rank checks)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import brute_force_oracle as bfo  # noqa: E402
from oracle_adapter import OracleInner  # noqa: E402
from pathway_b200.engine import BruteForceKnnMetricKind  # noqa: E402
from pathway_b200.external_index import B200BruteForceKNNIndex  # noqa: E402
from pathway_b200.sharded import ShardedKnnIndex  # noqa: E402

local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
for metric in (BruteForceKnnMetricKind.COS, BruteForceKnnMetricKind.L2SQ):
    rng = np.random.default_rng(5)
    d = 100
    sharded = ShardedKnnIndex(B200BruteForceKNNIndex(d, 8, metric=metric, device=local), device=torch.device("cuda", local))
    single = OracleInner(d, 8, 1024 * 128, metric, impl="c")
    live = []
    for step in range(5):
        rows = bfo.bf16_round(rng.standard_normal((3000, d)))
        batch = [(int(rng.integers(1, 2 ** 62)), rows[i]) for i in range(3000)]
        rm = [live.pop(int(rng.integers(0, len(live)))) for _ in range(min(len(live), 700))]
        sharded.remove(rm), single.remove(rm)
        sharded.add(batch), single.add(batch)
        live += [k for k, _ in batch]
        assert sharded.global_size() == single.current_size
        qs = [(i, bfo.bf16_round(rng.standard_normal(d)), int(rng.integers(0, 15))) for i in range(50)]
        for (k1, r1), (k2, r2) in zip(sharded.search(qs), single.search(qs)):
            assert [m for m, _ in r1] == [m for m, _ in r2], (rank, step)
            np.testing.assert_allclose([s for _, s in r1], [s for _, s in r2], rtol=1e-9, atol=1e-12)
    print(f"rank {rank}/{world} metric {metric.name}: sharded NCCL path matches the oracle "
          f"(local rows {sharded.local_size} of {single.current_size})", flush=True)
    sharded.inner.close()
# device tier: packed output + NCCL all-gather + library merge kernel  ==  torch stable-sort merge
from pathway_b200.sharded import allgather_merge_topk  # noqa: E402

dev = torch.device("cuda", local)
n_loc, d, Q, K = 20_000, 128, 300, 10
g = torch.Generator(device=dev); g.manual_seed(100 + rank)
rows = torch.randn((n_loc, d), generator=g, device=dev).to(torch.bfloat16)
rows[5] = rows[3]  # an exact duplicate inside the shard: tie broken by position
keys = np.zeros((n_loc, 2), dtype=np.uint64); keys[:, 0] = np.arange(rank * n_loc, (rank + 1) * n_loc, dtype=np.uint64)
ix = B200BruteForceKNNIndex(d, n_loc, metric=BruteForceKnnMetricKind.COS, device=local)
torch.cuda.synchronize(dev)
ix.add_device(keys, rows)
gq = torch.Generator(device=dev); gq.manual_seed(7)
q = torch.randn((Q, d), generator=gq, device=dev).to(torch.bfloat16)
q[0] = rows[3] if rank == 0 else q[0]
dist.broadcast(q, src=0)
ext = torch.cuda.ExternalStream(ix.stream_ptr, device=dev)
with torch.cuda.stream(ext):
    slots = torch.empty((Q, K), dtype=torch.int32, device=dev); dd = torch.empty((Q, K), dtype=torch.float64, device=dev)
    ix.search_device(q, K, slots, dd)
    ref_ids, ref_d = allgather_merge_topk((slots.to(torch.int64) + rank * n_loc).unsqueeze(-1), dd, K)
    packed = torch.empty((Q, K, 2), dtype=torch.int64, device=dev)
    gathered = torch.empty((world, Q, K, 2), dtype=torch.int64, device=dev)
    out_ids = torch.empty((Q, K), dtype=torch.int64, device=dev); out_d = torch.empty((Q, K), dtype=torch.float64, device=dev)
    ix.search_device_packed(q, K, rank * n_loc, packed)
    dist.all_gather_into_tensor(gathered, packed)
    ix.merge_shards_device(gathered, out_ids, out_d)
    ext.synchronize()
assert (out_ids == ref_ids[..., 0]).all() and (out_d == ref_d).all(), "library shard merge != torch merge"
assert out_ids[0, 0].item() == 3 and out_ids[0, 1].item() == 5
print(f"rank {rank}: packed search + all-gather + merge_shards kernel matches the torch merge", flush=True)
ix.close()
dist.barrier()
dist.destroy_process_group()
