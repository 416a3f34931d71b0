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
dist.barrier()
dist.destroy_process_group()
