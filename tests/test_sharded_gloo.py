"""World-size-2 `gloo` test (CPU) of the N>1 path's host logic: delta routing to the owning shard,
the single all-gather of per-shard (key, distance) top-k and the k-way merge.  The innermost index is
the CPU oracle here; on the GPU box the same class wraps the B200 index (bench.py --gpus N uses
`allgather_merge_topk` over NCCL)."""

import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, metric):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle_adapter import OracleInner
        from pathway_b200.sharded import ShardedKnnIndex, allgather_merge_topk

        rng = np.random.default_rng(11)  # same stream on every rank: every rank sees every delta
        d = 12
        sharded = ShardedKnnIndex(OracleInner(d, 4, 1024 * 128, metric))
        single = OracleInner(d, 4, 1024 * 128, metric)
        live = []
        for step in range(6):
            rows = rng.standard_normal((40, d))
            batch = [(int(rng.integers(1, 2 ** 62)) * 2 + (i % 2), rows[i]) for i in range(40)]
            rm = [live.pop(int(rng.integers(0, len(live)))) for _ in range(min(len(live), 15))]
            sharded.remove(rm), single.remove(rm)
            sharded.add(batch), single.add(batch)
            live += [k for k, _ in batch]
            assert sharded.global_size() == single.current_size == len(live)
            qs = [(i, rng.standard_normal(d), int(rng.integers(0, 9))) for i in range(7)]
            got, want = sharded.search(qs), single.search(qs)
            for (k1, r1), (k2, r2) in zip(got, want):
                assert k1 == k2
                assert [m for m, _ in r1] == [m for m, _ in r2], (rank, step, r1, r2)
                np.testing.assert_allclose([s for _, s in r1], [s for _, s in r2], rtol=1e-12)
        # both shards hold rows (owner = key mod world)
        assert 0 < sharded.local_size < len(live)
        # raw merge helper: ties keep (rank, position) order, NaN last, +inf padding dropped last
        pay = torch.tensor([[[10 * rank + 0], [10 * rank + 1], [10 * rank + 2]]], dtype=torch.int64)
        dd = torch.tensor([[0.5, 1.0 + rank, float("inf")]], dtype=torch.float64)
        mp_, md_ = allgather_merge_topk(pay, dd, 4)
        assert mp_[0, :, 0].tolist() == [0, 10, 1, 11] and md_[0].tolist() == [0.5, 0.5, 1.0, 2.0]
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("metric", [0, 1])
def test_sharded_world2_gloo(metric):
    mp.spawn(_worker, args=(2, _free_port(), metric), nprocs=2, join=True)
