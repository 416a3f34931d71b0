"""Row-sharded KNN index over `torch.distributed` (one process per GPU).

The reference replicates the whole index on every timely worker: the index stream is
`broadcast()` to all workers and queries are partitioned (/root/reference
src/engine/dataflow/operators/external_index.rs:90-101).  Here every rank still RECEIVES every
delta (same broadcast contract), but keeps only the rows it owns, every rank sees every query, and a
query batch ends with ONE all-gather of the per-shard (key, distance) top-k followed by a k-way merge
(SURVEY.md section 8e).  The merge orders by (distance, rank, position-in-shard): inside a shard that is
the reference's (distance, slot) order; across shards exact distance ties fall back to the rank order
(documented deviation: the reference's slot order is history-dependent and not shardable).
"""

from __future__ import annotations

from typing import Any, Callable, Optional

import numpy as np
import torch
import torch.distributed as dist


def allgather_merge_topk(payload: torch.Tensor, distance: torch.Tensor, k_out: int, group=None):
    """payload [Q,K,C] int64, distance [Q,K] float64 (ascending per row, +inf / NaN padded) on every rank
    -> the global k_out best per query, identical on every rank: (payload [Q,k_out,C], distance [Q,k_out]).

    One collective: the distance bits travel as an extra int64 column of the payload."""
    world = dist.get_world_size(group)
    q, k, c = payload.shape
    packed = torch.cat([payload, distance.contiguous().view(torch.int64).unsqueeze(-1)], dim=-1).contiguous()
    if dist.get_backend(group) == "nccl":
        gathered = torch.empty((world, q, k, c + 1), dtype=torch.int64, device=packed.device)
        dist.all_gather_into_tensor(gathered, packed, group=group)
    else:  # gloo (CPU tests)
        parts = [torch.empty_like(packed) for _ in range(world)]
        dist.all_gather(parts, packed, group=group)
        gathered = torch.stack(parts, 0)
    d_all = gathered[..., c].contiguous().view(torch.float64).permute(1, 0, 2).reshape(q, world * k)
    p_all = gathered[..., :c].permute(1, 0, 2, 3).reshape(q, world * k, c)
    # stable ascending sort: NaN last (ordered-float), ties keep (rank, position) order
    d_sorted, order = torch.sort(d_all, dim=1, stable=True)
    k_out = min(k_out, world * k)
    idx = order[:, :k_out]
    return torch.gather(p_all, 1, idx.unsqueeze(-1).expand(q, k_out, c)), d_sorted[:, :k_out]


def default_owner(key: int, world: int) -> int:
    return int(key) % world


class ShardedKnnIndex:
    """NonFilteringExternalIndex-shaped wrapper: add / remove / search with the trait's list shapes
    (see pathway_b200.external_index.B200BruteForceKNNIndex), sharded over the process group."""

    def __init__(self, inner, group=None, device: Optional[torch.device] = None,
                 owner: Callable[[int, int], int] = default_owner):
        self.inner = inner
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.device = device if device is not None else torch.device("cpu")
        self.owner = owner

    @property
    def local_size(self) -> int:
        return self.inner.current_size

    def global_size(self) -> int:
        t = torch.tensor([self.inner.current_size], dtype=torch.int64, device=self.device)
        dist.all_reduce(t, group=self.group)
        return int(t.item())

    def add(self, add_data: list[tuple[int, Any]]):
        mine = [(k, v) for k, v in add_data if self.owner(k, self.world) == self.rank]
        res = dict(self.inner.add(mine)) if mine else {}
        return [(k, res.get(k)) for k, _ in add_data]

    def remove(self, keys: list[int]):
        mine = [k for k in keys if self.owner(k, self.world) == self.rank]
        res = dict(self.inner.remove(mine)) if mine else {}
        return [(k, res.get(k)) for k in keys]

    def search(self, queries: list[tuple[int, Any, int]]):
        """Every rank passes the same queries; every rank gets the same merged answer."""
        if not queries:
            return []
        local = self.inner.search(queries)
        kmax = max(1, max(lim for _, _, lim in queries))
        nq = len(queries)
        payload = np.full((nq, kmax, 2), -1, dtype=np.int64)
        distance = np.full((nq, kmax), np.inf, dtype=np.float64)
        errors = {}
        for qi, (key, res) in enumerate(local):
            if isinstance(res, Exception):
                errors[qi] = res
                continue
            for r, (mk, score) in enumerate(res[:kmax]):
                payload[qi, r, 0] = np.int64(np.uint64(mk & 0xFFFFFFFFFFFFFFFF))
                payload[qi, r, 1] = np.int64(np.uint64((mk >> 64) & 0xFFFFFFFFFFFFFFFF))
                distance[qi, r] = -score
        p_t = torch.from_numpy(payload).to(self.device)
        d_t = torch.from_numpy(distance).to(self.device)
        valid = torch.from_numpy(np.array([[r < len(res) if not isinstance(res, Exception) else False
                                            for r in range(kmax)] for _, res in local])).to(self.device)
        # keep "valid" next to the payload so NaN distances (zero vectors) are not confused with padding
        p_t = torch.cat([p_t, valid.to(torch.int64).unsqueeze(-1)], dim=-1)
        # padding must sort after real NaN entries: stable sort keeps them behind only inside one rank,
        # so order by (is_padding, distance)
        mp, md = allgather_merge_topk(p_t, torch.where(valid, d_t, torch.full_like(d_t, float("inf"))), kmax * self.world,
                                      self.group)
        mp, md = mp.cpu().numpy(), md.cpu().numpy()
        out = []
        for qi, (key, _vec, lim) in enumerate(queries):
            if qi in errors:
                out.append((key, errors[qi]))
                continue
            rows = [(int(np.uint64(mp[qi, r, 0])) | (int(np.uint64(mp[qi, r, 1])) << 64), -float(md[qi, r]))
                    for r in range(mp.shape[1]) if mp[qi, r, 2] == 1]
            # real matches with NaN distance were sorted behind +inf padding of other ranks: they are still
            # in `rows` (valid flag) and keep their relative order; finite ones come first already.
            out.append((key, rows[:lim]))
        return out
