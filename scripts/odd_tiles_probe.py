"""Planner probe for odd query-tile counts (Q = 384, 640, 896, 1152): slabs / lead_tiles / cta_pair combinations on
one 10 M x 768 index; fused-kernel ms from CUDA events (library profile option)."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pathway_b200.engine import BruteForceKnnMetricKind  # noqa: E402
from pathway_b200.external_index import B200BruteForceKNNIndex  # noqa: E402

dev = torch.device("cuda", 0)
N, D, K = 10_000_000, 768, 10
ix = B200BruteForceKNNIndex(D, N, metric=BruteForceKnnMetricKind.COS, device=0)
CH = 250_000
for c in range(N // CH):
    g = torch.Generator(device=dev); g.manual_seed(1234 + c)
    rows = torch.randn((CH, D), generator=g, device=dev, dtype=torch.float32).to(torch.bfloat16)
    keys = np.zeros((CH, 2), dtype=np.uint64); keys[:, 0] = np.arange(c * CH, (c + 1) * CH, dtype=np.uint64)
    torch.cuda.synchronize(dev)
    ix.add_device(keys, rows)
    del rows
ext = torch.cuda.ExternalStream(ix.stream_ptr, device=dev)


def run(Q, reps=6):
    g = torch.Generator(device=dev); g.manual_seed(4321)
    q = torch.randn((Q, D), generator=g, device=dev, dtype=torch.float32).to(torch.bfloat16)
    os_ = torch.empty((Q, K), dtype=torch.int32, device=dev); od = torch.empty((Q, K), dtype=torch.float64, device=dev)
    with torch.cuda.stream(ext):
        for _ in range(2):
            ix.search_device(q, K, os_, od)
        ext.synchronize()
        ix.set_option("profile", 1)
        for _ in range(reps):
            ix.search_device(q, K, os_, od)
        ext.synchronize()
    fused = ix.stat("fused_ns_sum") / max(1, ix.stat("fused_samples")) * 1e-6
    ix.set_option("profile", 0)
    return round(fused, 3), int(os_.to(torch.int64).sum().item()), ix.stat("last_slabs"), ix.stat("last_grid"), ix.stat("last_ctas")


cases = {
    384: [(0, 0, 4), (1, 49, 4), (1, 49, 2), (1, 49, 1), (1, 98, 2), (2, 37, 4), (1, 148, 1)],
    640: [(0, 0, 4), (1, 29, 4), (1, 29, 2), (1, 148, 2), (2, 0, 4), (2, 24, 4)],
    896: [(0, 0, 4), (1, 21, 4), (1, 21, 2), (1, 42, 4), (2, 0, 4), (2, 18, 4)],
    1152: [(0, 0, 4), (1, 16, 4), (1, 32, 4), (2, 0, 4), (2, 14, 4), (2, 29, 4)],
    512: [(0, 0, 4), (1, 37, 4), (2, 37, 4)],
}
for rnd in range(2):
    for Q, cfgs in cases.items():
        for cp, slabs, lead in cfgs:
            ix.set_option("cta_pair", cp); ix.set_option("slabs", slabs); ix.set_option("lead_tiles", lead)
            ms, chk, s, grid, ctas = run(Q)
            print(json.dumps({"Q": Q, "cta_pair": cp, "slabs_opt": slabs, "lead": lead, "fused_ms": ms, "slabs": s,
                              "grid": grid, "ctas": ctas, "tflops": round(2.0 * Q * N * D / ms / 1e9, 1),
                              "checksum": chk}), flush=True)
ix.close()
