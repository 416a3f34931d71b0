"""Per-CTA run times of the fused kernel (diagnostics): are there straggler CTA pairs?  10 M x 768, Q = 1024."""
import os, sys, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from pathway_b200.engine import BruteForceKnnMetricKind
from pathway_b200.external_index import B200BruteForceKNNIndex
dev = torch.device("cuda", 0); N, D, K = int(os.environ.get("ROWS", 10_000_000)), 768, 10
ix = B200BruteForceKNNIndex(D, N, metric=BruteForceKnnMetricKind.COS, device=0)
CH = 250_000
for c in range(N // CH):
    g = torch.Generator(device=dev); g.manual_seed(1234 + c)
    rows = torch.randn((CH, D), generator=g, device=dev, dtype=torch.float32).to(torch.bfloat16)
    keys = np.zeros((CH, 2), dtype=np.uint64); keys[:, 0] = np.arange(c * CH, (c + 1) * CH, dtype=np.uint64)
    torch.cuda.synchronize(dev); ix.add_device(keys, rows); del rows
for Q in (1024, 512, 256):
    g = torch.Generator(device=dev); g.manual_seed(4321)
    q = torch.randn((Q, D), generator=g, device=dev, dtype=torch.float32).to(torch.bfloat16)
    os_ = torch.empty((Q, K), dtype=torch.int32, device=dev); od = torch.empty((Q, K), dtype=torch.float64, device=dev)
    for lead in (4, 0):
        ix.set_option("lead_tiles", lead)
        ix.set_option("cta_times", 0)
        for _ in range(20): ix.search_device(q, K, os_, od, stream_ptr=ix.stream_ptr)
        torch.cuda.synchronize(dev)
        ix.set_option("cta_times", 1)
        for rep in range(3):
            ix.search_device(q, K, os_, od, stream_ptr=ix.stream_ptr)
            torch.cuda.synchronize(dev)
            grid = ix.stat("last_grid")
            dur = np.array([ix.stat(f"cta_time_{i}") for i in range(grid)]) * 1e-6
            st = np.array([ix.stat(f"cta_start_{i}") for i in range(grid)]) * 1e-6
            end = st + dur
            order = np.argsort(end)
            print(json.dumps(dict(Q=Q, lead=lead, rep=rep, grid=int(grid), slabs=ix.stat("last_slabs"), kernel_ms=round(float(end.max() - st.min()), 3),
                                  end_ms_pct=[round(float(np.percentile(end, p)), 3) for p in (0, 10, 50, 90, 99, 100)],
                                  start_ms_max=round(float(st.max()), 3), slowest_ctas=[int(x) for x in order[-10:]],
                                  slowest_end=[round(float(end[x]), 3) for x in order[-10:]])), flush=True)
ix.close()
