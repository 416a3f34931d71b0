import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pathway_b200.engine import BruteForceKnnMetricKind
from pathway_b200.external_index import B200BruteForceKNNIndex
import torch
dev = torch.device("cuda", 0)
N, D, Q, K = 1_250_000, 768, 1024, 10
ix = B200BruteForceKNNIndex(D, N, metric=BruteForceKnnMetricKind.COS, device=0)
for c in range(5):
    g = torch.Generator(device=dev); g.manual_seed(c)
    rows = torch.randn((250_000, D), generator=g, device=dev).to(torch.bfloat16)
    keys = np.zeros((250_000, 2), dtype=np.uint64); keys[:, 0] = np.arange(c * 250_000, (c + 1) * 250_000, dtype=np.uint64)
    torch.cuda.synchronize(); ix.add_device(keys, rows)
q = np.random.default_rng(0).standard_normal((Q, D))
lim = np.full(Q, K, dtype=np.uint32)
for _ in range(5): ix.search_arrays(q, lim)
ts, st, sh, en, sy = [], [], [], [], []
ix.set_option("profile", 1)
for _ in range(30):
    t0 = time.perf_counter(); ix.search_arrays(q, lim); ts.append(time.perf_counter() - t0)
    st.append(ix.stat("stage_ns")); sh.append(ix.stat("search_host_ns")); en.append(ix.stat("enqueue_ns")); sy.append(ix.stat("sync_ns"))
print("fused ms", ix.stat("fused_ns_sum") / ix.stat("fused_samples") / 1e6, "enqueue-done ms", np.median(en) / 1e6, "sync-done ms", np.median(sy) / 1e6)
print("PWKNN_HOST_THREADS", os.environ.get("PWKNN_HOST_THREADS"), "python call ms", round(np.median(ts) * 1e3, 3),
      "C call ms", round(np.median(sh) / 1e6, 3), "stage ms", round(np.median(st) / 1e6, 3), "nproc", os.cpu_count(),
      "affinity", len(os.sched_getaffinity(0)))
