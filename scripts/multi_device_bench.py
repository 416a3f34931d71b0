"""One process, one handle, G GPUs (pwknn_create_multi): the drop-in case of a single Pathway worker on a multi-GPU
box.  Times pwknn_search (HOST f64 buffers in, host keys/scores out -- the call the Rust shim makes) on BASELINE's C2
workload split over the devices, next to the same handle on one GPU.

  python scripts/multi_device_bench.py --devices 0,1 [--rows 10000000] [--steps 20]

Prints one JSON line per device count.  The corpus is generated on each device and bulk-loaded through the shard
handles (pwknn_shard + pwknn_add_device_bf16); keys are picked so that the router owns them on that shard."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pathway_b200.engine import BruteForceKnnMetricKind  # noqa: E402
from pathway_b200.external_index import B200BruteForceKNNIndex, keys_to_array  # noqa: E402


def load(ix, devices, rows_total, dims, chunk=250_000):
    g_count = len(devices)
    next_key = 0
    filled = [0] * g_count
    # fill with consecutive keys until the total is reached: every shard gets what the router gives it
    while sum(filled) < rows_total:
        cand = keys_to_array(range(next_key, next_key + chunk * g_count))
        next_key += chunk * g_count
        owner = ix.shards_of_keys(cand)
        for g in range(g_count):
            mine = cand[owner == g][: max(0, rows_total - sum(filled))]
            if not len(mine):
                continue
            sh = ix.shard(g)
            gen = torch.Generator(device=f"cuda:{sh.device}")
            gen.manual_seed(1234 + 7919 * g + filled[g])
            t = torch.randn((len(mine), dims), generator=gen, device=f"cuda:{sh.device}", dtype=torch.float32)
            t = t.to(torch.bfloat16)
            torch.cuda.synchronize(sh.device)
            st = sh.add_device(mine, t)
            assert (st == 0).all()
            filled[g] += len(mine)
            del t
    return sum(filled)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--devices", default="0")
    ap.add_argument("--rows", type=int, default=10_000_000)
    ap.add_argument("--dims", type=int, default=768)
    ap.add_argument("--queries", type=int, default=1024)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--calibrate", action="store_true", help="weight the placement by measured per-GPU speed")
    a = ap.parse_args()
    devices = [int(x) for x in a.devices.split(",")]
    weights = None
    if a.calibrate:
        from pathway_b200.placement import calibrate_device_weights
        weights = calibrate_device_weights(devices, a.dims)
    ix = B200BruteForceKNNIndex(a.dims, a.rows, metric=BruteForceKnnMetricKind.COS, devices=devices, weights=weights)
    t0 = time.perf_counter()
    n = load(ix, devices, a.rows, a.dims)
    load_s = time.perf_counter() - t0
    assert ix.current_size == n
    rng = np.random.default_rng(4321)
    batches = [rng.standard_normal((a.queries, a.dims)) for _ in range(4)]  # f64 host buffers, like the trait
    limits = np.full(a.queries, a.k, dtype=np.uint32)
    times = []
    for it in range(a.warmup + a.steps):
        t0 = time.perf_counter()
        keys, scores, counts, status = ix.search_arrays(batches[it % 4], limits)
        dt = time.perf_counter() - t0
        if it >= a.warmup:
            times.append(dt)
    assert (counts == a.k).all() and (status == 0).all()
    ms = float(np.median(times)) * 1e3
    detail = {}
    if len(devices) > 1:  # each shard alone (the other devices idle), then the fused kernel's device time per shard
        ix.set_option("profile", 1)
        for it in range(8):
            ix.search_arrays(batches[it % 4], limits)
        for g in range(len(devices)):
            sh = ix.shard(g)
            detail[f"shard{g}_fused_ms_concurrent"] = round(sh.stat("fused_ns_sum") / max(1, sh.stat("fused_samples")) / 1e6, 3)
        ix.set_option("profile", 0)
        for g in range(len(devices)):
            sh = ix.shard(g)
            ts = []
            for it in range(8):
                t0 = time.perf_counter()
                sh.search_arrays(batches[it % 4], limits)
                ts.append(time.perf_counter() - t0)
            detail[f"shard{g}_alone_ms"] = round(float(np.median(ts[2:])) * 1e3, 3)
    print(json.dumps({
        "mode": "one handle, %d device(s)" % len(devices), "devices": devices, "rows": n, "dims": a.dims,
        "queries": a.queries, "k": a.k, "ms_per_batch_median": round(ms, 3), "ms_min": round(min(times) * 1e3, 3),
        "qps_e2e_host_buffers": round(a.queries / ms * 1e3), "load_s": round(load_s, 1),
        "stage_ms": ix.stat("stage_ns") / 1e6, "enqueue_ms": ix.stat("enqueue_ns") / 1e6,
        "sync_ms": ix.stat("sync_ns") / 1e6, "call_ms": ix.stat("search_host_ns") / 1e6,
        "h2d_bytes": ix.stat("h2d_bytes"), "weights": weights,
        "shard_rows": [ix.shard(g).current_size for g in range(len(devices))], **detail}))
    ix.close()


if __name__ == "__main__":
    main()
