"""Option / batch-size sweep on one GPU (10 M x 768 by default): builds the index once, then times
search_device under different library options with CUDA events on the index stream."""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pathway_b200.engine import BruteForceKnnMetricKind  # noqa: E402
from pathway_b200.external_index import B200BruteForceKNNIndex  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=10_000_000)
ap.add_argument("--dims", type=int, default=768)
ap.add_argument("--k", type=int, default=10)
ap.add_argument("--reps", type=int, default=15)
ap.add_argument("--what", default="q")
ap.add_argument("--ab", default="", help="option=v1,v2,...: interleaved A/B of one library option at --q")
ap.add_argument("--q", type=int, default=1024)
ap.add_argument("--rounds", type=int, default=6)
ap.add_argument("--set", default="", help="name=value[,name=value]: library options fixed for the whole run")
args = ap.parse_args()
dev = torch.device("cuda", 0)
N, D, K = args.rows, args.dims, args.k
ix = B200BruteForceKNNIndex(D, N, metric=BruteForceKnnMetricKind.COS, device=0)
CH = 250_000
for c in range((N + CH - 1) // CH):
    m = min(CH, N - c * CH)
    g = torch.Generator(device=dev); g.manual_seed(1234 + c)
    rows = torch.randn((m, D), generator=g, device=dev, dtype=torch.float32).to(torch.bfloat16)
    keys = np.zeros((m, 2), dtype=np.uint64); keys[:, 0] = np.arange(c * CH, c * CH + m, dtype=np.uint64)
    torch.cuda.synchronize(dev)
    ix.add_device(keys, rows)
    del rows
ext = torch.cuda.ExternalStream(ix.stream_ptr, device=dev)
for kv in filter(None, args.set.split(",")):
    ix.set_option(kv.split("=")[0], int(kv.split("=")[1]))


def time_q(Q, reps=args.reps):
    g = torch.Generator(device=dev); g.manual_seed(4321)
    q = torch.randn((Q, D), generator=g, device=dev, dtype=torch.float32).to(torch.bfloat16)
    os_ = torch.empty((Q, K), dtype=torch.int32, device=dev); od = torch.empty((Q, K), dtype=torch.float64, device=dev)
    with torch.cuda.stream(ext):
        for _ in range(3):
            ix.search_device(q, K, os_, od)
        ext.synchronize()
        ix.set_option("profile", 1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ext)
        for _ in range(reps):
            ix.search_device(q, K, os_, od)
        e1.record(ext)
        ext.synchronize()
    ms = e0.elapsed_time(e1) / reps
    fused = ix.stat("fused_ns_sum") / max(1, ix.stat("fused_samples")) * 1e-6
    ix.set_option("profile", 0)
    flops = 2.0 * Q * N * D
    byts = N * D * 2 + N * 8
    return dict(Q=Q, ms=round(ms, 4), fused_ms=round(fused, 4), qps=round(Q / ms * 1e3, 1),
                tflops=round(flops / fused / 1e9, 1), hbm_gbs=round(byts / fused / 1e6, 1),
                slabs=ix.stat("last_slabs"), grid=ix.stat("last_grid"), kprime=ix.stat("last_kprime"),
                checksum=int(os_.to(torch.int64).sum().item()))


what = args.what.split(",")
if args.ab:
    name, vals = args.ab.split("=")
    vals = [int(v) for v in vals.split(",")]
    res = {v: [] for v in vals}
    for r in range(args.rounds):  # interleave so that thermal / power drift hits every setting equally
        for v in (vals if r % 2 == 0 else vals[::-1]):
            ix.set_option(name, v)
            r_ = time_q(args.q, reps=args.reps)
            res[v].append(r_["fused_ms"])
            print("AB-sample", name, v, "step_ms", r_["ms"], "fused_ms", r_["fused_ms"], "tail_us", round((r_["ms"] - r_["fused_ms"]) * 1e3, 1), flush=True)
    for v in vals:
        xs = sorted(res[v])
        print("AB", name, v, "fused_ms median", xs[len(xs) // 2], "min", xs[0], "max", xs[-1], "all", xs, flush=True)
    what = []
if "lead" in what:
    for lt in (0, 2, 4, 8, 16, 32, 0, 8):
        ix.set_option("lead_tiles", lt)
        print("lead_tiles", lt, json.dumps(time_q(1024, reps=25)), flush=True)
    ix.set_option("lead_tiles", 8)
if "diag" in what:
    for cp, stages in ((1, 2), (1, 3), (1, 4), (2, 3), (2, 4), (2, 5), (2, 6)):
        ix.set_option("cta_pair", cp); ix.set_option("max_stages", stages)
        for dbg in (0, 1):
            ix.set_option("debug_flags", dbg)
            print("diag cta_pair", cp, "stages", stages, "no_epilogue", dbg, json.dumps(time_q(1024, reps=8)), flush=True)
    ix.set_option("cta_pair", 0); ix.set_option("max_stages", 0); ix.set_option("debug_flags", 0)
if "slabs" in what:
    for s in (18, 37, 74, 148):
        ix.set_option("slabs", s)
        print("slabs", s, json.dumps(time_q(1024)), flush=True)
    ix.set_option("slabs", 0)
if "q" in what:
    for cp in (1, 2):
        ix.set_option("cta_pair", cp)
        for Q in (1, 64, 128, 129, 256, 384, 512, 1024, 2048, 4096):
            print("q-sweep cta_pair", cp, json.dumps(time_q(Q, reps=max(5, args.reps // 2))), flush=True)
ix.close()
