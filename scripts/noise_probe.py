"""Measures the noise of the tensor-core ranking pass: fp32 scores of the fused kernel (debug dump) against the
f64 ranking score of (a) the stored bf16 values (fp32-accumulation noise only: the default store) and (b) raw f32
inputs (adds the bf16 rounding noise: the exact re-rank store).  The certainty test of the search
(pwknn_host.cpp: uncertainty escalation) uses bounds derived from these numbers."""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pathway_b200.engine import BruteForceKnnMetricKind  # noqa: E402
from pathway_b200.external_index import B200BruteForceKNNIndex, keys_to_array  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=200_000)
ap.add_argument("--nq", type=int, default=128)
args = ap.parse_args()
dev = torch.device("cuda", 0)
for D in (64, 768, 1536):
    for metric in (BruteForceKnnMetricKind.COS, BruteForceKnnMetricKind.L2SQ):
        for dist_kind in ("gauss", "sparse"):
            g = torch.Generator(device=dev); g.manual_seed(D)
            raw = torch.randn((args.rows, D), generator=g, device=dev, dtype=torch.float32)
            q_raw = torch.randn((args.nq, D), generator=g, device=dev, dtype=torch.float32)
            if dist_kind == "sparse":  # heavy-tailed components: a few coordinates carry most of the norm
                raw = raw * torch.exp(2.0 * torch.randn((args.rows, D), generator=g, device=dev))
                q_raw = q_raw * torch.exp(2.0 * torch.randn((args.nq, D), generator=g, device=dev))
            xb, qb = raw.to(torch.bfloat16), q_raw.to(torch.bfloat16)
            ix = B200BruteForceKNNIndex(D, args.rows, metric=metric, device=0)
            torch.cuda.synchronize(dev)
            ix.add_device(keys_to_array(range(args.rows)), xb)
            out = torch.full((args.nq, args.rows), float("nan"), dtype=torch.float32, device=dev)
            torch.cuda.synchronize(dev)
            ix.debug_scores_device(qb, out)
            torch.cuda.synchronize(dev)

            def true_scores(x, q):  # f64 ranking score: COS dot/|x|, L2SQ dot - |x|^2/2
                x64, q64 = x.double(), q.double()
                dot = q64 @ x64.T
                nx = x64.pow(2).sum(1)
                return (dot / nx.sqrt()[None, :]) if metric is BruteForceKnnMetricKind.COS else (dot - 0.5 * nx[None, :]), nx

            res = {}
            for name, (x, q) in {"stored_bf16": (xb, qb), "raw_f32": (raw, q_raw)}.items():
                t, nx = true_scores(x, q)
                qn = q.double().pow(2).sum(1).sqrt()[:, None]
                if metric is BruteForceKnnMetricKind.COS:
                    scale = qn  # score magnitude <= |q|
                else:
                    scale = qn * nx.sqrt()[None, :] + 0.5 * nx[None, :]
                e = (out.double() - t).abs() / scale
                res[name] = dict(max=float(e.max()), rms=float(e.pow(2).mean().sqrt()),
                                 p9999=float(torch.quantile(e.flatten()[:4_000_000].float(), 0.9999)))
            print(json.dumps(dict(D=D, metric=metric.name, data=dist_kind, **res)), flush=True)
            ix.close()
            del raw, q_raw, xb, qb, out
