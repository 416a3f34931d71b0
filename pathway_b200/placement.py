"""Placement weights for a multi-device handle (pwknn_multi_set_weights).

A search on a multi-device handle costs its slowest device, and under the per-GPU power cap the GPUs of one box run
the same tensor-bound kernel at persistently different speeds (measured: up to 35 % apart).  `calibrate_device_weights`
loads every device with the index's own search kernel on a scratch corpus for about a second -- all devices at once,
as in production -- and returns weights proportional to (rows per second) ** 0.6; the exponent damps the correction
because a device that is given more rows also draws more power and slows a little (measured on 2 and 8 GPUs: the full
correction overshoots by about half)."""

from __future__ import annotations

import threading
import time

import numpy as np

from .engine import BruteForceKnnMetricKind
from .external_index import B200BruteForceKNNIndex, keys_to_array


def calibrate_device_weights(devices: list[int], dimensions: int, *, rows: int = 400_000, queries: int = 1024,
                             k: int = 10, seconds: float = 1.2, damping: float = 0.6) -> list[float]:
    """-> one positive weight per entry of `devices` (mean 1.0).  Needs torch (device RNG and streams)."""
    import torch

    if len(set(devices)) <= 1:
        return [1.0] * len(devices)
    uniq = sorted(set(devices))
    state = {}
    for d in uniq:
        dev = torch.device("cuda", d)
        ix = B200BruteForceKNNIndex(dimensions, rows, metric=BruteForceKnnMetricKind.COS, device=d)
        g = torch.Generator(device=dev)
        g.manual_seed(99 + d)
        chunk = 100_000
        for r0 in range(0, rows, chunk):
            n = min(chunk, rows - r0)
            x = torch.randn((n, dimensions), generator=g, device=dev, dtype=torch.float32).to(torch.bfloat16)
            torch.cuda.synchronize(dev)
            ix.add_device(keys_to_array(range(r0, r0 + n)), x)
            del x
        q = torch.randn((queries, dimensions), generator=g, device=dev, dtype=torch.float32).to(torch.bfloat16)
        out_s = torch.empty((queries, k), dtype=torch.int32, device=dev)
        out_d = torch.empty((queries, k), dtype=torch.float64, device=dev)
        torch.cuda.synchronize(dev)
        state[d] = (ix, q, out_s, out_d)

    rate = {}

    def load(d):
        ix, q, out_s, out_d = state[d]
        ext = torch.cuda.ExternalStream(ix.stream_ptr, device=torch.device("cuda", d))

        def burst(t_len):
            t_end = time.perf_counter() + t_len
            while time.perf_counter() < t_end:
                for _ in range(8):
                    ix.search_device(q, k, out_s, out_d)
                ext.synchronize()

        burst(seconds / 2)           # let the power-cap controller settle
        ix.set_option("profile", 1)  # second half is measured: device time of the fused kernel
        burst(seconds / 2)
        n_samples = ix.stat("fused_samples")
        rate[d] = rows / (ix.stat("fused_ns_sum") / n_samples) if n_samples > 0 and ix.stat("fused_ns_sum") > 0 else 0.0
        ix.set_option("profile", 0)

    try:
        threads = [threading.Thread(target=load, args=(d,)) for d in uniq]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
    finally:
        for ix, *_ in state.values():
            ix.close()
    return clamp_weights([rate.get(d, 0.0) for d in devices], damping)


def clamp_weights(rates: list[float], damping: float = 0.6, band: float = 0.25) -> list[float]:
    """rates (rows per second per device; <= 0 or missing = unknown) -> weights with mean 1.0, every one inside
    [1 - band, 1 + band] of the mean: a calibration hiccup must not unbalance the index for its lifetime.  Clipping and
    renormalising are iterated to a fixed point (one clip + one renormalisation can leave the band again)."""
    r = np.array(rates, dtype=np.float64)
    known = np.isfinite(r) & (r > 0)
    if not known.any():
        return [1.0] * len(rates)
    r = np.where(known, r, r[known].mean())  # a device without a sample gets the mean speed
    w = r ** damping
    w = w / w.mean()
    for _ in range(32):
        c = np.clip(w, 1.0 - band, 1.0 + band)
        c = c / c.mean()
        if np.allclose(c, w, rtol=0, atol=1e-12):
            break
        w = c
    return np.clip(w, 1.0 - band, 1.0 + band).tolist()
