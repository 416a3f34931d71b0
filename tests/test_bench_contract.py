"""bench.py contract (no GPU): the reference arm prints ONE JSON line with the keys the driver reads."""

import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "1",
                          "--warmup", "1", "--cpu-rows", "20000"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr
    lines = [ln for ln in out.stdout.strip().splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert d["impl"] == "reference" and d["metric"] == base["metric"] and d["unit"] == "queries/s"
    for key in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["value"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["config"]["workload"].startswith("BruteForceKnn 10000000x768")


def test_shard_ranges_cover_the_corpus():
    """N > 1 placement: contiguous ranges, exact cover, proportional to the weights, tile-aligned inner boundaries."""
    sys.path.insert(0, ROOT)
    import bench

    n = 10_000_000
    for weights in ([1.0] * 8, [1.0, 1.3], [0.9, 1.0, 1.1, 1.0], [1e-9, 1.0], [1.0]):
        r = bench.shard_ranges(n, weights)
        assert r[0][0] == 0 and r[-1][1] == n and len(r) == len(weights)
        assert all(a[1] == b[0] for a, b in zip(r, r[1:])) and all(lo <= hi for lo, hi in r)
        assert all(hi % 256 == 0 for _, hi in r[:-1])
        total = sum(weights)
        for (lo, hi), w in zip(r, weights):
            assert abs((hi - lo) - n * w / total) <= 256
    assert bench.shard_ranges(100, [1.0, 1.0, 1.0]) == [(0, 0), (0, 0), (0, 100)] or \
        sum(hi - lo for lo, hi in bench.shard_ranges(100, [1.0, 1.0, 1.0])) == 100
