"""The two CPU restatements (numpy, C) agree with each other on random delta / query sequences:
sizes, allocation rule, storage layout, ids and scores.  The numpy one is the readable spec, the C one
is what bench.py times; both are pinned to the reference's KATs in test_oracle_golden.py."""

import ctypes as C

import numpy as np
import pytest

from oracle import brute_force_oracle as bfo
from oracle.c_oracle import CBruteForceKNNIndex, lib


def _c_rows(ix, d):
    n = ix.current_size
    ptr = lib().pwo_rows(ix._h)
    return np.ctypeslib.as_array((C.c_double * (n * d)).from_address(ptr)).reshape(n, d).copy()


@pytest.mark.parametrize("metric", [bfo.L2SQ, bfo.COS])
@pytest.mark.parametrize("reserved,aux", [(0, 1), (3, 50), (64, 100000)])
def test_random_sequences(metric, reserved, aux):
    rng = np.random.default_rng(reserved * 31 + metric)
    d = 6
    a = bfo.BruteForceKNNIndex(d, reserved, aux, metric)
    b = CBruteForceKNNIndex(d, reserved, aux, metric)
    live, nxt = [], 1
    for step in range(25):
        n_rm = int(rng.integers(0, min(len(live), 30) + 1))
        rm = [live.pop(int(rng.integers(0, len(live)))) for _ in range(n_rm)]
        if step % 4 == 1 and live:
            rm.append(10 ** 12 + step)  # MissingKey: state unchanged
        ra, rb = a.remove(rm), b.remove(rm)
        assert [(k, e is None) for k, e in ra] == [(k, e is None) for k, e in rb]
        n_add = int(rng.integers(0, 40))
        rows = rng.standard_normal((n_add, d))
        if n_add:
            rows[0] = 0.0  # a zero vector: NaN cosine distance, ranked last
        batch = [((nxt + i) << 70 | (nxt + i), rows[i]) for i in range(n_add)]  # exercise the hi half of the key
        nxt += n_add
        a.add(batch), b.add(batch)
        live += [k for k, _ in batch]
        assert a.current_size == b.current_size == len(live)
        assert a.current_allocated == b.current_allocated
        if live:
            np.testing.assert_array_equal(a.index_array[: a.current_size], _c_rows(b, d))
        qs = [(i, rng.standard_normal(d), int(rng.integers(0, 12))) for i in range(6)]
        for (k1, r1), (k2, r2) in zip(a.search(qs), b.search(qs, workers=3)):
            assert k1 == k2 and [m for m, _ in r1] == [m for m, _ in r2]
            np.testing.assert_allclose([s for _, s in r1], [s for _, s in r2], rtol=1e-12, atol=1e-15, equal_nan=True)


def test_bf16_round_is_rne_and_idempotent():
    x = np.array([0.1, 0.6, 1.0, -3.14159, 1e-8, 65504.0, 0.0, 1.00390625, 1.01171875])
    r = bfo.bf16_round(x)
    np.testing.assert_array_equal(bfo.bf16_round(r), r)
    assert r[0] == 0.10009765625 and r[1] == 0.6015625
    assert r[7] == 1.0 and r[8] == 1.015625  # ties to even: 1 + 1/256 -> 1, 1 + 3/256 -> 1 + 4/256
    import torch

    t = torch.tensor(x, dtype=torch.float64).to(torch.float32).to(torch.bfloat16).to(torch.float64).numpy()
    np.testing.assert_array_equal(r, t)


def test_config1_shape_static_table_k3():
    """BASELINE.json configs[0] shape (1 k x 64 random embeddings, k = 3, static table, one worker, no GPU):
    the two CPU restatements agree row for row; this is the plumbing case the reference runs on its CPU engine."""
    rng = np.random.default_rng(0)
    rows = rng.standard_normal((1000, 64))
    qs = rng.standard_normal((100, 64))
    for metric in (bfo.COS, bfo.L2SQ):
        a = bfo.BruteForceKNNIndex(64, 400, 1024 * 128, metric)
        b = CBruteForceKNNIndex(64, 400, 1024 * 128, metric)
        data = [(i + 1, rows[i]) for i in range(1000)]
        a.add(data), b.add(data)
        q = [(i, qs[i], 3) for i in range(100)]
        for (_, r1), (_, r2) in zip(a.search(q), b.search(q, workers=2)):
            assert [m for m, _ in r1] == [m for m, _ in r2] and len(r1) == 3
            np.testing.assert_allclose([x for _, x in r1], [x for _, x in r2], rtol=1e-12)
