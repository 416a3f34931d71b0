"""Multi-device handle (pwknn_create_multi): one NonFilteringExternalIndex whose rows are spread over several
shards.  Parity against the CPU oracle on the same seeded inputs; runs on a single-GPU box by listing device 0
several times and, when the box has more GPUs, over real distinct devices too."""

import ctypes as C

import numpy as np
import pytest

from golden import kats
from kat_harness import run_case
from oracle_adapter import OracleInner
from pathway_b200 import _native
from pathway_b200.debug import ref_scalar
from pathway_b200.engine import BruteForceKnnMetricKind
from pathway_b200.external_index import B200BruteForceKNNIndex, MissingKey
from test_gpu_parity import _check_same, _rounded

pytestmark = pytest.mark.gpu

COS, L2SQ = BruteForceKnnMetricKind.COS, BruteForceKnnMetricKind.L2SQ


def _device_sets():
    import torch

    sets = [[0, 0, 0]]
    if torch.cuda.device_count() >= 2:
        sets.append(list(range(min(4, torch.cuda.device_count()))))
    return sets


@pytest.mark.parametrize("metric", [COS, L2SQ])
def test_incremental_parity(metric):
    """adds / upserts / removes / missing keys / searches with mixed limits, checked step by step."""
    for devices in _device_sets():
        rng = np.random.default_rng(11)
        d = 96
        gpu = B200BruteForceKNNIndex(d, 16, metric=metric, devices=devices)
        ora = OracleInner(d, 16, 1024 * 128, metric, impl="c")
        assert gpu.stat("shards") == len(devices)
        live: list[int] = []
        next_key = 0
        for step in range(6):
            n_add = [700, 50, 1200, 3, 400, 64][step]
            fresh = [ref_scalar("k", next_key + i) for i in range(n_add)]
            next_key += n_add
            data = [(k, _rounded(rng, d)) for k in fresh]
            gone = [live[int(i)] for i in rng.choice(len(live), size=min(len(live), 150 * (step % 3)), replace=False)] \
                if live else []
            bogus = [ref_scalar("nope", step)] if live else []  # removing from an empty index panics in the reference
            r1, r2 = gpu.remove(gone + bogus), ora.remove(gone + bogus)
            assert [(k, type(e)) for k, e in r1] == [(k, type(e)) for k, e in r2]
            assert not bogus or isinstance(r1[-1][1], MissingKey)
            gone_set = set(gone)
            live = [k for k in live if k not in gone_set]
            # upsert = remove-then-add of a live key inside one timestamp (mod.rs:144-150)
            redo = live[:5]
            gpu.remove(redo), ora.remove(redo)
            data += [(k, _rounded(rng, d)) for k in redo]
            assert all(e is None for _, e in gpu.add(data))
            ora.add(data)
            live += fresh
            assert gpu.current_size == ora.current_size == len(live)
            qs = [(1000 * step + j, _rounded(rng, d), [1, 3, 10, 0, 40][j % 5]) for j in range(37)]
            _check_same(gpu.search(qs), ora.search(qs), (devices, metric, step))
        gpu.close()


def test_limits_beyond_shard_and_index_size():
    """limit > rows on a shard and > rows in the whole index; empty shards; empty index."""
    rng = np.random.default_rng(12)
    gpu = B200BruteForceKNNIndex(32, 4, metric=L2SQ, devices=[0, 0, 0, 0])
    ora = OracleInner(32, 4, 1024 * 128, L2SQ, impl="c")
    qs = [(j, _rounded(rng, 32), 7) for j in range(5)]
    assert gpu.search(qs) == [(j, []) for j in range(5)]  # empty index -> empty replies (rs:193-198)
    data = [(ref_scalar(i), _rounded(rng, 32)) for i in range(5)]  # fewer rows than shards * limit
    gpu.add(data), ora.add(data)
    qs = [(j, _rounded(rng, 32), [1, 5, 7, 100][j % 4]) for j in range(9)]
    _check_same(gpu.search(qs), ora.search(qs))
    gpu.close()


def test_large_k_and_exact_kernel():
    """k = 100 (optimistic two-pass lists) and the any-k exact path, through the router."""
    rng = np.random.default_rng(13)
    d, n = 64, 6000
    rows = _rounded(rng, (n, d))
    data = [(ref_scalar(i), rows[i]) for i in range(n)]
    for kernel in (0, 2):
        gpu = B200BruteForceKNNIndex(d, n, metric=COS, devices=[0, 0])
        ora = OracleInner(d, n, 1024 * 128, COS, impl="c")
        gpu.set_option("kernel", kernel)
        gpu.add(data), ora.add(data)
        qs = [(j, _rounded(rng, d), [100, 300][j % 2] if kernel == 0 else [100, 2500][j % 2]) for j in range(20)]
        _check_same(gpu.search(qs), ora.search(qs), kernel)
        gpu.close()


def test_bad_vectors_and_f32_entry_points():
    rng = np.random.default_rng(14)
    d = 48
    gpu = B200BruteForceKNNIndex(d, 8, metric=COS, devices=[0, 0, 0])
    ora = OracleInner(d, 8, 1024 * 128, COS, impl="c")
    rows = _rounded(rng, (200, d))
    keys = [ref_scalar(i) for i in range(200)]
    bad = rows[7].copy()
    bad[3] = np.inf
    res = gpu.add([(keys[i], rows[i]) for i in range(100)] + [(ref_scalar("bad"), bad)])
    assert [e is None for _, e in res] == [True] * 100 + [False]
    # f32 ABI entry points through the router
    from pathway_b200.external_index import keys_to_array
    st = gpu.add_arrays(keys_to_array(keys[100:]), rows[100:].astype(np.float32))
    assert (st == 0).all()
    ora.add([(keys[i], rows[i]) for i in range(200)])
    assert gpu.current_size == 200
    q = _rounded(rng, (6, d))
    nanq = q[2].copy()
    nanq[0] = np.nan
    qs = [(0, q[0], 4), (1, q[1], 4), (2, nanq, 4), (3, q[3], 9)]
    got = gpu.search(qs)
    assert isinstance(got[2][1], Exception)
    good = [qs[0], qs[1], qs[3]]
    _check_same([got[0], got[1], got[3]], ora.search(good))
    gpu.close()


def test_filters_route_with_their_rows():
    rng = np.random.default_rng(15)
    d, n = 40, 900
    rows = _rounded(rng, (n, d))
    keys = [ref_scalar(i) for i in range(n)]
    multi = B200BruteForceKNNIndex(d, 8, metric=L2SQ, devices=[0, 0, 0])
    single = B200BruteForceKNNIndex(d, 8, metric=L2SQ)
    for ix in (multi, single):
        ix.add([(keys[i], rows[i]) for i in range(n)])
    even, third = multi.filter_create(), multi.filter_create()
    assert (even, third) == (single.filter_create(), single.filter_create())
    for ix in (multi, single):
        ix.filter_set(even, keys, [i % 2 == 0 for i in range(n)])
        ix.filter_set(third, keys, [i % 3 == 0 for i in range(n)])
        ix.remove(keys[:100])  # swap-removes: the bits must follow their rows on every shard
    qs = [(j, _rounded(rng, d), 1 + j % 12, [even, third, None][j % 3]) for j in range(40)]
    a, b = multi.search_filtered(qs), single.search_filtered(qs)
    _check_same(a, b)
    by_key = {k: i for i, k in enumerate(keys)}
    for (_, reply), (_, _, _, f) in zip(a, qs):
        for m, _ in reply:
            assert by_key[m] >= 100 and (f is None or by_key[m] % (2 if f == even else 3) == 0)
    multi.filter_destroy(even)
    multi.close(), single.close()


def test_per_device_entry_points_are_refused():
    gpu = B200BruteForceKNNIndex(16, 8, metric=COS, devices=[0, 0])
    L = _native.lib()
    out = np.zeros(16)
    assert L.pwknn_read_rows(gpu._h, 0, 1, out.ctypes.data) == _native.E_INVALID
    assert b"multi-device" in L.pwknn_last_error(gpu._h)
    assert L.pwknn_search_device(gpu._h, 1, 1, 16, 1, 1, 1, None) == _native.E_INVALID
    h = C.c_void_p()
    assert L.pwknn_create_multi(16, 8, 0, 1, None, 0, 0, C.byref(h)) == _native.E_INVALID
    bad = (C.c_int * 1)(99)
    assert L.pwknn_create_multi(16, 8, 0, 1, bad, 1, 0, C.byref(h)) == _native.E_INVALID and not h.value
    gpu.close()


@pytest.mark.parametrize("name", sorted(kats.CASES))
def test_reference_kats_on_three_shards(name):
    def make_inner(dimensions, reserved_space, auxiliary_space, metric):
        return B200BruteForceKNNIndex(dimensions, reserved_space, auxiliary_space, metric, devices=[0, 0, 0])
    inner = run_case(name, make_inner)
    inner.close()


def test_exact_rerank_store_on_shards():
    """PWKNN_EXACT_RERANK_F32 on every shard: raw (un-rounded) f32 inputs give the oracle's ids and distances."""
    rng = np.random.default_rng(16)
    d, n = 80, 3000
    rows = rng.standard_normal((n, d)).astype(np.float32).astype(np.float64)
    gpu = B200BruteForceKNNIndex(d, 8, metric=COS, devices=[0, 0, 0], exact_rerank="f32")
    ora = OracleInner(d, 8, 1024 * 128, COS, impl="c")
    data = [(ref_scalar(i), rows[i]) for i in range(n)]
    gpu.add(data), ora.add(data)
    assert gpu.stat("exact_rerank") == 1
    qs = [(j, rng.standard_normal(d).astype(np.float32).astype(np.float64), 10) for j in range(64)]
    _check_same(gpu.search(qs), ora.search(qs))
    gpu.close()


def test_bulk_load_through_shards():
    """Rows loaded straight into the shards (device-resident bf16, pwknn_add_device_bf16) are found by the router."""
    import torch

    from pathway_b200.external_index import array_to_keys, keys_to_array

    for devices in _device_sets():
        rng = np.random.default_rng(17)
        d, n = 128, 4000
        rows = _rounded(rng, (n, d))
        keys = keys_to_array([ref_scalar("bulk", i) for i in range(n)])
        gpu = B200BruteForceKNNIndex(d, 8, metric=COS, devices=devices)
        ora = OracleInner(d, 8, 1024 * 128, COS, impl="c")
        owner = gpu.shards_of_keys(keys)
        assert set(owner.tolist()) == set(range(len(devices)))
        with pytest.raises(IndexError):
            gpu.shard(len(devices))
        for g in range(len(devices)):
            sh = gpu.shard(g)
            mine = np.nonzero(owner == g)[0]
            t = torch.tensor(rows[mine], dtype=torch.float32, device=f"cuda:{sh.device}").to(torch.bfloat16)
            torch.cuda.synchronize(sh.device)
            assert (sh.add_device(keys[mine], t) == 0).all()
            np.testing.assert_array_equal(sh.read_rows(0, 3), rows[mine[:3]])
            sh.close()  # borrowed: must not destroy the shard
        ora.add(list(zip(array_to_keys(keys), rows)))
        assert gpu.current_size == n
        qs = [(j, _rounded(rng, d), 10) for j in range(50)]
        _check_same(gpu.search(qs), ora.search(qs), devices)
        # and the router still owns them: remove through the handle
        some = array_to_keys(keys[:40])
        assert all(e is None for _, e in gpu.remove(some))
        ora.remove(some)
        _check_same(gpu.search(qs), ora.search(qs), devices)
        gpu.close()


def test_snapshot_round_trip(tmp_path):
    rng = np.random.default_rng(18)
    d, n = 72, 2500
    rows = _rounded(rng, (n, d))
    data = [(ref_scalar("s", i), rows[i]) for i in range(n)]
    gpu = B200BruteForceKNNIndex(d, 8, metric=L2SQ, devices=[0, 0, 0])
    gpu.add(data)
    gpu.remove([k for k, _ in data[:300]])
    qs = [(j, _rounded(rng, d), 12) for j in range(30)]
    want = gpu.search(qs)
    path = str(tmp_path / "multi.pwknn")
    gpu.snapshot_save(path)
    size, allocated = gpu.current_size, gpu.current_allocated
    gpu.close()
    with pytest.raises(_native.PwknnError, match="pwknn_snapshot_load_multi"):
        B200BruteForceKNNIndex.snapshot_load(path)
    with pytest.raises(_native.PwknnError, match="3 shards"):
        B200BruteForceKNNIndex.snapshot_load(path, devices=[0, 0])
    back = B200BruteForceKNNIndex.snapshot_load(path, devices=[0, 0, 0])
    assert (back.current_size, back.current_allocated) == (size, allocated)
    _check_same(back.search(qs), want)
    # still a live index: deltas keep working after the restore
    assert all(e is None for _, e in back.add(data[:50]))
    assert back.current_size == size + 50
    back.close()


def test_placement_weights(tmp_path):
    """pwknn_multi_set_weights: hash ranges proportional to the weights, only on an empty index, kept by snapshots."""
    from pathway_b200.external_index import keys_to_array

    rng = np.random.default_rng(19)
    d, n = 48, 6000
    rows = _rounded(rng, (n, d))
    keys = [ref_scalar("w", i) for i in range(n)]
    gpu = B200BruteForceKNNIndex(d, 8, metric=COS, devices=[0, 0, 0], weights=[3.0, 1.0, 1.0])
    ora = OracleInner(d, 8, 1024 * 128, COS, impl="c")
    owner = gpu.shards_of_keys(keys_to_array(keys))
    share = np.bincount(owner, minlength=3) / n
    assert abs(share[0] - 0.6) < 0.03 and abs(share[1] - 0.2) < 0.03 and abs(share[2] - 0.2) < 0.03
    data = list(zip(keys, rows))
    gpu.add(data), ora.add(data)
    assert [gpu.shard(g).current_size for g in range(3)] == np.bincount(owner, minlength=3).tolist()
    with pytest.raises(_native.PwknnError, match="must be empty"):
        gpu.set_weights([1.0, 1.0, 1.0])
    qs = [(j, _rounded(rng, d), 9) for j in range(40)]
    _check_same(gpu.search(qs), ora.search(qs))
    path = str(tmp_path / "weighted.pwknn")
    gpu.snapshot_save(path)
    gpu.close()
    back = B200BruteForceKNNIndex.snapshot_load(path, devices=[0, 0, 0])
    np.testing.assert_array_equal(back.shards_of_keys(keys_to_array(keys)), owner)  # the weights came back
    gone = keys[:500]
    assert all(e is None for _, e in back.remove(gone))  # every key is found on the shard that holds it
    ora.remove(gone)
    _check_same(back.search(qs), ora.search(qs))
    back.close()
    with pytest.raises(_native.PwknnError):
        B200BruteForceKNNIndex(d, 8, metric=COS, devices=[0, 0], weights=[1.0, 0.0])
    with pytest.raises(_native.PwknnError):
        B200BruteForceKNNIndex(d, 8, metric=COS).set_weights([1.0])
