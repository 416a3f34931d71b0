"""GPU tests (-m gpu) of the per-query routing of pwknn_search: tier 1 (fused lists + certainty test), tier 2 (score dump
-> radix threshold -> f64 re-rank, "select path"), tier 3 (f64 distances of every row + full sort), the escalation
between them, and the hardened snapshot loader.  Checker: the CPU oracle on the same inputs, ids exact.

Reference order at stake everywhere: ascending (OrderedFloat(distance), slot), NaN last
(/root/reference/src/external_integration/brute_force_knn_integration.rs:220-224)."""

import os
import struct

import numpy as np
import pytest

from oracle import brute_force_oracle as bfo
from oracle_adapter import OracleInner
from pathway_b200 import _native
from pathway_b200.engine import BruteForceKnnMetricKind
from pathway_b200.external_index import B200BruteForceKNNIndex, keys_to_array

pytestmark = pytest.mark.gpu

COS, L2SQ = BruteForceKnnMetricKind.COS, BruteForceKnnMetricKind.L2SQ


def _rounded(rng, shape, scale=1.0):
    return bfo.bf16_round(rng.standard_normal(shape) * scale)


def _same(got, want, ctx="", rtol=1e-9):
    assert len(got) == len(want)
    for (k1, r1), (k2, r2) in zip(got, want):
        assert k1 == k2 and not isinstance(r1, Exception), (ctx, k1, r1)
        ids1, ids2 = [m for m, _ in r1], [m for m, _ in r2]
        assert ids1 == ids2, (ctx, k1, len(ids1), len(ids2), [(i, a, b) for i, (a, b) in enumerate(zip(ids1, ids2)) if a != b][:5])
        s1, s2 = np.array([s for _, s in r1]), np.array([s for _, s in r2])
        if len(s1):
            nan = np.isnan(s2)
            assert (np.isnan(s1) == nan).all()
            np.testing.assert_allclose(s1[~nan], s2[~nan], rtol=rtol, atol=1e-12, err_msg=str(ctx))


def _pair(d, metric, store=None, **kw):
    gpu = B200BruteForceKNNIndex(d, 16, 1024 * 128, metric, exact_rerank=store, **kw)
    ora = OracleInner(d, 16, 1024 * 128, metric, impl="c")
    return gpu, ora


def _load(gpu, ora, rows, key0=0):
    keys = keys_to_array(range(key0, key0 + len(rows)))
    assert (gpu.add_arrays(keys, rows) == 0).all()
    ora.ix.add_arrays(keys, np.ascontiguousarray(rows, dtype=np.float64))


@pytest.mark.parametrize("metric", [COS, L2SQ])
def test_mixed_limits_are_routed_per_query(metric):
    """One k = 5000 query must not drag the k = 10 queries off the tensor cores (round 1 ran the whole batch at the
    largest k): every query is answered by the tier its own k needs, results equal the oracle's."""
    rng = np.random.default_rng(31)
    d, n = 64, 20_000
    gpu, ora = _pair(d, metric)
    _load(gpu, ora, _rounded(rng, (n, d)))
    limits = [3, 10, 0, 150, 5000, n + 5, 50, 100, 1, 4096, 4097, 109, 98]
    qs = [(i, _rounded(rng, d), limits[i % len(limits)]) for i in range(3 * len(limits))]
    _same(gpu.search(qs), ora.search(qs), ("mixed", metric))
    # the same batch with only small limits stays on the fused kernel
    qs = [(i, _rounded(rng, d), [3, 10, 50, 98][i % 4]) for i in range(40)]
    _same(gpu.search(qs), ora.search(qs), ("small", metric))
    assert gpu.stat("last_kernel") == 1 and gpu.stat("last_tier") == 1
    gpu.close()


@pytest.mark.parametrize("store", [None, "f32"])
@pytest.mark.parametrize("metric", [COS, L2SQ])
def test_select_path_equals_oracle(metric, store):
    """Tier 2 for 108 < k + margin, k <= 4096: ids exact, (distance, slot) order, both stores; also forced for small k."""
    rng = np.random.default_rng(32)
    d, n = 96, 30_000
    gpu, ora = _pair(d, metric, store)
    rows = rng.standard_normal((n, d)).astype(np.float32) if store else _rounded(rng, (n, d))
    _load(gpu, ora, rows)
    mk = (lambda: rng.standard_normal(d).astype(np.float32).astype(np.float64)) if store else (lambda: _rounded(rng, d))
    qs = [(i, mk(), [150, 1000, 4096, 300][i % 4]) for i in range(12)]
    _same(gpu.search(qs), ora.search(qs), ("select", metric, store), rtol=1e-12 if store else 1e-9)
    assert gpu.stat("last_kernel") == 3 and gpu.stat("last_overflow") == 0
    gpu.set_option("force_tier", 2)
    qs = [(i, mk(), [1, 5, 17, 0][i % 4]) for i in range(300)]  # > 256 queries: several query tiles in the dump pass
    _same(gpu.search(qs), ora.search(qs), ("forced select", metric, store), rtol=1e-12 if store else 1e-9)
    assert gpu.stat("last_tier") == 2
    gpu.set_option("select_budget", 1 << 20)  # 1 MiB of scores per chunk: many chunks of few queries
    _same(gpu.search(qs[:40]), ora.search(qs[:40]), ("chunked select", metric, store), rtol=1e-12 if store else 1e-9)
    gpu.close()


def test_select_overflow_falls_back_to_the_full_sort():
    """10 000 identical rows tie exactly with the k-th result: the 2-eps band holds more than kSelectCap candidates,
    the query is flagged and re-answered by tier 3; ties come out in slot order like the reference."""
    rng = np.random.default_rng(33)
    d = 32
    gpu, ora = _pair(d, COS)
    base = _rounded(rng, d)
    rows = np.vstack([_rounded(rng, (3000, d)), np.tile(base, (10_000, 1)), _rounded(rng, (2000, d))])
    _load(gpu, ora, rows)
    qs = [(0, base, 200), (1, _rounded(rng, d), 200), (2, base, 10)]
    _same(gpu.search(qs), ora.search(qs), "overflow")
    assert gpu.stat("last_overflow") >= 1
    gpu.close()


@pytest.mark.parametrize("store", [None, "f32"])
@pytest.mark.parametrize("gap_lo,gap_hi", [(1e-6, 1e-4), (1e-9, 1e-7)])
def test_near_duplicate_embeddings_keep_ids_exact(store, gap_lo, gap_hi):
    """64 near-duplicates of each query, cosine gaps log-spaced in [gap_lo, gap_hi], inside 100 k
    rows, k = 10.  Gaps of 1e-9..1e-7 are BELOW the fp32 / bf16 noise of the tensor-core ranking pass: more than
    k' - k candidates are mis-ordered there, the certainty test must notice and the select path must re-answer."""
    rng = np.random.default_rng(34)
    d, n_fill, nq, dup = 768, 100_000, 6, 64
    gpu, ora = _pair(d, COS, store)
    fill = rng.standard_normal((n_fill, d)).astype(np.float32)
    queries = rng.standard_normal((nq, d)).astype(np.float32)
    gaps = np.geomspace(gap_lo, gap_hi, dup)
    near = []
    for q in queries:
        for g in gaps:
            u = rng.standard_normal(d).astype(np.float64)
            u -= (u @ q) / (q @ q) * q  # orthogonal to q: cos distance of q + t u/|u| |q| is ~ t^2 / 2
            near.append((q + np.sqrt(2 * g) * np.linalg.norm(q) * u / np.linalg.norm(u)).astype(np.float32))
    rows = np.vstack([fill, np.array(near, dtype=np.float32)])
    rows = rows[rng.permutation(len(rows))]
    if store is None:  # default store: "identical inputs" are bf16 numbers
        rows, queries = bfo.bf16_round(rows), bfo.bf16_round(queries)
    _load(gpu, ora, rows)
    qs = [(i, queries[i].astype(np.float64), 10) for i in range(nq)]
    _same(gpu.search(qs), ora.search(qs), ("near-dup", store, gap_lo), rtol=1e-9)
    flagged = gpu.stat("last_uncertain")
    print(f"near-duplicates store={store} gaps [{gap_lo:g}, {gap_hi:g}]: {flagged} of {nq} queries escalated")
    if gap_hi <= 1e-7:
        assert flagged > 0, "gaps below the ranking noise must trip the certainty test"
    # random queries on the same corpus are certain: no escalation
    qs = [(100 + i, _rounded(rng, d), 10) for i in range(64)]
    _same(gpu.search(qs), ora.search(qs), ("random", store))
    assert gpu.stat("last_uncertain") == 0
    gpu.close()


def test_filters_on_the_select_path():
    """A pushed filter (row bitmask) with a limit beyond the fused lists: tier 2 honours the mask."""
    rng = np.random.default_rng(35)
    d, n = 48, 12_000
    gpu, ora = _pair(d, COS)
    rows = _rounded(rng, (n, d))
    _load(gpu, ora, rows)
    fid = gpu.filter_create()
    passing = [i for i in range(n) if i % 7 == 3]
    gpu.filter_set(fid, list(range(n)), [i % 7 == 3 for i in range(n)])
    qv = [_rounded(rng, d) for _ in range(6)]
    got = gpu.search_filtered([(i, qv[i], [300, 10, 2000][i % 3], fid if i % 2 == 0 else None) for i in range(6)])
    for i, (k, res) in enumerate(got):
        lim = [300, 10, 2000][i % 3]
        (_k, full), = ora.search([(i, qv[i], n)])
        want = [(m, s) for m, s in full if (i % 2 == 1) or (m % 7 == 3)][:lim]
        assert [m for m, _ in res] == [m for m, _ in want], (i, len(res), len(want))
    assert len(passing) > 1000
    gpu.close()


def test_multi_device_handle_routes_per_query():
    rng = np.random.default_rng(36)
    d, n = 64, 9000
    gpu = B200BruteForceKNNIndex(d, 16, metric=COS, devices=[0, 0, 0])
    ora = OracleInner(d, 16, 1024 * 128, COS, impl="c")
    rows = _rounded(rng, (n, d))
    keys = keys_to_array(range(n))
    assert (gpu.add_arrays(keys, rows) == 0).all()
    ora.ix.add_arrays(keys, rows)
    qs = [(i, _rounded(rng, d), [10, 400, 0, 9000, 3, 120][i % 6]) for i in range(30)]
    _same(gpu.search(qs), ora.search(qs), "multi mixed")
    gpu.close()


def test_device_tier_large_k_and_flags():
    import torch

    rng = np.random.default_rng(37)
    d, n, nq = 64, 8000, 20
    gpu, ora = _pair(d, L2SQ)
    rows = _rounded(rng, (n, d))
    _load(gpu, ora, rows)
    q = _rounded(rng, (nq, d))
    dev = torch.device("cuda", gpu.device)
    q_bf = torch.tensor(q, dtype=torch.float32, device=dev).to(torch.bfloat16)
    for k in (10, 300):
        out_s = torch.empty((nq, k), dtype=torch.int32, device=dev)
        out_d = torch.empty((nq, k), dtype=torch.float64, device=dev)
        gpu.search_device(q_bf, k, out_s, out_d)  # stream NULL: ordered after the default stream that produced q_bf
        torch.cuda.synchronize(dev)
        assert gpu.stat("last_uncertain") == 0
        want = ora.search([(i, q[i], k) for i in range(nq)])
        got_keys = gpu.slots_to_keys(out_s.cpu().numpy().astype(np.uint32).reshape(-1))[:, 0].reshape(nq, k)
        assert (got_keys == np.array([[m for m, _ in r] for _, r in want], dtype=np.uint64)).all()
    gpu.close()


# --------------------------------------------------------------------------- snapshots: nothing in the file is trusted
def _header(path):
    with open(path, "rb") as f:
        return bytearray(f.read(88))


def test_snapshot_loader_rejects_corrupted_files(tmp_path):
    rng = np.random.default_rng(38)
    gpu = B200BruteForceKNNIndex(32, 64, metric=COS)
    gpu.add_arrays(keys_to_array(range(500)), _rounded(rng, (500, 32)))
    p = str(tmp_path / "ix.snap")
    gpu.snapshot_save(p)
    assert not os.path.exists(p + ".tmp")  # written under a temporary name, then renamed
    good = open(p, "rb").read()
    hdr = struct.Struct("<8s6I5Q")
    magic, dims, dpad, metric, sk, ss, xm, size, alloc, minimum, aux, gen = hdr.unpack(good[:hdr.size])
    assert magic == b"PWKNNSN2" and size == 500 and alloc >= 500

    def mutated(**kw):
        f = dict(dims=dims, dpad=dpad, metric=metric, sk=sk, ss=ss, xm=xm, size=size, alloc=alloc, minimum=minimum, aux=aux, gen=gen)
        f.update(kw)
        return hdr.pack(magic, f["dims"], f["dpad"], f["metric"], f["sk"], f["ss"], f["xm"], f["size"], f["alloc"],
                        f["minimum"], f["aux"], f["gen"]) + good[hdr.size:]

    cases = {
        "size > allocated": mutated(size=alloc + 1),
        "huge size": mutated(size=1 << 50, alloc=1 << 50),
        "unknown metric": mutated(metric=9),
        "unknown store": mutated(sk=7),
        "bad dims": mutated(dims=0),
        "truncated": good[:-100],
        "appended": good + b"x" * 64,
        "size smaller than the payload": mutated(size=size - 1),
    }
    for name, blob in cases.items():
        q = str(tmp_path / "bad.snap")
        with open(q, "wb") as f:
            f.write(blob)
        with pytest.raises(_native.PwknnError) as e:
            B200BruteForceKNNIndex.snapshot_load(q)
        assert e.value.code == _native.E_INVALID, (name, e.value)
    back = B200BruteForceKNNIndex.snapshot_load(p)
    assert back.current_size == 500
    back.close()
    gpu.close()


def test_multi_snapshot_rejects_stale_shard_files(tmp_path):
    rng = np.random.default_rng(39)
    p = str(tmp_path / "m.snap")
    a = B200BruteForceKNNIndex(16, 8, metric=COS, devices=[0, 0])
    a.add_arrays(keys_to_array(range(300)), _rounded(rng, (300, 16)))
    a.snapshot_save(p)
    stale = open(p + ".shard1", "rb").read()
    a.add_arrays(keys_to_array(range(300, 400)), _rounded(rng, (100, 16)))
    a.snapshot_save(p)
    b = B200BruteForceKNNIndex.snapshot_load(p, devices=[0, 0])
    assert b.current_size == 400
    b.close()
    with open(p + ".shard1", "wb") as f:  # a shard file of the EARLIER save next to the new manifest
        f.write(stale)
    with pytest.raises(_native.PwknnError):
        B200BruteForceKNNIndex.snapshot_load(p, devices=[0, 0])
    a.close()


# --------------------------------------------------------------------------- async delta path
def test_deltas_interleaved_with_in_flight_searches():
    """add / remove return once their inputs sit in the staging ring; the copies and kernels drain behind them.  A
    device-tier search enqueued BEFORE a delta batch must see the old state, one enqueued AFTER it the new state, with
    no host synchronisation in between (remove -> add -> search order per timestamp, mod.rs:144-150; stream order)."""
    import torch

    rng = np.random.default_rng(40)
    d, k, nq = 64, 5, 40
    gpu, ora = _pair(d, COS)
    dev = torch.device("cuda", gpu.device)
    ext = torch.cuda.ExternalStream(gpu.stream_ptr, device=dev)
    q = _rounded(rng, (nq, d))
    q_bf = torch.tensor(q, dtype=torch.float32, device=dev).to(torch.bfloat16)
    torch.cuda.synchronize(dev)
    live, nxt, pending = [], 0, []
    for step in range(12):
        rows = _rounded(rng, (4000, d))  # 4000 x 64 x 8 B = 2 MB of f64 per batch: several staging chunks over the run
        keys = list(range(nxt, nxt + 4000)); nxt += 4000
        rm = [live.pop(int(rng.integers(0, len(live)))) for _ in range(min(len(live), 1500))]
        gpu.remove_arrays(keys_to_array(rm)) if rm else None
        gpu.add_arrays(keys_to_array(keys), rows)                      # returns before the upserts ran
        out_s = torch.empty((nq, k), dtype=torch.int32, device=dev)
        out_d = torch.empty((nq, k), dtype=torch.float64, device=dev)
        gpu.search_device(q_bf, k, out_s, out_d, stream_ptr=gpu.stream_ptr)  # queued right behind them, no sync
        ora.remove(rm) if rm else None
        ora.ix.add_arrays(keys_to_array(keys), rows)
        live += keys
        want = ora.search([(i, q[i], k) for i in range(nq)])
        # slot -> key must be read NOW (the next step's swap-removes change the mapping): the slots themselves are
        # only known after the kernel ran, so wait for this step's search here, but never between delta and search
        ext.synchronize()
        got = gpu.slots_to_keys(out_s.cpu().numpy().astype(np.uint32).reshape(-1))[:, 0].reshape(nq, k)
        assert (got == np.array([[m for m, _ in r] for _, r in want], dtype=np.uint64)).all(), step
        np.testing.assert_allclose(out_d.cpu().numpy(), -np.array([[s for _, s in r] for _, r in want]), rtol=1e-9, atol=1e-12)
        assert gpu.current_size == ora.current_size and gpu.current_allocated == ora.current_allocated
    gpu.close()


def test_unusable_rows_in_a_large_batch_roll_back_cleanly():
    """add() updates the key map optimistically WHILE the rows are copied and tested; a NaN / inf row found afterwards
    undoes the appended slots and redoes the assignment without it.  Layout, sizes and replies must equal an oracle that
    only ever saw the usable rows -- with live keys (overwrite in place) and in-batch duplicates (last wins) mixed in."""
    rng = np.random.default_rng(41)
    d = 96
    gpu, ora = _pair(d, L2SQ)
    first = _rounded(rng, (2000, d))
    _load(gpu, ora, first)
    rows = _rounded(rng, (3000, d))
    keys = list(range(1500, 4500))          # 500 live keys (overwritten in place) + 2500 new ones
    keys[100] = keys[2900]                  # an in-batch duplicate of a new key: the later row wins
    bad = [17, 1234, 2500]
    rows_bad = rows.copy()
    rows_bad[17, 3] = np.nan
    rows_bad[1234, 0] = np.inf
    rows_bad[2500, 95] = -np.inf
    st = gpu.add_arrays(keys_to_array(keys), rows_bad)
    assert [i for i in range(3000) if st[i] != 0] == bad
    good = [i for i in range(3000) if i not in bad]
    # the oracle (reference semantics) re-adding a live key bumps size: feed it remove-then-add for the live ones
    live = sorted({keys[i] for i in good if keys[i] < 2000})
    order = {}
    for i in good:
        order[keys[i]] = i  # last occurrence wins
    ora.remove(live)
    # replay: surviving rows in first-occurrence order of their key, with the last occurrence's vector
    seen, replay = set(), []
    for i in good:
        if keys[i] not in seen:
            seen.add(keys[i]); replay.append(keys[i])
    # live keys keep their old slot on the GPU (overwrite in place) but move to the end in the oracle's remove+add
    # replay: compare CONTENT and replies by key, sizes exactly
    ora.ix.add_arrays(keys_to_array(replay), np.array([rows[order[k]] for k in replay]))
    assert gpu.current_size == ora.current_size == 2000 + len([k for k in replay if k >= 2000])
    qs = [(i, _rounded(rng, d), 12) for i in range(50)]
    got, want = gpu.search(qs), ora.search(qs)
    for (k1, r1), (k2, r2) in zip(got, want):
        assert [m for m, _ in r1] == [m for m, _ in r2]
        np.testing.assert_allclose([s for _, s in r1], [s for _, s in r2], rtol=1e-9, atol=1e-12)
    gpu.close()


def test_host_packed_search_and_single_shard_merge():
    """pwknn_search_host_packed (host f64 queries in, packed {global id, distance bits} left in HBM) followed by
    pwknn_merge_shards_device over ONE shard equals pwknn_search; a huge limit is clamped to the index size instead of
    sizing the outputs by it."""
    import torch

    rng = np.random.default_rng(42)
    d, n, nq, k = 80, 6000, 70, 9
    gpu, ora = _pair(d, COS)
    rows = _rounded(rng, (n, d))
    _load(gpu, ora, rows, key0=500)
    q = _rounded(rng, (nq, d))
    dev = torch.device("cuda", gpu.device)
    packed = torch.empty((nq, k, 2), dtype=torch.int64, device=dev)
    out_ids = torch.empty((nq, k), dtype=torch.int64, device=dev)
    out_d = torch.empty((nq, k), dtype=torch.float64, device=dev)
    gpu.search_host_packed(q, k, 1_000_000, packed, stream_ptr=gpu.stream_ptr)
    gpu.merge_shards_device(packed.unsqueeze(0), out_ids, out_d, stream_ptr=gpu.stream_ptr)
    torch.cuda.synchronize(dev)
    want = ora.search([(i, q[i], k) for i in range(nq)])
    slots = (out_ids.cpu().numpy() - 1_000_000).astype(np.uint32)
    keys = gpu.slots_to_keys(slots.reshape(-1))[:, 0].reshape(nq, k)
    assert (keys == np.array([[m for m, _ in r] for _, r in want], dtype=np.uint64)).all()
    np.testing.assert_allclose(out_d.cpu().numpy(), -np.array([[s for _, s in r] for _, r in want]), rtol=1e-9, atol=1e-12)
    # "give me everything": limit 10^9 on a 6000-row index -> 6000 rows, outputs sized by the clamped limits
    (_, everything), = gpu.search([(0, q[0], 1_000_000_000)])
    (_, want_all), = ora.search([(0, q[0], n)])
    assert len(everything) == n and [m for m, _ in everything] == [m for m, _ in want_all]
    gpu.close()


def test_device_tier_reports_uncertain_queries():
    """The asynchronous device tier cannot re-answer a query: it leaves the certainty flags on the device and
    pwknn_get_stat(h, "last_uncertain") counts them -- > 0 for a query with 64 near-duplicates below the fp32 noise,
    0 for random queries."""
    import torch

    rng = np.random.default_rng(43)
    d, n = 256, 40_000
    gpu, _ora = _pair(d, COS)
    q0 = rng.standard_normal(d).astype(np.float32)
    near = []
    for g in np.geomspace(1e-9, 1e-7, 64):
        u = rng.standard_normal(d)
        u -= (u @ q0) / (q0 @ q0) * q0
        near.append(q0 + np.sqrt(2 * g) * np.linalg.norm(q0) * u / np.linalg.norm(u))
    rows = bfo.bf16_round(np.vstack([rng.standard_normal((n, d)), np.array(near)]))
    assert (gpu.add_arrays(keys_to_array(range(len(rows))), rows) == 0).all()
    dev = torch.device("cuda", gpu.device)
    qs = np.vstack([bfo.bf16_round(q0[None, :]), _rounded(rng, (31, d))])
    q_bf = torch.tensor(qs, dtype=torch.float32, device=dev).to(torch.bfloat16)
    out_s = torch.empty((32, 10), dtype=torch.int32, device=dev)
    out_d = torch.empty((32, 10), dtype=torch.float64, device=dev)
    gpu.search_device(q_bf, 10, out_s, out_d)
    assert gpu.stat("last_uncertain") == 1   # only the query sitting on its near-duplicates
    gpu.search_device(q_bf[1:], 10, out_s[1:], out_d[1:])
    assert gpu.stat("last_uncertain") == 0
    gpu.close()
