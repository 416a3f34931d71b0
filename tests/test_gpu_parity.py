"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI (ctypes), against the CPU
oracle on the same seeded inputs.

Bar (BASELINE.json north_star): returned ids bit-exact, distances within 1e-3 relative (we assert
1e-9: the final re-rank is f64 on both sides).  "Identical inputs" = vectors pre-rounded to bf16
(oracle.brute_force_oracle.bf16_round), the precision the HBM corpus stores.
"""

import numpy as np
import pytest

from golden import kats
from kat_harness import run_case
from oracle import brute_force_oracle as bfo
from oracle_adapter import OracleInner
from pathway_b200.debug import ref_scalar, table_from_rows
from pathway_b200.engine import BruteForceKnnMetricKind
from pathway_b200.external_index import B200BruteForceKNNIndex, MissingKey, keys_to_array
from pathway_b200.stdlib.indexing import BruteForceKnn

pytestmark = pytest.mark.gpu

COS, L2SQ = BruteForceKnnMetricKind.COS, BruteForceKnnMetricKind.L2SQ
REL_TOL = 1e-9  # north_star allows 1e-3 relative; the f64 re-rank makes it far tighter
ABS_TOL = 1e-12


def _rounded(rng, shape, scale=1.0):
    return bfo.bf16_round(rng.standard_normal(shape) * scale)


def _check_same(got, want, ctx=""):
    """got/want: [(key, [(match_key, score)])] in query order."""
    assert len(got) == len(want)
    for (k1, r1), (k2, r2) in zip(got, want):
        assert k1 == k2
        assert not isinstance(r1, Exception), (ctx, r1)
        ids1, ids2 = [m for m, _ in r1], [m for m, _ in r2]
        assert ids1 == ids2, (ctx, k1, ids1[:12], ids2[:12])
        s1, s2 = np.array([s for _, s in r1]), np.array([s for _, s in r2])
        if len(s1):
            nan = np.isnan(s2)
            assert (np.isnan(s1) == nan).all(), (ctx, s1, s2)
            np.testing.assert_allclose(s1[~nan], s2[~nan], rtol=REL_TOL, atol=ABS_TOL, err_msg=str(ctx))


def _pair(dims, reserved, metric, **kw):
    gpu = B200BruteForceKNNIndex(dims, reserved, 1024 * 128, metric, **kw)
    ora = OracleInner(dims, reserved, 1024 * 128, metric, impl="c")
    return gpu, ora


# --------------------------------------------------------------------------- plumbing first
def test_create_add_read_rows_roundtrip():
    rng = np.random.default_rng(1)
    gpu = B200BruteForceKNNIndex(70, 4, metric=COS)
    rows = rng.standard_normal((300, 70))
    keys = list(range(1000, 1300))
    res = gpu.add([(k, r) for k, r in zip(keys, rows)])
    assert all(e is None for _, e in res)
    assert gpu.current_size == 300 and gpu.current_allocated == 300  # max(2*4, 300)
    back = gpu.read_rows(0, 300)
    np.testing.assert_array_equal(back, bfo.bf16_round(rows))
    tail = gpu.read_rows(300, gpu.stat("phys_rows") - 300)
    assert (tail == 0).all()  # rows >= size stay zero (kernel invariant)
    gpu.close()


def test_exact_path_small():
    """kernel=2: f64 CUDA-core distances + stable segmented sort (the any-k path), both metrics."""
    rng = np.random.default_rng(2)
    for metric in (COS, L2SQ):
        gpu, ora = _pair(24, 8, metric)
        gpu.set_option("kernel", 2)
        rows = _rounded(rng, (500, 24))
        data = [(ref_scalar(i), rows[i]) for i in range(500)]
        gpu.add(data), ora.add(data)
        qs = [(i, _rounded(rng, 24), [0, 1, 7, 500, 900][i % 5]) for i in range(23)]
        _check_same(gpu.search(qs), ora.search(qs), metric)
        assert gpu.stat("last_kernel") == 2
        gpu.close()


# --------------------------------------------------------------------------- the fused tcgen05 kernel
@pytest.mark.parametrize("n,d,nq", [(700, 128, 200), (256, 64, 1), (5000, 768, 130), (9, 3, 4)])
def test_tcgen05_scores_vs_torch(n, d, nq):
    """Raw fp32 ranking scores of the TMA + tcgen05 pass against a plain torch fp32 reference of the
    same op (bf16 inputs, fp32 accumulate): catches descriptor / swizzle / pipeline errors directly."""
    import torch

    rng = np.random.default_rng(3)
    gpu = B200BruteForceKNNIndex(d, n, metric=COS)
    rows = _rounded(rng, (n, d))
    gpu.add_arrays(keys_to_array(range(n)), rows)
    q = _rounded(rng, (nq, d))
    dev = torch.device("cuda", gpu.device)
    q_bf = torch.tensor(q, dtype=torch.float32, device=dev).to(torch.bfloat16)
    out = torch.full((nq, n), float("nan"), dtype=torch.float32, device=dev)
    torch.cuda.synchronize(dev)
    gpu.debug_scores_device(q_bf, out)
    torch.cuda.synchronize(dev)
    x = torch.tensor(rows, dtype=torch.float32, device=dev)
    ref = (q_bf.float() @ x.T) / x.double().pow(2).sum(1).sqrt().float()[None, :]
    err = (out - ref).abs().max().item()
    assert not torch.isnan(out).any(), f"unwritten scores: {torch.isnan(out).sum().item()} of {out.numel()}"
    assert err < 2e-3 * max(1.0, ref.abs().max().item()), f"max |score - ref| = {err}"
    gpu.close()


@pytest.mark.parametrize("n,d,nq,k", [(5000, 768, 300, 10), (900, 64, 130, 3), (40000, 200, 513, 20), (3000, 96, 260, 50)])
@pytest.mark.parametrize("metric", [COS, L2SQ])
def test_cta_pair_variant(metric, n, d, nq, k):
    """cta_group::2: a CTA pair shares each corpus tile (each CTA loads half), leader issues M=256 MMAs.
    Raw scores vs torch and end-to-end parity vs the oracle, including an odd number of query tiles."""
    import torch

    rng = np.random.default_rng(n + nq)
    gpu, ora = _pair(d, 16, metric)
    gpu.set_option("cta_pair", 2)
    rows = _rounded(rng, (n, d))
    data = [(ref_scalar("p", i), rows[i]) for i in range(n)]
    gpu.add(data), ora.add(data)
    q = _rounded(rng, (nq, d))
    if metric is COS:
        dev = torch.device("cuda", gpu.device)
        q_bf = torch.tensor(q, dtype=torch.float32, device=dev).to(torch.bfloat16)
        out = torch.full((nq, n), float("nan"), dtype=torch.float32, device=dev)
        torch.cuda.synchronize(dev)
        gpu.debug_scores_device(q_bf, out)
        torch.cuda.synchronize(dev)
        assert gpu.stat("last_ctas") == 2
        x = torch.tensor(rows, dtype=torch.float32, device=dev)
        ref = (q_bf.float() @ x.T) / x.double().pow(2).sum(1).sqrt().float()[None, :]
        assert not torch.isnan(out).any(), f"unwritten scores: {torch.isnan(out).sum().item()}"
        assert (out - ref).abs().max().item() < 2e-3 * max(1.0, ref.abs().max().item())
    qs = [(i, q[i], k if i % 5 else 0) for i in range(nq)]
    _check_same(gpu.search(qs), ora.search(qs), ("pair", metric, n, d))
    assert gpu.stat("last_kernel") == 1 and gpu.stat("last_ctas") == 2
    gpu.close()


@pytest.mark.parametrize("name", sorted(kats.CASES))
def test_reference_kat_gpu(name):
    """The reference's own known-answer tests (tests/golden/kats.py) through the B200 index."""
    made = {}

    def make(**kw):
        made["ix"] = B200BruteForceKNNIndex(kw["dimensions"], kw["reserved_space"], kw["auxiliary_space"], kw["metric"])
        return made["ix"]

    run_case(name, make)
    if name == "space_shenanigans_resize":
        assert made["ix"].current_allocated == 12
    if name == "resize_after_delete":
        assert made["ix"].current_size == 9 and made["ix"].current_allocated == 12
    assert made["ix"].stat("last_kernel") == 1  # went through the tcgen05 kernel
    made["ix"].close()


def test_reference_stream_and_multi_removal_gpu():
    """test_index.py:1104-1147 and :1150-1190 through the B200 index."""
    points = table_from_rows(["coords"], [(np.array(c), t, 1) for c, t in kats.STREAM_POINTS], is_stream=True)
    queries = table_from_rows(["coords"], [(np.array(c), t, 1) for c, t in kats.STREAM_QUERIES], is_stream=True)
    index = BruteForceKnn(data_column=points.coords, dimensions=2, reserved_space=3, metric=L2SQ)
    res = index.query_as_of_now(queries.coords, number_of_matches=2)
    assert res.error_log == []
    coords_of = {r.key: tuple(int(x) for x in r.values["coords"]) for r in points.rows}
    got = {tuple(int(x) for x in r.values["coords"]): {coords_of[k] for k, _ in r.values["_pw_index_reply"]}
           for r in res.rows}
    assert got == kats.STREAM_EXPECTED_K2

    data = table_from_rows(["a", "v"], [(a, np.array(v), t, d) for a, v, t, d in kats.MULTI_REMOVAL_DATA],
                           id_from=["a"], is_stream=True)
    qs = table_from_rows(["a", "v"], [(a, np.array(v), t, d) for a, v, t, d in kats.MULTI_REMOVAL_QUERIES],
                         id_from=["a"], is_stream=True)
    index = BruteForceKnn(data_column=data.v, dimensions=3, reserved_space=0, metric=COS)
    res = index.query_as_of_now(qs.v, number_of_matches=1)
    assert res.error_log == []
    got = {r.values["a"]: [k for k, _ in r.values["_pw_index_reply"]] for r in res.rows}
    assert got == {q: [ref_scalar(d)] for q, d in kats.MULTI_REMOVAL_EXPECTED.items()}


@pytest.mark.parametrize("metric", [COS, L2SQ])
@pytest.mark.parametrize("n,d,nq", [(3000, 96, 300), (1, 5, 3), (257, 64, 129), (20000, 200, 64)])
def test_random_parity(metric, n, d, nq):
    rng = np.random.default_rng(n * 7 + d)
    gpu, ora = _pair(d, 16, metric)
    rows = _rounded(rng, (n, d))
    data = [(ref_scalar("k", i), rows[i]) for i in range(n)]
    gpu.add(data), ora.add(data)
    qs = [(i, _rounded(rng, d), int(rng.integers(0, 21))) for i in range(nq)]
    _check_same(gpu.search(qs), ora.search(qs), (metric, n, d))
    assert gpu.stat("last_kernel") == 1
    gpu.close()


@pytest.mark.parametrize("metric", [COS, L2SQ])
@pytest.mark.parametrize("k", [27, 40, 57, 100])
@pytest.mark.parametrize("optimistic", [1, 0])
def test_mid_k_fused_paths(metric, k, optimistic):
    """32 < k + margin <= 108 (HybridIndex-style top-100 re-rank, config C4).  optimistic=1: pass a keeps 16-entry
    register lists, the merge flags queries whose candidates may be incomplete, pass b redoes flagged groups with
    full shared-memory lists.  optimistic=0: one pass with full (warp-cooperative) shared-memory lists."""
    rng = np.random.default_rng(k)
    gpu, ora = _pair(48, 8, metric)
    gpu.set_option("optimistic", optimistic)
    rows = _rounded(rng, (7000, 48))
    data = [(i + 5, rows[i]) for i in range(7000)]
    gpu.add(data), ora.add(data)
    qs = [(i, _rounded(rng, 48), k if i % 3 else k // 2) for i in range(140)]
    _check_same(gpu.search(qs), ora.search(qs), (metric, k, optimistic))
    assert gpu.stat("last_kernel") == 1
    if k + 6 > 32:
        assert gpu.stat("last_two_pass") == optimistic
        assert gpu.stat("last_kprime") == (16 if optimistic else k + max(6, k // 8 if k <= 64 else k // 16))
    gpu.close()


@pytest.mark.parametrize("metric", [COS, L2SQ])
def test_optimistic_pass_falls_back_when_a_slab_holds_the_whole_answer(metric):
    """Adversarial layout for the optimistic scheme: the 100 nearest rows of some queries are contiguous (one
    slab), so their 16-entry lists are exhausted, the queries are flagged and redone with full lists; other
    queries of the same batch stay on the fast pass.  Results must still equal the oracle's."""
    rng = np.random.default_rng(77)
    d, n, k = 32, 30000, 100
    gpu, ora = _pair(d, 8, metric)
    rows = _rounded(rng, (n, d))
    centre = _rounded(rng, d) * 4
    rows[5000:5300] = bfo.bf16_round(centre[None, :] + 0.05 * rng.standard_normal((300, d)))
    data = [(i, rows[i]) for i in range(n)]
    gpu.add(data), ora.add(data)
    qs = [(i, centre if i < 7 else _rounded(rng, d), k) for i in range(260)]
    got, want = gpu.search(qs), ora.search(qs)
    assert gpu.stat("last_two_pass") == 1
    flagged = gpu.stat("last_incomplete")
    assert 7 <= flagged < 260, flagged
    _check_same(got, want, "optimistic fallback")
    gpu.close()


@pytest.mark.parametrize("n,nq,k", [(150, 5, 100), (50000, 1, 100), (257, 130, 57), (40, 300, 33), (20000, 129, 64)])
def test_large_k_edge_shapes(n, nq, k):
    """k close to / above the index size, a single query, odd tile counts: whichever plan is picked
    (optimistic two-pass, full smem lists, CTA pair or single) must agree with the oracle."""
    rng = np.random.default_rng(n + k)
    gpu, ora = _pair(40, 4, COS)
    rows = _rounded(rng, (n, 40))
    data = [(i + 1, rows[i]) for i in range(n)]
    gpu.add(data), ora.add(data)
    qs = [(i, _rounded(rng, 40), k) for i in range(nq)]
    _check_same(gpu.search(qs), ora.search(qs), (n, nq, k))
    assert gpu.stat("last_kernel") == 1
    gpu.close()


def test_config3_shape_dims_1536_k8():
    """BASELINE.json configs[2] shape per shard (D = 1536, k = 8), reduced rows: parity with the oracle."""
    rng = np.random.default_rng(15)
    gpu, ora = _pair(1536, 8, COS)
    rows = _rounded(rng, (2500, 1536))
    data = [(i, rows[i]) for i in range(2500)]
    gpu.add(data), ora.add(data)
    qs = [(i, _rounded(rng, 1536), 8) for i in range(64)]
    _check_same(gpu.search(qs), ora.search(qs), "C3 shape")
    gpu.close()


def test_slab_count_does_not_change_results():
    rng = np.random.default_rng(5)
    gpu, ora = _pair(64, 8, COS)
    rows = _rounded(rng, (6000, 64))
    data = [(i + 1, rows[i]) for i in range(6000)]
    gpu.add(data), ora.add(data)
    qs = [(i, _rounded(rng, 64), 10) for i in range(40)]
    want = ora.search(qs)
    for slabs in (1, 2, 5, 24):
        gpu.set_option("slabs", slabs)
        _check_same(gpu.search(qs), want, f"slabs={slabs}")
        assert gpu.stat("last_slabs") == slabs
    gpu.close()


def test_ties_duplicates_zero_vectors_and_limits():
    """Equal distances are ordered by slot; zero vectors give NaN cosine distance, ranked last
    (ordered-float), limit > size returns size matches, limit 0 returns none."""
    gpu, ora = _pair(8, 2, COS)
    base = np.array([1, 2, 3, 4, 0.5, 0.25, 0, 1.0])
    data = [(100 + i, base) for i in range(40)]  # duplicate documents: all distances tie exactly
    data += [(200 + i, np.zeros(8)) for i in range(3)]  # NaN distances
    data += [(300, -base)]
    gpu.add(data), ora.add(data)
    qs = [(0, base, 5), (1, base, 44), (2, base, 100), (3, -base, 2), (4, base, 0), (5, np.zeros(8), 3)]
    _check_same(gpu.search(qs), ora.search(qs), "ties")
    gpu.close()


def test_churn_matches_oracle_layout():
    """Interleaved add / swap-remove batches: identical size, allocation and slot layout
    (brute_force_knn_integration.rs:114-187), identical search results, MissingKey on unknown keys."""
    rng = np.random.default_rng(6)
    d = 32
    gpu, ora = _pair(d, 4, L2SQ)
    live, next_key = [], 1
    for step in range(30):
        n_add = int(rng.integers(0, 200))
        rows = _rounded(rng, (n_add, d))
        batch = [(next_key + i, rows[i]) for i in range(n_add)]
        next_key += n_add
        n_rm = int(rng.integers(0, min(len(live), 150) + 1))
        rm = [live.pop(int(rng.integers(0, len(live)))) for _ in range(n_rm)] if live else []
        if step % 5 == 0 and live:
            rm.append(10 ** 9 + step)  # unknown key -> MissingKey, state unchanged
        r_g, r_o = gpu.remove(rm), ora.remove(rm)
        assert [(k, type(e)) for k, e in r_g] == [(k, type(e)) for k, e in r_o]
        if step % 5 == 0 and rm:
            assert isinstance(r_g[-1][1], MissingKey)
        gpu.add(batch), ora.add(batch)
        live += [k for k, _ in batch]
        assert gpu.current_size == ora.current_size == len(live)
        assert gpu.current_allocated == ora.current_allocated
        if gpu.current_size:
            np.testing.assert_array_equal(gpu.read_rows(0, gpu.current_size), _oracle_rows(ora, d))
            qs = [(i, _rounded(rng, d), 7) for i in range(5)]
            _check_same(gpu.search(qs), ora.search(qs), f"step {step}")
    gpu.close()


def _oracle_rows(ora, d):
    import ctypes as C

    from oracle.c_oracle import lib

    n = ora.current_size
    ptr = lib().pwo_rows(ora.ix._h)
    return np.ctypeslib.as_array((C.c_double * (n * d)).from_address(ptr)).reshape(n, d).copy()


def test_large_k_routes_to_exact_path():
    rng = np.random.default_rng(8)
    gpu, ora = _pair(16, 8, COS)
    rows = _rounded(rng, (3000, 16))
    data = [(i, rows[i]) for i in range(3000)]
    gpu.add(data), ora.add(data)
    qs = [(0, _rounded(rng, 16), 250), (1, _rounded(rng, 16), 3000), (2, _rounded(rng, 16), 2)]
    _check_same(gpu.search(qs), ora.search(qs), "large k")
    assert gpu.stat("last_kernel") == 3  # per-query routing: the k = 2 query stays fused, the large ones take the select path
    gpu.close()


def test_errors_are_per_row_not_fatal():
    gpu = B200BruteForceKNNIndex(4, 2, metric=COS)
    res = gpu.add([(1, [1.0, 0, 0, 0]), (2, [1.0, 2.0]), (3, [np.nan, 0, 0, 1.0]), (4, [0, 1.0, 0, 0]),
                   (5, [1e39, 0, 0, 0]), (6, [np.inf, 0, 0, 0])])  # 1e39 overflows the bf16/f32 range
    assert [e is None for _, e in res] == [True, False, False, True, False, False]
    assert gpu.current_size == 2
    big = np.zeros((20000, 4)); big[:, 0] = 1.0; big[777, 1] = 1e40  # vectorised check on a > 1 MiB batch
    st = gpu.add_arrays(keys_to_array(range(100, 20100)), big)
    assert st[777] == 2 and (np.delete(st, 777) == 0).all() and gpu.current_size == 2 + 19999
    gpu.remove([100 + i for i in range(20000) if i != 777])
    out = gpu.search([(9, [1.0, 0, 0, 0], 2), (10, [1.0], 1)])
    assert [m for m, _ in out[0][1]] == [1, 4]
    assert isinstance(out[1][1], Exception)
    assert gpu.remove([77])[0][1].__class__ is MissingKey
    gpu.close()


def test_filtered_search_k_doubling_gpu():
    """DerivedFilteredSearchIndex over the GPU index: JMESPath / globmatch filters with the k-doubling loop
    (src/external_integration/mod.rs:457-504,583-601)."""
    from pathway_b200.external_index import AddDataEntry, DerivedFilteredSearchIndex, QueryEntry

    gpu = B200BruteForceKNNIndex(3, 4, metric=L2SQ)
    ix = DerivedFilteredSearchIndex(gpu)
    n = 64
    ix.add([AddDataEntry(i, (float(i), 0.0, 0.0), {"path": f"dir{i % 4}/f{i}.py", "k": i % 2}) for i in range(n)])
    res = dict(ix.search([
        QueryEntry(1, (0.0, 0.0, 0.0), 3, "globmatch('dir3/*', path)"),
        QueryEntry(2, (0.0, 0.0, 0.0), 2, "k == `1`"),
        QueryEntry(3, (0.0, 0.0, 0.0), 5, "k == `7`"),
        QueryEntry(4, (0.0, 0.0, 0.0), 2, None),
    ]))
    assert [k for k, _ in res[1]] == [3, 7, 11]
    assert [k for k, _ in res[2]] == [1, 3]
    assert res[3] == ()
    assert [k for k, _ in res[4]] == [0, 1]
    gpu.close()


def test_full_size_properties_10m_x_768():
    """BASELINE.json configs[1] at full size (10 M x 768 bf16, Q = 1024, k = 10) through the device tier:
    size-independent properties -- a query that is a stored row finds that row first at distance ~0,
    distances ascend, results are reproducible, slab count does not matter -- plus a torch f64
    brute-force cross-check of 4 queries."""
    import torch

    n, d, q, k, chunk = 10_000_000, 768, 1024, 10, 250_000
    dev = torch.device("cuda", 0)
    free, _total = torch.cuda.mem_get_info(dev)
    if free < 40 * 2 ** 30:
        pytest.skip("needs ~20 GB of free HBM")
    gpu = B200BruteForceKNNIndex(d, n, metric=COS, device=0)
    picks = torch.arange(q, device=dev) * 9_765 + 17  # rows used as queries
    q_rows = torch.empty((q, d), dtype=torch.bfloat16, device=dev)
    for c in range(n // chunk):
        g = torch.Generator(device=dev)
        g.manual_seed(99 + c)
        rows = torch.randn((chunk, d), generator=g, device=dev, dtype=torch.float32).to(torch.bfloat16)
        keys = np.zeros((chunk, 2), dtype=np.uint64)
        keys[:, 0] = np.arange(c * chunk, (c + 1) * chunk, dtype=np.uint64)
        sel = (picks >= c * chunk) & (picks < (c + 1) * chunk)
        q_rows[sel] = rows[picks[sel] - c * chunk]
        torch.cuda.synchronize(dev)
        assert (gpu.add_device(keys, rows) == 0).all()
        del rows
    assert gpu.current_size == n
    out_slots = torch.empty((q, k), dtype=torch.int32, device=dev)
    out_dist = torch.empty((q, k), dtype=torch.float64, device=dev)
    ext = torch.cuda.ExternalStream(gpu.stream_ptr, device=dev)

    def run():
        gpu.search_device(q_rows, k, out_slots, out_dist)
        ext.synchronize()
        return out_slots.cpu().numpy().copy(), out_dist.cpu().numpy().copy()

    s1, d1 = run()
    assert gpu.stat("last_kernel") == 1 and gpu.stat("last_slabs") == 37
    assert (s1[:, 0] == picks.cpu().numpy()).all(), "a stored row must be its own nearest neighbour"
    assert (np.abs(d1[:, 0]) < 1e-12).all()
    assert (np.diff(d1, axis=1) >= 0).all(), "distances ascend"
    assert (d1[:, 1] > 0.5).all()  # random 768-d data: everything else is far
    s2, d2 = run()
    assert (s1 == s2).all() and (d1 == d2).all(), "reproducible (the shared threshold is a pure optimisation)"
    gpu.set_option("slabs", 11)
    s3, d3 = run()
    assert (s1 == s3).all() and (d1 == d3).all(), "slab count must not change results"
    gpu.set_option("slabs", 0)
    # torch f64 brute force for 4 queries
    qq = q_rows[:4].double()
    best_d = torch.full((4, 2 * k), float("inf"), dtype=torch.float64, device=dev)
    best_i = torch.full((4, 2 * k), -1, dtype=torch.int64, device=dev)
    back = gpu.read_rows  # rows come back from the index itself (what is stored is what is searched)
    for c in range(0, n, 500_000):
        x = torch.from_numpy(back(c, 500_000)).to(dev)
        dist = 1.0 - (qq @ x.T) / torch.sqrt((x * x).sum(1)[None, :] * (qq * qq).sum(1)[:, None])
        cd, ci = torch.topk(dist, 2 * k, dim=1, largest=False)
        alld, alli = torch.cat([best_d, cd], 1), torch.cat([best_i, ci + c], 1)
        o = torch.argsort(alld, dim=1, stable=True)[:, : 2 * k]
        best_d, best_i = torch.gather(alld, 1, o), torch.gather(alli, 1, o)
    assert (best_i[:, :k].cpu().numpy() == s1[:4]).all()
    np.testing.assert_allclose(best_d[:, :k].cpu().numpy(), d1[:4], rtol=1e-9, atol=1e-12)
    gpu.close()


# --------------------------------------------------------------------------- exact re-rank store (raw-input parity)
@pytest.mark.parametrize("store,metric", [("f64", COS), ("f64", L2SQ), ("f32", COS), ("f32", L2SQ)])
def test_exact_rerank_store_matches_reference_on_raw_inputs(store, metric):
    """With PWKNN_EXACT_RERANK_F32/F64 the index keeps an exact copy of every row for the f64 re-rank: ids and
    distances equal the oracle's on the RAW (un-rounded) inputs -- f64 data with the f64 store, f32-representable
    data (what an f32 embedding model produces) with the f32 store -- through adds, swap-removes and a regrow."""
    rng = np.random.default_rng(31)
    d = 70
    gpu = B200BruteForceKNNIndex(d, 8, 1024 * 128, metric, exact_rerank=store)
    ora = OracleInner(d, 8, 1024 * 128, metric, impl="c")
    assert gpu.stat("exact_rerank") == (2 if store == "f64" else 1)

    def raw(shape):
        x = rng.standard_normal(shape) * 3.0
        return x if store == "f64" else x.astype(np.float32).astype(np.float64)

    live, nxt = [], 1
    for step in range(6):
        rows = raw((1500, d))
        batch = [(nxt + i, rows[i]) for i in range(1500)]
        nxt += 1500
        rm = [live.pop(int(rng.integers(0, len(live)))) for _ in range(min(len(live), 900))]
        gpu.remove(rm), ora.remove(rm)
        gpu.add(batch), ora.add(batch)
        live += [k for k, _ in batch]
        assert gpu.current_size == ora.current_size and gpu.current_allocated == ora.current_allocated
        qs = [(i, raw(d), int(rng.integers(0, 25))) for i in range(150)]
        got, want = gpu.search(qs), ora.search(qs)
        for (k1, r1), (k2, r2) in zip(got, want):
            assert [m for m, _ in r1] == [m for m, _ in r2], (store, metric, step)
            np.testing.assert_allclose([x for _, x in r1], [x for _, x in r2], rtol=1e-12, atol=1e-13)
    # any-k path (f64 distances of the exact copy + stable sort) and the f32 host entry point
    qs = [(0, raw(d), 400), (1, raw(d), gpu.current_size + 5)]
    got, want = gpu.search(qs), ora.search(qs)
    assert gpu.stat("last_kernel") in (2, 3)  # select path (k <= 4096) or the full sort
    for (_, r1), (_, r2) in zip(got, want):
        assert [m for m, _ in r1] == [m for m, _ in r2]
        np.testing.assert_allclose([x for _, x in r1], [x for _, x in r2], rtol=1e-12, atol=1e-13)
    if store == "f32":
        q32 = rng.standard_normal((64, d)).astype(np.float32)
        ok, sc, cnt, st = gpu.search_arrays(q32, np.full(64, 5, dtype=np.uint32))
        want = ora.search([(i, q32[i].astype(np.float64), 5) for i in range(64)])
        got_ids = [[(int(ok[q * 5 + r, 1]) << 64) | int(ok[q * 5 + r, 0]) for r in range(5)] for q in range(64)]
        assert got_ids == [[m for m, _ in r] for _, r in want]
        np.testing.assert_allclose(sc.reshape(64, 5), [[x for _, x in r] for _, r in want], rtol=1e-12)
    gpu.close()


@pytest.mark.parametrize("name", sorted(kats.CASES))
def test_reference_kat_gpu_exact_distances(name):
    """The reference's KAT inputs (0.1, 0.6 ... are not bf16 numbers): with the f64 re-rank store the replies equal
    the oracle's to the last bits, not only after rounding to 2 decimals."""
    from kat_harness import get_ret, make_list
    from pathway_b200.debug import table_from_markdown
    from pathway_b200.engine import ExternalIndexFactory

    case = kats.CASES[name]
    f = dict(case["factory"])
    metric = BruteForceKnnMetricKind[f.pop("metric")]
    index = table_from_markdown(case["index"], id_from=["pk_source"]).with_columns(data=lambda r: make_list(r["data"]))
    queries = table_from_markdown(case["queries"], id_from=["pk_source"]).with_columns(data=lambda r: make_list(r["data"]))
    factory = ExternalIndexFactory.brute_force_knn_factory(metric=metric, **f)
    gpu = B200BruteForceKNNIndex(f["dimensions"], f["reserved_space"], f["auxiliary_space"], metric, exact_rerank="f64")
    _ret_g, raw_g = get_ret(queries, index, factory, gpu)
    _ret_o, raw_o = get_ret(queries, index, factory, OracleInner(metric=metric, **f))
    for rg, ro in zip(sorted(raw_g.rows, key=lambda r: r.key), sorted(raw_o.rows, key=lambda r: r.key)):
        a, b = rg.values["_pw_index_reply"], ro.values["_pw_index_reply"]
        assert [m for m, _ in a] == [m for m, _ in b]
        np.testing.assert_allclose([x for _, x in a], [x for _, x in b], rtol=1e-13, atol=1e-15)
    gpu.close()


def test_usearch_api_mirror_gpu():
    """The reference's usearch KAT (test_index.py:858-900) and an IP-metric parity run through the B200 index."""
    from test_oracle_golden import _usearch_distances

    made = []

    def inner(d):
        made.append(B200BruteForceKNNIndex(d, 400, metric=1))
        return made[-1]

    _usearch_distances(inner)
    made[0].close()
    rng = np.random.default_rng(4)
    gpu = B200BruteForceKNNIndex(24, 8, metric=2)  # PWKNN_METRIC_IP
    ora = OracleInner(24, 8, 1024 * 128, 2, impl="numpy")
    rows = _rounded(rng, (3000, 24))
    data = [(i, rows[i]) for i in range(3000)]
    gpu.add(data), ora.add(data)
    qs = [(i, _rounded(rng, 24), int(rng.integers(0, 12))) for i in range(150)]
    _check_same(gpu.search(qs), ora.search(qs), "IP")
    gpu.close()


def test_plain_c_caller_of_the_abi():
    """examples/c_abi_demo.c (built by __graft_entry__.build()): remove -> add -> search from C, no Python, no torch."""
    import os
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "examples", "c_abi_demo")
    if not os.path.exists(exe):
        import __graft_entry__

        __graft_entry__.build()
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "c_abi_demo ok" in out.stdout, out.stderr


def test_two_handles_two_threads_and_growth_at_scale():
    """Handles are independent (one per timely worker): two threads drive two indices concurrently; one of them
    grows from reserved_space=1 to 300 k rows through the reference's doubling rule (reallocs in HBM)."""
    import threading

    rng = np.random.default_rng(9)
    d = 64
    rows = bfo.bf16_round(rng.standard_normal((300_000, d)))
    q = bfo.bf16_round(rng.standard_normal((64, d)))
    results = {}

    def worker(name, reserved, chunk):
        gpu = B200BruteForceKNNIndex(d, reserved, metric=COS)
        for c0 in range(0, len(rows), chunk):
            keys = np.zeros((min(chunk, len(rows) - c0), 2), dtype=np.uint64)
            keys[:, 0] = np.arange(c0, c0 + len(keys), dtype=np.uint64)
            assert (gpu.add_arrays(keys, rows[c0:c0 + len(keys)]) == 0).all()
        ok, sc, cnt, st = gpu.search_arrays(q, np.full(64, 10, dtype=np.uint32))
        results[name] = (ok[:, 0].copy(), sc.copy(), gpu.current_allocated)
        gpu.close()

    t1 = threading.Thread(target=worker, args=("grow", 1, 7000))
    t2 = threading.Thread(target=worker, args=("fixed", 300_000, 50_000))
    t1.start(); t2.start(); t1.join(); t2.join()
    assert (results["grow"][0] == results["fixed"][0]).all() and (results["grow"][1] == results["fixed"][1]).all()
    assert results["fixed"][2] == 300_000 and results["grow"][2] >= 300_000
    # spot-check against torch on the same (bf16-exact) data
    import torch

    x = torch.tensor(rows, dtype=torch.float64, device="cuda")
    qq = torch.tensor(q, dtype=torch.float64, device="cuda")
    dist = 1.0 - (qq @ x.T) / torch.sqrt((x * x).sum(1)[None, :] * (qq * qq).sum(1)[:, None])
    top = torch.topk(dist, 10, dim=1, largest=False)
    assert (top.indices.cpu().numpy().reshape(-1) == results["fixed"][0].astype(np.int64)).all()
    np.testing.assert_allclose(-results["fixed"][1], top.values.cpu().numpy().reshape(-1), rtol=1e-9, atol=1e-12)


def test_config1_shape_static_table_k3_gpu():
    """BASELINE.json configs[0] shape (1 k x 64, k = 3, 100 queries, static table) through the B200 index with the
    f64 re-rank store: identical to the CPU reference path on the raw random embeddings."""
    rng = np.random.default_rng(0)
    rows, qs = rng.standard_normal((1000, 64)), rng.standard_normal((100, 64))
    for metric in (COS, L2SQ):
        gpu = B200BruteForceKNNIndex(64, 400, 1024 * 128, metric, exact_rerank="f64")
        ora = OracleInner(64, 400, 1024 * 128, metric, impl="c")
        data = [(i + 1, rows[i]) for i in range(1000)]
        gpu.add(data), ora.add(data)
        q = [(i, qs[i], 3) for i in range(100)]
        for (_, r1), (_, r2) in zip(gpu.search(q), ora.search(q)):
            assert [m for m, _ in r1] == [m for m, _ in r2] and len(r1) == 3
            np.testing.assert_allclose([x for _, x in r1], [x for _, x in r2], rtol=1e-12, atol=1e-14)
        gpu.close()


def test_device_tier_on_a_caller_stream():
    """`stream` arguments of the device tier are plain cudaStream_t handles: a torch-created stream works, the
    work is ordered on it, and the result equals the run on the index's own stream."""
    import torch

    rng = np.random.default_rng(12)
    dev = torch.device("cuda", 0)
    gpu = B200BruteForceKNNIndex(96, 20000, metric=COS, device=0)
    rows = torch.tensor(_rounded(rng, (20000, 96)), dtype=torch.float32, device=dev).to(torch.bfloat16)
    keys = np.zeros((20000, 2), dtype=np.uint64); keys[:, 0] = np.arange(20000, dtype=np.uint64)
    side = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(side):
        rows2 = rows.clone()  # produced on the side stream; add_device waits for that stream
        st = gpu.add_device(keys, rows2, stream_ptr=side.cuda_stream)
    assert (st == 0).all()
    q = torch.tensor(_rounded(rng, (300, 96)), dtype=torch.float32, device=dev).to(torch.bfloat16)
    a_s = torch.empty((300, 10), dtype=torch.int32, device=dev); a_d = torch.empty((300, 10), dtype=torch.float64, device=dev)
    b_s = torch.empty_like(a_s); b_d = torch.empty_like(a_d)
    torch.cuda.synchronize(dev)
    gpu.search_device(q, 10, a_s, a_d)  # own stream
    torch.cuda.ExternalStream(gpu.stream_ptr, device=dev).synchronize()
    with torch.cuda.stream(side):
        gpu.search_device(q, 10, b_s, b_d, stream_ptr=side.cuda_stream)
        total = b_s.to(torch.int64).sum()  # consumer on the same stream, no explicit sync in between
    side.synchronize()
    assert (a_s == b_s).all() and (a_d == b_d).all() and int(total.item()) == int(a_s.to(torch.int64).sum().item())
    gpu.close()


def test_filter_pushdown_equals_reference_k_doubling_loop():
    """Filter push-down (row bitmasks in the kernel epilogue, one pass) against the reference's post-filter +
    k-doubling loop (DerivedFilteredSearchIndex over the CPU oracle): same replies through adds, swap-removes,
    selective and empty filters, non-boolean / malformed expressions and filter-cache eviction."""
    from pathway_b200.external_index import AddDataEntry, DerivedFilteredSearchIndex, QueryEntry

    rng = np.random.default_rng(21)
    d = 24
    gpu = B200BruteForceKNNIndex(d, 8, metric=COS)
    push = DerivedFilteredSearchIndex(gpu)  # push-down on (the B200 index offers filter_create)
    assert push.pushdown
    push.MAX_PUSHED_FILTERS = 3
    loop = DerivedFilteredSearchIndex(OracleInner(d, 8, 1024 * 128, COS, impl="c"), pushdown=False)

    def meta(i):
        return {"path": f"dir{i % 7}/sub{i % 3}/f{i}.{'py' if i % 4 else 'md'}", "owner": f"u{i % 100}", "n": int(i % 13)}

    live, nxt = [], 0
    filters = ["n == `3`", "globmatch('dir2/*/*.py', path)", "owner == 'u17'", "n > `5` && contains(path, 'sub1')",
               "owner == 'nobody'", None, "n", "n ==", "n < `100`"]
    for step in range(5):
        rows = _rounded(rng, (2500, d))
        batch = [AddDataEntry(nxt + i, tuple(float(x) for x in rows[i]), meta(nxt + i)) for i in range(2500)]
        rm = [live.pop(int(rng.integers(0, len(live)))) for _ in range(min(len(live), 1200))]
        for ix in (push, loop):
            ix.remove(rm)
            ix.add([AddDataEntry(e.key, e.data, e.filter_data) for e in batch])
        live += list(range(nxt, nxt + 2500)); nxt += 2500
        qs = []
        for i in range(60):
            v = tuple(float(x) for x in _rounded(rng, d))
            qs.append(QueryEntry(1000 * step + i, v, int(rng.integers(0, 12)), filters[i % len(filters)]))
        got, want = dict(push.search(qs)), dict(loop.search(qs))
        assert set(got) == set(want)
        for k in want:
            if isinstance(want[k], Exception):
                assert isinstance(got[k], Exception), (step, k)
                continue
            assert [m for m, _ in got[k]] == [m for m, _ in want[k]], (step, k, got[k][:3], want[k][:3])
            np.testing.assert_allclose([x for _, x in got[k]], [x for _, x in want[k]], rtol=1e-9, atol=1e-12)
    assert 0 < len(push._pushed) <= 3  # LRU-capped cache of pushed expressions (the overflow used the lazy loop)
    (kq, rq), = push.search([QueryEntry(5, tuple(float(x) for x in _rounded(rng, d)), 5, "n == `3`")])
    assert gpu.stat("last_kernel") == 1 and len(rq) == 5  # a pushed filter is one pass of the fused kernel
    # large limit on a selective filter: the any-k path must honour the mask too
    q = QueryEntry(1, tuple(float(x) for x in _rounded(rng, d)), 400, "owner == 'u17'")
    (k1, r1), = push.search([q]); (k2, r2), = loop.search([q])
    assert [m for m, _ in r1] == [m for m, _ in r2] and 0 < len(r1) < 400
    gpu.close()


@pytest.mark.parametrize("store", [None, "f32"])
def test_snapshot_roundtrip(tmp_path, store):
    """pwknn_snapshot_save / _load: the restored index answers exactly like the original (same slots, sizes,
    allocation rule state) and keeps accepting deltas."""
    rng = np.random.default_rng(44)
    d = 50
    gpu = B200BruteForceKNNIndex(d, 8, metric=L2SQ, exact_rerank=store)
    ora = OracleInner(d, 8, 1024 * 128, L2SQ, impl="c")
    prep = (lambda x: x.astype(np.float32).astype(np.float64)) if store else bfo.bf16_round
    rows = prep(rng.standard_normal((5000, d)))
    data = [(i * 3 + 1, rows[i]) for i in range(5000)]
    gpu.add(data), ora.add(data)
    rm = [k for k, _ in data[100:1900]]
    gpu.remove(rm), ora.remove(rm)
    path = tmp_path / "index.pwknn"
    gpu.snapshot_save(path)
    qs = [(i, prep(rng.standard_normal(d)), 9) for i in range(80)]
    want = gpu.search(qs)
    size, alloc = gpu.current_size, gpu.current_allocated
    gpu.close()
    back = B200BruteForceKNNIndex.snapshot_load(path)
    assert back.current_size == size == ora.current_size and back.current_allocated == alloc
    got = back.search(qs)
    assert got == want
    _check_same(got, ora.search(qs), "snapshot")
    more = prep(rng.standard_normal((700, d)))
    batch = [(100000 + i, more[i]) for i in range(700)]
    back.add(batch), ora.add(batch)
    back.remove([4, 7]), ora.remove([4, 7])
    _check_same(back.search(qs), ora.search(qs), "snapshot + deltas")
    assert back.current_allocated == ora.current_allocated
    back.close()


def test_soak_constant_memory():
    """Reference soak (integration_tests/stdlib/test_index.py:48-60): 400 k queries through a small index must not
    grow memory.  Here: 1600 batches x 256 queries with interleaved upserts / removes; host RSS and device memory
    stay flat after warm-up."""
    import resource

    import torch

    rng = np.random.default_rng(8)
    d = 64
    gpu = B200BruteForceKNNIndex(d, 16, metric=COS)
    rows = rng.standard_normal((6, d))
    gpu.add([(i + 1, rows[i]) for i in range(6)])
    q = rng.standard_normal((256, d))
    lim = np.full(256, 3, dtype=np.uint32)

    def cycle(i):
        gpu.search_arrays(q, lim)
        if i % 8 == 0:
            k = 1000 + (i // 8) % 50
            gpu.add([(k, rng.standard_normal(d))])
            if i % 16 == 0:
                gpu.remove([k])

    for i in range(200):
        cycle(i)
    rss0 = resource.getrusage(resource.RUSAGE_SELF).ru_maxrss
    free0, _ = torch.cuda.mem_get_info(0)
    for i in range(200, 1800):
        cycle(i)
    rss1 = resource.getrusage(resource.RUSAGE_SELF).ru_maxrss
    free1, _ = torch.cuda.mem_get_info(0)
    assert rss1 - rss0 < 64 * 1024, f"host RSS grew by {(rss1 - rss0) / 1024:.1f} MB over 409 600 queries"
    assert free0 - free1 < 64 * 2 ** 20, f"device memory grew by {(free0 - free1) / 2 ** 20:.1f} MB"
    gpu.close()
