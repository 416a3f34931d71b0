/*
 * pwknn.h -- C ABI of the B200-native brute-force KNN index (libpwknn.so).
 *
 * This is the drop-in boundary for Pathway's incremental nearest-neighbour
 * operator.  It sits exactly below the Rust trait
 *
 *   trait NonFilteringExternalIndex<Vec<f64>, Vec<f64>>          (reference:
 *       fn add(&mut self, Vec<(Key, Vec<f64>)>) -> Vec<(Key, DynResult<()>)>
 *       fn remove(&mut self, Vec<Key>)          -> Vec<(Key, DynResult<()>)>
 *       fn search(&self, &[(Key, Vec<f64>, usize)])
 *                 -> Vec<(Key, DynResult<Vec<KeyScoreMatch>>)>
 *   )  src/external_integration/mod.rs:375-382
 *
 * and replaces `BruteForceKNNIndex`
 *   (src/external_integration/brute_force_knn_integration.rs:22-239)
 * plus its `KeyToU64IdMapper` (mod.rs:209-256).  The Rust-side binding a
 * maintainer adds is shown in INTEGRATION.md.
 *
 * Conventions
 *  - plain C, no CUDA/torch types in any signature; `void*` device pointers
 *    are raw CUDA device addresses, `stream` is a `cudaStream_t` cast to void*
 *    (NULL = the index's own stream).
 *  - stream ordering: the index's own stream is non-blocking, i.e. NOT implicitly ordered against the legacy default
 *    stream.  A call with stream == NULL therefore first makes the index stream wait for everything already queued on
 *    the legacy default stream (where the producer of d_rows / d_queries usually ran); a producer on any OTHER stream
 *    must be passed as `stream` or synchronised by the caller.  A call with its own `stream` is ordered after the
 *    deltas still draining on the index stream.  Device-tier searches are asynchronous: do not mutate the index
 *    (add / remove) before a search running on a stream other than the index's own has completed.
 *  - pwknn_add / pwknn_add_f32 / pwknn_remove return as soon as their inputs sit in the library's pinned staging
 *    ring; the H2D copies and kernels drain in stream order BEFORE any later search of the same handle (the
 *    remove -> add -> search contract of a timestamp holds without a host wait).  A CUDA failure of a drained delta
 *    surfaces as the error of the next call.
 *  - `Key` (u128, src/engine/value.rs:34,41) is passed as two u64: {lo, hi}.
 *  - the caller owns every buffer; inputs are consumed before the call
 *    returns, outputs are complete when it returns (the operator emits them
 *    immediately: src/engine/dataflow/operators/external_index.rs:145-151).
 *  - never aborts the process.  Function result: PWKNN_OK or a PWKNN_E_* code
 *    for a whole-call failure (text via pwknn_last_error).  Per-row results go
 *    to `status[i]`:
 *        0 PWKNN_ROW_OK
 *        1 PWKNN_ROW_MISSING_KEY   -> DataError::MissingKey(key)  mod.rs:248-252
 *        2 PWKNN_ROW_BAD_VECTOR    -> dimension / limit error (the reference has
 *                                     no check here; see DESIGN.md "quirks")
 *        3 PWKNN_ROW_DEVICE_ERROR
 *  - threading: one handle per timely worker, used from one thread at a time
 *    (external_index.rs:118-153); different handles may be used concurrently.
 *  - call order per timestamp: remove(all -diff) -> add(all +diff) -> search
 *    (mod.rs:144-150).
 */
#ifndef PWKNN_H
#define PWKNN_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PWKNN_OK 0
#define PWKNN_E_INVALID 1   /* bad argument */
#define PWKNN_E_CUDA 2      /* CUDA runtime/driver error (see pwknn_last_error) */
#define PWKNN_E_NOGPU 3     /* no usable sm_100 device / driver */
#define PWKNN_E_OOM 4       /* device or pinned allocation failed */

#define PWKNN_ROW_OK 0
#define PWKNN_ROW_MISSING_KEY 1
#define PWKNN_ROW_BAD_VECTOR 2
#define PWKNN_ROW_DEVICE_ERROR 3

/* BruteForceKnnMetricKind  (brute_force_knn_integration.rs:16-20) */
#define PWKNN_METRIC_L2SQ 0
#define PWKNN_METRIC_COS 1
/* usearch MetricKind::IP (distance = 1 - x.q), for the USearchKnn API mirror: usearch_integration.rs:34-40 */
#define PWKNN_METRIC_IP 2

typedef struct pwknn_index pwknn_index;

/* ABI version of this header (bumped on any signature change). */
uint32_t pwknn_abi_version(void);

/* Thread-local text of the last failure of a call that had no handle
 * (pwknn_create), else use pwknn_last_error(h). */
const char* pwknn_last_global_error(void);

/* BruteForceKNNIndex::new(dimensions, reserved_space, auxiliary_space, metric)
 * (brute_force_knn_integration.rs:34-51).  `auxiliary_space` is accepted for
 * signature parity; it only chunks queries in the reference (:205) and never
 * changes results (test_brute_force_knn.py:63-70), so it is ignored here.
 * `device` is the CUDA ordinal holding this index (one shard). */
int pwknn_create(uint32_t dimensions, uint64_t reserved_space, uint64_t auxiliary_space,
                 int metric, int device, pwknn_index** out);

/* Same, with flags.  PWKNN_EXACT_RERANK_F32 / _F64 additionally keep an exact f32 / f64 copy of every row in HBM
 * (4 / 8 bytes per element; 10 M x 768 f32 = 30.7 GB).  The tensor-core pass still ranks candidates on the bf16
 * matrix; the final f64 re-rank then reads the exact copy and the caller's ORIGINAL query vectors, so ids and
 * distances are those of the reference on the raw inputs (not only on bf16-pre-rounded ones).  f32 is lossless
 * for embeddings that come from f32 models (python_api.rs:418-428 widens f32 arrays to f64).
 * pwknn_create uses the environment variable PWKNN_EXACT_RERANK = f32 | f64 for the flags. */
#define PWKNN_EXACT_RERANK_F32 1u
#define PWKNN_EXACT_RERANK_F64 2u
int pwknn_create_ex(uint32_t dimensions, uint64_t reserved_space, uint64_t auxiliary_space,
                    int metric, int device, uint32_t flags, pwknn_index** out);

/* One handle, several GPUs of this process (SURVEY.md 8b "one instance per timely worker" + 8e): the same trait
 * surface -- pwknn_add / _add_f32 / _remove / _search / _search_f32 / _search_filtered / _filter_* / _size /
 * _allocated / _set_option / _get_stat -- over rows spread across `devices` (a key lives on the device that owns its
 * hash range, equal ranges unless pwknn_multi_set_weights says otherwise; each device keeps the reference's
 * dense-slot bookkeeping for its rows).  A search stages
 * the queries once, runs on every device concurrently and merges the per-device replies on the host ordered by
 * (distance, device position in `devices`, slot); within a device that is the reference's (distance, slot) order,
 * across devices exact distance ties fall back to the device order (the reference's slot order is history-dependent
 * and not shardable).  A device may be listed more than once (several shards on one GPU; used by the single-GPU
 * tests).  The device-pointer entry points, pwknn_read_rows and pwknn_slots_to_keys are per-device and return
 * PWKNN_E_INVALID on such a handle (reach them through pwknn_shard); snapshots: pwknn_snapshot_load_multi.  The Rust shim needs no change beyond calling this constructor: a single
 * worker (the reference's default, PATHWAY_THREADS=1) then drives all GPUs. */
int pwknn_create_multi(uint32_t dimensions, uint64_t reserved_space, uint64_t auxiliary_space, int metric,
                       const int* devices, uint32_t n_devices, uint32_t flags, pwknn_index** out);

/* Placement weights of a multi-device handle: device g receives the share weights[g] / sum(weights) of the key-hash
 * space (default: equal shares).  A search costs the slowest device, and the GPUs of one box run at persistently
 * different speeds under the per-GPU power cap, so weights proportional to measured speed shorten the step
 * (pathway_b200/placement.py measures them).  Only on an EMPTY index: a key's device is a pure function of the key.
 * The weights are part of a snapshot. */
int pwknn_multi_set_weights(pwknn_index* h, uint32_t n_devices, const double* weights);

/* The i-th shard of a multi-device handle: a single-device handle owned by `h` (never destroy it) on which the
 * per-device entry points work -- bulk loads with pwknn_add_device_bf16, pwknn_read_rows, statistics.  Rows added
 * straight to a shard must belong to it: pwknn_shards_of_keys tells where the router would put each key
 * (all 0 for a single-device handle).  NULL for a single-device handle or i out of range. */
pwknn_index* pwknn_shard(pwknn_index* h, uint32_t i);
int pwknn_shards_of_keys(const pwknn_index* h, uint64_t n, const uint64_t* keys, uint32_t* out_shards);

/* Drop -> free HBM, pinned staging, the stream. */
int pwknn_destroy(pwknn_index* h);

/* NonFilteringExternalIndex::add  (brute_force_knn_integration.rs:114-144).
 * keys: 2*n u64 {lo,hi}; rows: n x dimensions row-major f64 (host).  Rows are
 * rounded f64 -> f32 -> bf16 (RNE twice) into the HBM corpus. A key that is
 * already live overwrites its row in place (reference quirk fenced off, see
 * DESIGN.md).  status may be NULL. */
int pwknn_add(pwknn_index* h, uint64_t n, const uint64_t* keys, const double* rows, int32_t* status);
/* Same, rows as f32 (host): the "next" ingest path of SURVEY.md 8(f)-1. */
int pwknn_add_f32(pwknn_index* h, uint64_t n, const uint64_t* keys, const float* rows, int32_t* status);
/* Same, rows already bf16 and resident in HBM on this index's device
 * (row_stride_elems >= dimensions).  Used for bulk loads and by bench.py. */
int pwknn_add_device_bf16(pwknn_index* h, uint64_t n, const uint64_t* keys, const void* d_rows,
                          uint64_t row_stride_elems, void* stream, int32_t* status);

/* NonFilteringExternalIndex::remove  (brute_force_knn_integration.rs:146-187):
 * swap-with-last per key in order, then the halve-when-quarter-full rule. */
int pwknn_remove(pwknn_index* h, uint64_t n, const uint64_t* keys, int32_t* status);

/* NonFilteringExternalIndex::search  (brute_force_knn_integration.rs:189-238).
 * Every query is answered by the tier its OWN limit needs (a single large limit does not slow the others down):
 *   tier 1  k + margin <= 108   fused tcgen05 kernel (per-slab candidate lists) + f64 re-rank, followed by a CERTAINTY
 *                               TEST: could a row outside the re-ranked candidates still belong to the top k, given the
 *                               noise bound of the fp32 (and, with an exact re-rank store, bf16-rounded) ranking pass?
 *                               Queries that fail it (near-duplicate embeddings, exact ties at the k-th place) are
 *                               re-answered by tier 2 before the call returns;
 *   tier 2  k <= 4096           score dump + radix threshold + f64 re-rank of every row within 2 eps of the k-th score;
 *   tier 3  any k               f64 distances of all rows + stable sort (also takes tier-2 queries with > 8192 near-ties).
 * queries: nq x dimensions f64 (host); limits: nq (k per query).
 * Outputs are packed per query at offset sum(limits[0..q)):
 *   out_keys   2*sum(limits) u64 {lo,hi}   KeyScoreMatch.key
 *   out_scores   sum(limits) f64           KeyScoreMatch.score = -distance (:230)
 *   out_counts   nq u32                    matches written for query q
 *                                          (= min(limit, size); 0 on empty index, :193-198)
 * ascending (distance, slot) order inside each query (:220-224). */
int pwknn_search(pwknn_index* h, uint64_t nq, const double* queries, const uint32_t* limits,
                 uint64_t* out_keys, double* out_scores, uint32_t* out_counts, int32_t* status);
int pwknn_search_f32(pwknn_index* h, uint64_t nq, const float* queries, const uint32_t* limits,
                     uint64_t* out_keys, double* out_scores, uint32_t* out_counts, int32_t* status);

/* Filter push-down.  The reference filters AFTER the search and re-searches with k, 2k, 4k, ...
 * until enough candidates survive (DerivedFilteredSearchIndex, mod.rs:457-504,583-601).  Here a filter is a per-row
 * boolean the library keeps next to the rows (it follows swap-removes; new rows start excluded): the host evaluates
 * its predicate once per row (pwknn_filter_set) and a filtered query is answered in ONE pass -- rows whose bit is
 * clear are skipped inside the kernel epilogue.  filter_ids: nq entries, PWKNN_NO_FILTER for unfiltered queries.
 * out_counts[q] = min(limit, number of passing rows). */
#define PWKNN_NO_FILTER 0xFFFFFFFFu
int pwknn_filter_create(pwknn_index* h, uint32_t* out_filter_id);
int pwknn_filter_destroy(pwknn_index* h, uint32_t filter_id);
int pwknn_filter_set(pwknn_index* h, uint32_t filter_id, uint64_t n, const uint64_t* keys, const uint8_t* pass,
                     int32_t* status);
int pwknn_search_filtered(pwknn_index* h, uint64_t nq, const double* queries, const uint32_t* limits,
                          const uint32_t* filter_ids, uint64_t* out_keys, double* out_scores, uint32_t* out_counts,
                          int32_t* status);

/* Device-resident tier (what bench.py's `value` times, and what the sharded
 * multi-GPU host uses before its all-gather).  d_queries: nq x dimensions bf16
 * in HBM (row_stride_elems >= dimensions); one k for the whole batch.
 * d_out_slots: nq*k u32, d_out_dist: nq*k f64 in HBM; rows beyond
 * min(k,size) are filled with slot 0xFFFFFFFF / +inf.  Asynchronous on
 * `stream`; returns after enqueueing. */
int pwknn_search_device(pwknn_index* h, uint64_t nq, const void* d_queries, uint64_t row_stride_elems,
                        uint32_t k, void* d_out_slots, void* d_out_dist, void* stream);
/* (k + margin <= 108: only enqueues, and leaves the per-query certainty flags on the device -- read their count with
 * pwknn_get_stat(h, "last_uncertain") and re-ask those queries through pwknn_search if it is not 0.  Larger k runs the
 * select path and WAITS for it, because candidate overflow must be re-answered by the full sort before returning.) */

/* Multi-GPU building blocks (row-range shards, ONE all-gather of the per-shard top-k).
 * pwknn_search_device_packed: same search, but the result is written as nq*k pairs of int64
 *   {global id = slot + id_offset (or -1 = none), IEEE bits of the f64 distance} -- exactly the payload a rank
 *   contributes to the all-gather (NCCL, done by the host: torch.distributed / ncclAllGather).
 * pwknn_merge_shards_device: d_gathered = [world][nq][k][2] int64 as produced by the all-gather; writes the
 *   k best per query ordered by (distance, shard, position), NaN last: d_out_ids nq*k int64, d_out_dist nq*k f64. */
int pwknn_search_device_packed(pwknn_index* h, uint64_t nq, const void* d_queries, uint64_t row_stride_elems,
                               uint32_t k, int64_t id_offset, void* d_out_packed, void* stream);
int pwknn_merge_shards_device(pwknn_index* h, uint32_t world, uint64_t nq, uint32_t k, const void* d_gathered,
                              void* d_out_ids, void* d_out_dist, void* stream);
/* pwknn_search_host_packed: the same per-shard search with HOST f64 queries (what the trait hands over: nq x
 *   dimensions, staged through the pinned buffer exactly like pwknn_search) and the packed result left in HBM, so a
 *   rank's step is  H2D -> kernels -> all-gather -> merge -> ONE D2H.  Returns after enqueueing on `stream`; a query
 *   with a non-finite component fails the whole call (PWKNN_E_INVALID). */
int pwknn_search_host_packed(pwknn_index* h, uint64_t nq, const double* queries, uint32_t k, int64_t id_offset,
                             void* d_out_packed, void* stream);

/* Snapshot.  The reference has no index snapshot: after a restart the index is rebuilt by replaying
 * the input (src/engine/dataflow.rs:2634-2677 wires no persistence wrapper).  These two calls write / restore the
 * complete state (slot->key table, bf16 rows, row constants, the optional exact re-rank store; not the filter
 * masks) so a restart costs one sequential file read instead of re-embedding the corpus. */
int pwknn_snapshot_save(pwknn_index* h, const char* path);
int pwknn_snapshot_load(const char* path, int device, pwknn_index** out);
/* A multi-device handle saves a manifest at `path` plus one ordinary snapshot per shard (`path.shard<i>`); it is
 * restored onto the same NUMBER of devices (keys are placed by hash range per shard), not necessarily the same ones. */
int pwknn_snapshot_load_multi(const char* path, const int* devices, uint32_t n_devices, pwknn_index** out);

/* slot -> key for results of pwknn_search_device (host arrays; slots that are
 * 0xFFFFFFFF or >= size yield key {~0,~0}). */
int pwknn_slots_to_keys(pwknn_index* h, uint64_t n, const uint32_t* slots, uint64_t* out_keys);

uint64_t pwknn_size(const pwknn_index* h);       /* current_size       */
uint64_t pwknn_allocated(const pwknn_index* h);  /* current_allocated (reference growth rule) */
uint32_t pwknn_dimensions(const pwknn_index* h);
int pwknn_device(const pwknn_index* h);
void* pwknn_stream(const pwknn_index* h);        /* the index's cudaStream_t */
const char* pwknn_last_error(const pwknn_index* h);

/* Copy stored (bf16-rounded) rows [slot0, slot0+n) back as f64 n x dimensions (tests). */
int pwknn_read_rows(pwknn_index* h, uint64_t slot0, uint64_t n, double* out_rows);

/* Tuning / diagnostics.  name is one of:
 *   "kernel"      0 = auto (per-query tiers, see pwknn_search), 2 = every query on tier 3 (f64 CUDA-core distances
 *                 + full sort).  stat "last_kernel": 1 fused, 2 full sort, 3 select path; "last_tier" 1..3
 *   "certainty"   1 (default) = certainty test + escalation of uncertain queries; 0 = off (round-1 behaviour)
 *   "noise_sigmas" width of the statistical bf16-rounding bound used with an exact re-rank store (default 8)
 *   "force_tier"  0 = auto, 2 / 3 = send every query there (tests)
 *   "select_budget" bytes of score dump per select-path chunk (default 8 GiB)
 *                 stats "last_uncertain" / "last_overflow": queries of the last search that were escalated from
 *                 tier 1 / tier 2 (device tier: flagged, NOT re-answered -- the caller decides)
 *   "slabs"       corpus slabs per query tile for the tcgen05 path (0 = auto)
 *   "margin"      extra candidates kept by the fp32 pass before the f64 re-rank
 *   "profile"     1 = bracket every fused-kernel launch with CUDA events on its stream
 *                 (read back with stats "fused_ns_sum" / "fused_samples"); setting it resets the samples
 * Returns PWKNN_E_INVALID for unknown names. */
int pwknn_set_option(pwknn_index* h, const char* name, int64_t value);
int64_t pwknn_get_stat(const pwknn_index* h, const char* name);   /* "launches", "last_kernel", "last_slabs", ... */

/* Reciprocal-rank fusion of several retrievers' replies.  Replaces, behind the operator, the
 * reference's dataflow: python/pathway/stdlib/indexing/hybrid_index.py:43-66 (rank r = 1,2,.. inside each reply,
 * score 1/(k_rrf + r)), :75-86 (sum per (query, matched id) over the retrievers), :92-103 (per query, order by
 * descending summed score) and :118-120 (keep the first limits[q]).  Host-only and stateless (no handle, no GPU:
 * a few hundred ids per query).
 *
 * Every list l is laid out exactly like pwknn_search's output: reply q of list l starts at
 * off_l(q) = sum_{p<q} list_limits[l][p] in list_keys[l] (two u64 per key) and holds list_counts[l][q] keys in rank
 * order; list_limits[l] == NULL means packed replies (off_l(q) = sum of counts).  Output uses the same layout with
 * `limits`.  The sum over lists runs in list order; equal fused scores are ordered by ascending key (the reference
 * orders ties by an opaque hash of (query id, matched id): parity unpinned for exact ties). */
int pwknn_rrf_fuse(uint32_t n_lists, uint64_t nq, const uint64_t* const* list_keys, const uint32_t* const* list_counts,
                   const uint32_t* const* list_limits, double k_rrf, const uint32_t* limits, uint64_t* out_keys,
                   double* out_scores, uint32_t* out_counts);

/* Debug: raw fp32 ranking scores of the tcgen05 pass for nq x size (HBM out, row-major
 * nq x ld_out f32) -- dot*inv_norm (COS) or dot-0.5*|x|^2 (L2SQ).  Test use only. */
int pwknn_debug_scores_device(pwknn_index* h, uint64_t nq, const void* d_queries, uint64_t row_stride_elems,
                              void* d_out_scores, uint64_t ld_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PWKNN_H */
