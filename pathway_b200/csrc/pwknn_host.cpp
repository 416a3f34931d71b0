// Host side of libpwknn: the C ABI of include/pwknn.h.
//
// Mirrors, as C++ behind extern "C", the Rust struct BruteForceKNNIndex
//   (/root/reference/src/external_integration/brute_force_knn_integration.rs:22-239)
// and its KeyToU64IdMapper (src/external_integration/mod.rs:209-256): dense slots 0..size,
// append on add, swap-with-last on remove, the grow (x2) / shrink (/2 when < 1/4 full) rule.
// The slot map lives on the host (integer bookkeeping, O(delta)); rows live in HBM as bf16 and
// every arithmetic step runs in the CUDA kernels.  There is NO CPU compute fallback: without a
// usable sm_100 device pwknn_create fails with PWKNN_E_NOGPU.
//
// Host-only building blocks live in their own CPU-tested headers: key_slot_map.h (sharded open-addressing key -> slot
// table), slot_book.h (batched swap-remove), host_pool.h (worker pool).
//
// Layout of this file, top to bottom:
//   DevBuf / PinBuf, struct pwknn_index                        state of one handle (incl. SearchState, staging ring)
//   fail / CU_TRY / ensure_dev / ensure_pin / realloc_corpus   error and memory helpers
//   plan_fused / make_plan                                     kernel planning of the fused tcgen05 pass
//   SubBatch, noise_bound, tier_of, run_fused / run_select / run_full
//                                                              the three search tiers on one (sub-)batch of queries
//   stream_copy_check / stream_narrow_check                    host staging (non-temporal copies + usability test)
//   ensure_ring / ring_*, add_impl, prepare_filters            delta path (pinned staging ring, single-pass add), filter masks
//   search_stage / enqueue_round / search_begin / search_finish / search_impl
//                                                              pwknn_search: classes by k, rounds, escalation
//   shard_of / multi_add / multi_remove / multi_search         multi-device handle (router over per-device handles)
//   extern "C": create / destroy / add / remove / search / filters / device tier / snapshots / rrf_fuse /
//               introspection / options and stats               (every entry point that can allocate is `guarded`)
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <emmintrin.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <condition_variable>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <numeric>
#include <thread>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/pwknn.h"
#include "host_pool.h"
#include "key_slot_map.h"
#include "slot_book.h"
#include "pwknn_internal.h"

using namespace pwknn;

namespace {

thread_local std::string g_last_error;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t cap = 0;  // elements
};
template <typename T>
struct PinBuf {
    T* p = nullptr;
    size_t cap = 0;
};

}  // namespace

struct pwknn_index {
    int device = 0;
    uint32_t dims = 0, dpad = 0;
    int metric = PWKNN_METRIC_COS;
    // --- BruteForceKNNIndex fields (rs:22-31)
    uint64_t current_size = 0, current_allocated = 0, minimum_allocated = 0, auxiliary_space = 0;
    // --- KeyToU64IdMapper (mod.rs:209-256): slots are dense, so id->key is a vector
    ShardedKeyMap key_to_slot;
    std::vector<Key> slot_to_key;
    // per-batch scratch indexed by slot ("stamp" arrays instead of per-batch hash maps)
    std::vector<SlotStamp> stamp;  // one record per slot and delta batch (slot_book.h): add -> {gen, a = batch row}; remove -> swap replay
    uint32_t generation = 0;
    // --- HBM corpus
    __nv_bfloat16* d_rows = nullptr;
    RowAux* d_aux = nullptr;
    uint64_t phys_rows = 0;
    // filter push-down: per-filter row bitmask (bit = row passes), kept in slot order on the host and mirrored in HBM
    struct Filter { std::vector<uint32_t> bits; bool live = false; bool dirty = false; };
    std::vector<Filter> filters;
    DevBuf<uint32_t> d_masks;     // [filters.size()][mask_words]
    uint32_t mask_words = 0;      // words per filter in d_masks (= phys_rows / 32 at upload time)
    DevBuf<uint32_t> d_qfilter;   // [qpad] filter index per query of the current batch
    PinBuf<uint32_t> h_qfilter;
    ShadowStore shadow;          // optional exact (f32 / f64) copy of the corpus used only by the f64 re-rank
    CUtensorMap tmap_corpus{};       // box [256 x 64]: one CTA loads the whole corpus tile
    CUtensorMap tmap_corpus_half{};  // box [128 x 64]: each CTA of a pair loads half of it
    // --- query side
    DevBuf<__nv_bfloat16> d_q;   // [qpad_cap x dpad]
    uint64_t qpad_cap = 0;
    CUtensorMap tmap_q{};
    // --- scratch
    DevBuf<uint8_t> d_stage;     // H2D staging of delta rows / queries
    DevBuf<uint32_t> d_slots;    // slots of a delta chunk / move pairs
    DevBuf<float> d_gthr;        // [qpad_cap] cross-slab thresholds
    DevBuf<uint32_t> d_progress; // [slabs][groups] sibling progress counters of the fused kernel
    DevBuf<uint32_t> d_flags;    // optimistic two-pass: [nq] incomplete + [groups] flags
    DevBuf<float> d_part_val;
    DevBuf<uint32_t> d_part_idx;
    DevBuf<uint32_t> d_out_slots;
    DevBuf<double> d_out_dist;
    DevBuf<uint8_t> d_exact_ws;
    PinBuf<uint8_t> h_stage;
    PinBuf<uint32_t> h_slots;
    PinBuf<uint32_t> h_out_slots;
    PinBuf<double> h_out_dist;
    // --- per-query routing (search_begin / search_finish): queries of one call are split into classes by k
    //   tier 1  k + margin <= 108          fused tcgen05 kernel with per-slab candidate lists, certainty-tested
    //   tier 2  k <= kSelectMaxK           score dump + radix threshold + f64 re-rank (also re-answers uncertain queries)
    //   tier 3  anything else              f64 distances of every row + full sort
    struct QueryClass {
        int tier = 1;
        uint32_t kmax = 0;                 // result rows per query in this class's output region
        std::vector<uint32_t> idx;         // batch positions
        uint64_t out_base = 0, flag_base = 0, meta_base = 0;
    };
    struct SearchState {
        std::vector<QueryClass> round, later;   // classes in flight / deferred to a later round (output budget)
        std::vector<uint32_t> kq;               // per batch position: min(limit, size); 0 = nothing to do
        std::vector<uint32_t> filter_ids;       // per batch position (empty: no filters)
        const void* src_p = nullptr; uint32_t src_kind = 0; uint64_t src_stride = 0;
        FilterView masks;                       // masks / words of this batch (q_filter is per class)
        // results, per batch position q: rows [res_off[q], res_off[q] + res_cnt[q])
        std::vector<uint64_t> res_off;
        std::vector<uint32_t> res_cnt, res_slots;
        std::vector<double> res_dist;
        int64_t escalated_uncertain = 0, escalated_overflow = 0;
    } ss;
    DevBuf<uint32_t> d_meta;      // per class: q_index[nq], q_k[nq], q_filter[qpad]
    PinBuf<uint32_t> h_meta;
    DevBuf<uint32_t> d_cflags;    // per class: uncertain / overflow flags
    PinBuf<uint32_t> h_cflags;
    DevBuf<float> d_scores;       // select path: score dump [q_chunk][phys_rows]
    DevBuf<uint8_t> d_select_ws;
    // --- delta staging ring (add / remove): kRingSlots pinned + device buffers; a job = fill the pinned slot (slot
    // numbers + rows, or (src, dst) move pairs), ONE H2D copy on the copy stream, one kernel on the index stream.
    // The calls return once their inputs sit in the ring: copies and kernels drain behind them (stream order keeps
    // remove -> add -> search correct), so staging of the next call overlaps the DMA / kernels of this one.
    static constexpr int kRingSlots = 8;
    struct RingSlot { uint8_t* h = nullptr; uint8_t* d = nullptr; cudaEvent_t copied = nullptr, done = nullptr; bool used = false; };
    RingSlot ring[kRingSlots];
    size_t ring_bytes = 0;
    int ring_next = 0;
    cudaStream_t copy_stream = nullptr;
    float* d_xmax_sq = nullptr;   // largest |x|^2 ever stored (noise bound of the ranking pass, L2SQ / IP)
    cudaEvent_t order_ev = nullptr;  // orders the index stream after the legacy default stream (stream == NULL calls)
    cudaStream_t stream = nullptr;
    cudaEvent_t stage_ev[2] = {nullptr, nullptr};
    bool stage_pending = false;  // pwknn_search_host_packed: an H2D from h_stage may be in flight (stage_ev[0])
    EncodeTiledFn encode = nullptr;
    int num_sms = 0;
    size_t smem_optin = 0;
    // options / stats
    int opt_kernel = 0;  // 0 auto, 1 tcgen05, 2 exact
    int opt_slabs = 0;
    int opt_margin = -1;
    int opt_lead_tiles = 4;  // (A/B: 4 -> 16.2 GB DRAM reads, 12.18 ms; 0 -> 30 GB, 12.71 ms) max tiles a CTA (pair) may run ahead of the slowest started sibling on its slab (0 = off)
    int opt_max_stages = 0;  // diagnostics: cap the smem ring depth (0 = as many as fit)
    int opt_debug_flags = 0; // diagnostics: FusedParams::debug_flags
    int opt_optimistic = 1;  // large k: 16-entry lists + verification + flagged re-run (0 = always full lists)
    int opt_certainty = 1;   // 1 = certainty test + escalation of uncertain queries to the select path (0 = off: round-1 behaviour)
    int opt_noise_sigmas = 8;  // exact re-rank store: width of the statistical bf16-rounding bound
    int64_t opt_select_budget = 8ll << 30;  // bytes of score dump per select chunk
    int opt_force_tier = 0;  // diagnostics / tests: 2 or 3 sends every query to that tier
    int opt_cta_times = 0;   // diagnostics: record %globaltimer at start / end of every CTA of the fused kernel
    DevBuf<unsigned long long> d_cta_times;
    int opt_cta_pair = 0;  // 0 auto, 1 single CTA per tile, 2 CTA pair (cta_group::2)
    int auto_ctas = 2;     // what "auto" picks for >= 2 query tiles (measured: +15% under the power cap); PWKNN_CTA_PAIR=1 reverts
    int64_t stat_launches = 0;
    int64_t stat_last_nq = 0, stat_h2d_bytes = 0;
    int64_t stat_stage_ns = 0, stat_search_host_ns = 0, stat_enqueue_ns = 0, stat_sync_ns = 0;  // last pwknn_search: staging copy / whole call
    // optional device timing of the dominant (fused) kernel: ring of event pairs on the launch stream
    int opt_profile = 0;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_events;
    size_t prof_used = 0;
    int64_t stat_last_ctas = 0, stat_last_two_pass = 0;
    int64_t stat_last_uncertain = 0, stat_last_overflow = 0, stat_last_tier = 0, stat_total_uncertain = 0;
    uint64_t dev_flags_n = 0;  // device tier: flags of the last search live in d_cflags[0, dev_flags_n)
    // host-side time spent inside add / remove since the last reset (set_option "delta_stats", 0), nanoseconds
    int64_t stat_add_check_ns = 0, stat_add_map_ns = 0, stat_add_stage_ns = 0, stat_add_total_ns = 0;
    int64_t stat_rm_map_ns = 0, stat_rm_total_ns = 0, stat_delta_calls = 0;
    int64_t stat_last_kernel = 0, stat_last_slabs = 0, stat_last_kprime = 0, stat_last_stages = 0, stat_last_grid = 0;
    std::string last_error;
    // --- multi-device handle (pwknn_create_multi): this object only routes; rows live in one index per device.
    // current_size / current_allocated above then hold the totals (allocated follows the reference's rule on totals).
    std::vector<pwknn_index*> shards;
    std::vector<uint32_t> shard_bounds;  // shard g owns the 32-bit key hashes in [bounds[g-1], bounds[g]); last = 2^32 - 1 inclusive
    bool multi = false;
    bool is_multi() const { return multi; }
};

namespace {

int fail(pwknn_index* h, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (h) h->last_error = buf;
    g_last_error = buf;
    return code;
}

#define SINGLE_DEVICE_ONLY(h, name) \
    do { if ((h)->is_multi()) return fail((h), PWKNN_E_INVALID, name ": not available on a multi-device handle"); } while (0)

#define CU_TRY(h, expr)                                                                                      \
    do {                                                                                                     \
        cudaError_t _e = (expr);                                                                             \
        if (_e != cudaSuccess) {                                                                             \
            const int _code = (_e == cudaErrorMemoryAllocation) ? PWKNN_E_OOM : PWKNN_E_CUDA;                 \
            return fail(h, _code, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
        }                                                                                                    \
    } while (0)

template <typename T>
int ensure_dev(pwknn_index* h, DevBuf<T>& b, size_t n) {
    if (n <= b.cap) return PWKNN_OK;
    size_t ncap = std::max(n, b.cap + b.cap / 2);
    if (b.p) {
        CU_TRY(h, cudaStreamSynchronize(h->stream));
        CU_TRY(h, cudaFree(b.p));
        b.p = nullptr; b.cap = 0;
    }
    CU_TRY(h, cudaMalloc(reinterpret_cast<void**>(&b.p), ncap * sizeof(T)));
    b.cap = ncap;
    return PWKNN_OK;
}
template <typename T>
int ensure_pin(pwknn_index* h, PinBuf<T>& b, size_t n) {
    if (n <= b.cap) return PWKNN_OK;
    size_t ncap = std::max(n, b.cap + b.cap / 2);
    if (b.p) {
        CU_TRY(h, cudaStreamSynchronize(h->stream));
        CU_TRY(h, cudaFreeHost(b.p));
        b.p = nullptr; b.cap = 0;
    }
    // portable: a multi-device handle lets every device copy from the buffer staged once
    CU_TRY(h, cudaHostAlloc(reinterpret_cast<void**>(&b.p), ncap * sizeof(T), cudaHostAllocPortable));
    b.cap = ncap;
    return PWKNN_OK;
}

uint64_t round_up(uint64_t x, uint64_t a) { return (x + a - 1) / a * a; }

int encode_2d(pwknn_index* h, CUtensorMap* m, void* base, uint64_t rows, uint32_t box_rows) {
    const cuuint64_t gdim[2] = {h->dpad, rows};
    const cuuint64_t gstride[1] = {static_cast<cuuint64_t>(h->dpad) * 2};
    const cuuint32_t box[2] = {kBlockK, box_rows};
    const cuuint32_t estride[2] = {1, 1};
    CUresult r = h->encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, gdim, gstride, box, estride,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(h, PWKNN_E_CUDA, "cuTensorMapEncodeTiled failed: CUresult %d", static_cast<int>(r));
    return PWKNN_OK;
}

// (Re)allocate the HBM corpus for `new_allocated` logical rows; keeps rows [0, keep).
int realloc_corpus(pwknn_index* h, uint64_t new_allocated, uint64_t keep) {
    const uint64_t new_phys = std::max<uint64_t>(round_up(std::max<uint64_t>(new_allocated, 1), kBlockN), kBlockN);
    if (new_phys != h->phys_rows || !h->d_rows) {
        __nv_bfloat16* nrows = nullptr;
        RowAux* naux = nullptr;
        CU_TRY(h, cudaMalloc(reinterpret_cast<void**>(&nrows), new_phys * h->dpad * sizeof(__nv_bfloat16)));
        cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&naux), new_phys * sizeof(RowAux));
        if (e != cudaSuccess) {
            cudaFree(nrows);
            return fail(h, PWKNN_E_OOM, "cudaMalloc(aux) failed: %s", cudaGetErrorString(e));
        }
        void* nshadow = nullptr;
        const size_t sh_row_bytes = static_cast<size_t>(h->shadow.stride) * (h->shadow.kind == 1 ? 4 : 8);
        if (h->shadow.kind) {
            e = cudaMalloc(&nshadow, new_phys * sh_row_bytes);
            if (e != cudaSuccess) {
                cudaFree(nrows); cudaFree(naux);
                return fail(h, PWKNN_E_OOM, "cudaMalloc(exact re-rank store, %zu bytes) failed: %s", new_phys * sh_row_bytes, cudaGetErrorString(e));
            }
        }
        keep = std::min(keep, new_phys);
        if (keep && h->d_rows && nshadow)
            CU_TRY(h, cudaMemcpyAsync(nshadow, h->shadow.rows, keep * sh_row_bytes, cudaMemcpyDeviceToDevice, h->stream));
        if (keep && h->d_rows) {
            CU_TRY(h, cudaMemcpyAsync(nrows, h->d_rows, keep * h->dpad * sizeof(__nv_bfloat16),
                                      cudaMemcpyDeviceToDevice, h->stream));
            CU_TRY(h, cudaMemcpyAsync(naux, h->d_aux, keep * sizeof(RowAux), cudaMemcpyDeviceToDevice, h->stream));
        }
        CU_TRY(h, launch_init_rows(nrows, naux, keep, new_phys, h->dpad, h->stream));
        h->stat_launches++;
        CU_TRY(h, cudaStreamSynchronize(h->stream));
        if (h->d_rows) cudaFree(h->d_rows);
        if (h->d_aux) cudaFree(h->d_aux);
        if (h->shadow.rows) cudaFree(h->shadow.rows);
        h->d_rows = nrows; h->d_aux = naux; h->phys_rows = new_phys; h->shadow.rows = nshadow;
        for (auto& f : h->filters)
            if (f.live) { f.bits.resize(new_phys / 32, 0u); f.dirty = true; }
        int rc = encode_2d(h, &h->tmap_corpus, h->d_rows, h->phys_rows, kBlockN);
        if (rc) return rc;
        rc = encode_2d(h, &h->tmap_corpus_half, h->d_rows, h->phys_rows, kBlockN / 2);
        if (rc) return rc;
    }
    h->current_allocated = new_allocated;
    return PWKNN_OK;
}

int ensure_queries(pwknn_index* h, uint64_t nq) {
    const uint64_t qpad = round_up(std::max<uint64_t>(nq, 1), 2 * kBlockM);  // room for a CTA pair's 2nd tile
    if (qpad > h->qpad_cap) {
        const uint64_t ncap = std::max(qpad, h->qpad_cap * 2);
        int rc = ensure_dev(h, h->d_q, ncap * h->dpad);
        if (rc) return rc;
        rc = ensure_dev(h, h->d_gthr, ncap);
        if (rc) return rc;
        h->qpad_cap = ncap;
        rc = encode_2d(h, &h->tmap_q, h->d_q.p, h->qpad_cap, kBlockM);
        if (rc) return rc;
    }
    return PWKNN_OK;
}

// stream == NULL means the index's own stream, which is created non-blocking: order it after whatever the caller
// already queued on the legacy default stream (where a producer of d_rows / d_queries usually ran).
int order_after_default_stream(pwknn_index* h, void* stream) {
    if (stream) return PWKNN_OK;
    CU_TRY(h, cudaEventRecord(h->order_ev, cudaStreamLegacy));
    CU_TRY(h, cudaStreamWaitEvent(h->stream, h->order_ev, 0));
    return PWKNN_OK;
}

struct FusedPlan {
    uint32_t kprime = 0;  // candidate-list length per (query, slab): 16 / 32 (registers) or 33..108 (smem)
    uint32_t stages = 0, slabs = 0, qtiles = 0, qgroups = 0, ctas = 1, tiles = 0, grid = 0;
    uint64_t qpad = 0;
    bool ok = false;
};
struct Plan {
    int kernel = 2;        // 1 tcgen05 (fused), 2 exact
    uint32_t k_need = 0;   // candidates re-ranked in f64 per query (k + margin, capped by size)
    FusedPlan a;           // the (first) fused pass
    bool two_pass = false; // optimistic: pass a keeps only 16-entry lists; pass b (full k_need lists) reruns flagged groups
    FusedPlan b;
};

uint32_t fused_margin(const pwknn_index* h, uint32_t k);
uint32_t gcd_u32(uint32_t a, uint32_t b) { while (b) { uint32_t t = a % b; a = b; b = t; } return a; }

FusedPlan plan_fused(const pwknn_index* h, uint64_t nq, uint32_t list_len) {
    FusedPlan p;
    const uint32_t n = static_cast<uint32_t>(h->current_size);
    p.kprime = list_len;
    p.qtiles = static_cast<uint32_t>((nq + kBlockM - 1) / kBlockM);
    p.tiles = (n + kBlockN - 1) / kBlockN;
    // CTA pair (cta_group::2): +8..15 % when tensor-bound (less smem/L2 energy per flop under the power cap), but a
    // single query tile is HBM-bound (7.3 vs 4.5 TB/s measured) and an odd tile count wastes half a pair.
    // (measured at 9 tiles: pair 14.6 ms vs single 15.4 ms, so from 9 tiles on the half-empty last pair is worth it)
    const bool pair_pays = p.qtiles >= 2 && (p.qtiles % 2 == 0 || p.qtiles >= 9);
    p.ctas = h->opt_cta_pair == 2 ? 2 : (h->opt_cta_pair == 1 ? 1 : (pair_pays ? h->auto_ctas : 1));
    p.stages = fused_pick_stages(list_len, h->smem_optin, p.ctas);
    if (p.ctas == 2 && p.stages < 2) { p.ctas = 1; p.stages = fused_pick_stages(list_len, h->smem_optin, 1); }
    if (h->opt_max_stages >= 2) p.stages = std::min<uint32_t>(p.stages, h->opt_max_stages);
    p.qgroups = (p.qtiles + p.ctas - 1) / p.ctas;
    p.qpad = static_cast<uint64_t>(p.qgroups) * p.ctas * kBlockM;
    p.ok = p.stages >= 2;
    const uint32_t units = static_cast<uint32_t>(h->num_sms) / p.ctas;  // CTAs or CTA pairs
    uint32_t s = h->opt_slabs > 0 ? static_cast<uint32_t>(h->opt_slabs) : units / gcd_u32(p.qgroups, units);
    // single-CTA mode with an odd tile count (3, 5, 7): units / gcd would mean `qgroups` rounds of short slabs whose
    // sibling CTAs start out of step from the second round on.  One round -- every CTA one (slab, tile) item, all
    // siblings of a slab starting together, units % qgroups SMs idle -- measured better at 10 M rows (interleaved
    // A/B): 3 tiles 5.62 vs 6.24 ms, 5 tiles 8.80 vs 9.70 ms, 7 tiles equal.
    if (h->opt_slabs <= 0 && p.ctas == 1 && p.qgroups > 1 && p.qgroups <= units && gcd_u32(p.qgroups, units) == 1)
        s = units / p.qgroups;
    s = std::min(s, p.tiles);
    s = std::min(s, kMergeMaxCand / list_len);
    s = std::min<uint32_t>(s, 160);  // merge_refine keeps <= 5 list heads per lane
    p.slabs = std::max<uint32_t>(s, 1);
    p.grid = std::min<uint32_t>(units, p.slabs * p.qgroups) * p.ctas;
    return p;
}

Plan make_plan(const pwknn_index* h, uint64_t nq, uint32_t k) {
    Plan p;
    const uint32_t n = static_cast<uint32_t>(h->current_size);
    const uint32_t margin = fused_margin(h, k);
    uint32_t k_need = static_cast<uint32_t>(std::min<uint64_t>(static_cast<uint64_t>(k) + margin, n));
    if (k_need < 1) k_need = 1;
    p.k_need = k_need;
    if (h->opt_kernel == 2 || k_need > kFastMaxKPrime) return p;  // exact path
    if (k_need <= 32) {
        p.a = plan_fused(h, nq, fused_round_kprime(k_need));
        p.k_need = p.a.kprime;  // lists hold 16 / 32 candidates: re-rank them all
    } else {
        // Large k: per-slab lists of k_need entries make the epilogue insert-bound.  Optimistic scheme: pass a keeps
        // only the 16 best per (query, slab) in registers; merge_refine flags a query whose merge exhausted a full
        // list (the only way a needed candidate can be missing); pass b redoes the flagged query groups with full
        // lists.  Worth it when the expected share of a slab in the final k_need is far below 16.
        FusedPlan opt = plan_fused(h, nq, 16);
        const bool feasible = opt.ok && static_cast<uint64_t>(opt.slabs) * 16 >= 2ull * k_need && k_need <= 4 * opt.slabs;
        p.b = plan_fused(h, nq, k_need);
        if (h->opt_optimistic != 0 && feasible && p.b.ok) {
            p.a = opt;
            p.two_pass = true;
        } else {
            p.a = p.b;
        }
    }
    p.kernel = p.a.ok ? 1 : 2;
    return p;
}

void fill_fused_params(const pwknn_index* h, const FusedPlan& p, uint64_t nq, FusedParams* fp) {
    fp->aux = h->d_aux; fp->n_valid = static_cast<uint32_t>(h->current_size); fp->n_q = static_cast<uint32_t>(nq);
    fp->n_qtiles = p.qtiles; fp->n_qgroups = p.qgroups; fp->ctas = p.ctas; fp->n_slabs = p.slabs; fp->n_tiles = p.tiles;
    fp->kprime = p.kprime; fp->k_chunks = h->dpad / kBlockK; fp->stages = p.stages;
    fp->part_val = h->d_part_val.p; fp->part_idx = h->d_part_idx.p; fp->gthr = h->d_gthr.p;
    fp->q_filter = nullptr; fp->filter_masks = nullptr; fp->filter_words = 0;
    fp->group_flags = nullptr; fp->dump = nullptr; fp->dump_ld = 0; fp->debug_flags = static_cast<uint32_t>(h->opt_debug_flags);
}

int launch_fused_pass(pwknn_index* h, const FusedPlan& fpn, FusedParams fp, cudaStream_t s) {
    fp.slab_progress = nullptr; fp.lead_tiles = 0;
    if (h->opt_lead_tiles > 0 && fpn.qgroups > 1) {
        const size_t np = static_cast<size_t>(fpn.slabs) * fpn.qgroups;
        int rc = ensure_dev(h, h->d_progress, np); if (rc) return rc;
        CU_TRY(h, cudaMemsetAsync(h->d_progress.p, 0, np * sizeof(uint32_t), s));
        // Without the shared threshold (first pass of the large-k scheme) the epilogue time per tile varies far more, and a
        // 4-tile leash makes sibling pairs wait on each other: 4x the leash there (k = 100, 2 M rows, Q = 1024,
        // interleaved A/B: 3.35 ms at 4 tiles, 2.79 ms at 16, 2.81 ms without a limit; no difference at Q = 256).
        fp.slab_progress = h->d_progress.p;
        fp.lead_tiles = static_cast<uint32_t>(h->opt_lead_tiles) * (fp.gthr ? 1u : 4u);
        // Short work items (the shard of an 8-GPU index: 1.25 M rows = 132 corpus tiles per item) lose more to waiting on
        // the leash than they gain in L2 hits: 4x the leash there (1.25 M rows, Q = 1024, interleaved A/B on one GPU:
        // 1.73 ms at 4 tiles, 1.65 at 8, 1.63 at 16; no difference from 264 tiles per item on, 14 % WORSE at 528).
        if (fpn.tiles / std::max<uint32_t>(fpn.slabs, 1) < 200) fp.lead_tiles *= 4;
    }
    fp.cta_times = nullptr;
    if (h->opt_cta_times) {
        int rc = ensure_dev(h, h->d_cta_times, 2 * 1024); if (rc) return rc;
        fp.cta_times = h->d_cta_times.p;
    }
    CU_TRY(h, launch_fused_knn(h->tmap_q, fpn.ctas == 2 ? h->tmap_corpus_half : h->tmap_corpus, fp, fpn.grid, s));
    h->stat_launches++;
    return PWKNN_OK;
}

// ---------------------------------------------------------------------------------------------------------
// One (sub-)batch of queries with ONE output width, as the three search tiers see it.
struct QuerySource {          // where the batch's query vectors live on the device
    const void* p = nullptr;
    uint32_t kind = 0;        // 0 bf16, 1 f32, 2 f64
    uint64_t stride = 0;      // elements per row
};
struct SubBatch {
    uint64_t nq = 0;
    uint32_t k = 0;                        // result rows written per query
    QuerySource src;
    const uint32_t* d_qindex = nullptr;    // sub-batch position -> row of src (nullptr = identity)
    const uint32_t* d_qk = nullptr;        // per-query k (certainty test / select rank); nullptr = k
    FilterView fv;                         // q_filter in sub-batch order (padded to a multiple of 256 with kNoFilter)
    uint32_t* d_out_slots = nullptr;       // [nq][k]
    double* d_out_dist = nullptr;
    long long* d_out_packed = nullptr;     // tier 1 only: [nq][k][2] {slot + id_offset | -1, distance bits}
    long long id_offset = 0;
    uint32_t* d_flags = nullptr;           // [nq]: tier 1 "uncertain", tier 2 "candidate overflow" (tier 3 writes 0)
};

size_t elem_bytes(uint32_t kind) { return kind == 0 ? 2 : (kind == 1 ? 4 : 8); }

// Noise bound of the fp32 tensor-core ranking pass (pwknn_internal.h NoiseBound).  c = worst case of the fp32
// accumulation: one rounding of relative size 2^-23 per accumulator update (D/16 MMAs of K = 16) plus slack for the
// epilogue FMA and the rounded per-row constants, relative to |x||q| -- scripts/noise_probe.py measured at most 0.4 c
// (D = 64..1536, Gaussian and heavy-tailed rows).  With an exact re-rank store the pass ranks bf16-ROUNDED copies of
// the raw values: add opt_noise_sigmas x the measured standard deviation 2.35e-3/sqrt(D) of that rounding noise.
NoiseBound noise_bound(const pwknn_index* h) {
    NoiseBound nb;
    const float c = (static_cast<float>(h->dims) / 16.f + 4.f) * 1.1920929e-7f;
    const float sig = h->shadow.kind ? static_cast<float>(h->opt_noise_sigmas) * 2.35e-3f / std::sqrt(static_cast<float>(h->dims)) : 0.f;
    if (h->metric == PWKNN_METRIC_COS) { nb.a0 = c + sig; }
    else if (h->metric == PWKNN_METRIC_IP) { nb.a1 = c + sig; }
    else { nb.a1 = c + 1.1920929e-7f + sig; nb.b = 1.1920929e-7f + sig; }
    nb.xmax_sq = h->d_xmax_sq;
    return nb;
}

// tier of a query that wants k results (k >= 1, already clamped to the index size)
uint32_t fused_margin(const pwknn_index* h, uint32_t k) {
    if (h->opt_margin >= 0) return static_cast<uint32_t>(h->opt_margin);
    // extra candidates the fp32 pass keeps for the f64 re-rank.  fp32 noise is ~1e-7 relative: 6 is generous.  With an
    // exact re-rank store the pass also carries the bf16 rounding noise (~8.5e-5 |q| at D = 768): keep 32-entry lists
    // (measured at 10 M x 768, Q = 1024, interleaved A/B: same kernel time as 16-entry lists).
    if (h->shadow.kind) return k <= 26 ? 32 - k : std::max<uint32_t>(6, k <= 64 ? k / 4 : k / 16);
    return std::max<uint32_t>(6, k <= 64 ? k / 8 : k / 16);
}
int tier_of(const pwknn_index* h, uint32_t k) {
    if (h->opt_force_tier == 2 && k <= kSelectMaxK) return 2;
    if (h->opt_force_tier == 3 || h->opt_kernel == 2) return 3;
    const uint64_t n = h->current_size;
    const uint64_t k_need = std::min<uint64_t>(static_cast<uint64_t>(k) + fused_margin(h, k), n);
    if (k_need <= kFastMaxKPrime) return 1;
    if (k <= kSelectMaxK) return 2;
    return 3;
}

int pack_subbatch(pwknn_index* h, const SubBatch& sb, uint64_t first, uint64_t count, cudaStream_t s) {
    int rc = ensure_queries(h, count); if (rc) return rc;
    const uint64_t qpad = round_up(count, 2 * kBlockM);
    const uint32_t* idx = sb.d_qindex ? sb.d_qindex + first : nullptr;
    const uint8_t* base = static_cast<const uint8_t*>(sb.src.p) + (sb.d_qindex ? 0 : first * sb.src.stride * elem_bytes(sb.src.kind));
    if (sb.src.kind == 0) {
        CU_TRY(h, launch_pack_queries_bf16(reinterpret_cast<const __nv_bfloat16*>(base), sb.src.stride, idx, count, h->dims,
                                           h->dpad, qpad, h->d_q.p, h->d_gthr.p, s));
    } else if (sb.src.stride != h->dims) {
        return fail(h, PWKNN_E_INVALID, "f32 / f64 query batches must be dense (stride == dimensions)");
    } else if (sb.src.kind == 1) {
        CU_TRY(h, launch_pack_queries_f32(reinterpret_cast<const float*>(base), idx, count, h->dims, h->dpad, qpad,
                                          h->d_q.p, h->d_gthr.p, s));
    } else {
        CU_TRY(h, launch_pack_queries_f64(reinterpret_cast<const double*>(base), idx, count, h->dims, h->dpad, qpad,
                                          h->d_q.p, h->d_gthr.p, s));
    }
    h->stat_launches++;
    return PWKNN_OK;
}

// What the f64 re-rank reads for queries [first, ...) of the sub-batch (whose packed bf16 copies sit in d_q rows 0..).
ExactView exact_view(const pwknn_index* h, const SubBatch& sb, uint64_t first) {
    ExactView xv{};
    if (h->shadow.kind) {
        xv.rows = h->shadow.rows; xv.row_kind = h->shadow.kind; xv.row_stride = h->shadow.stride; xv.len = h->dims;
    } else {
        xv.rows = h->d_rows; xv.row_kind = 0; xv.row_stride = h->dpad; xv.len = h->dpad;
    }
    if (h->shadow.kind && sb.src.kind != 0) {  // the caller's original f32 / f64 vectors
        xv.q_kind = sb.src.kind; xv.q_stride = static_cast<uint32_t>(sb.src.stride);
        if (sb.d_qindex) { xv.queries = sb.src.p; xv.q_index = sb.d_qindex + first; }
        else { xv.queries = static_cast<const uint8_t*>(sb.src.p) + first * sb.src.stride * elem_bytes(sb.src.kind); xv.q_index = nullptr; }
    } else {
        xv.queries = h->d_q.p; xv.q_kind = 0; xv.q_stride = h->dpad; xv.q_index = nullptr;
    }
    return xv;
}

// Tier 1: fused tcgen05 pass(es) + merge / f64 re-rank (+ certainty flags).
int run_fused(pwknn_index* h, const SubBatch& sb, cudaStream_t s) {
    const uint64_t nq = sb.nq;
    const uint32_t k = sb.k;
    const uint32_t n = static_cast<uint32_t>(h->current_size);
    int rc = pack_subbatch(h, sb, 0, nq, s); if (rc) return rc;
    const ExactView xv = exact_view(h, sb, 0);
    const Plan p = make_plan(h, nq, k);
    if (p.kernel != 1) return fail(h, PWKNN_E_INVALID, "fused path unavailable for k = %u", k);
    h->stat_last_kernel = 1; h->stat_last_kprime = p.a.kprime; h->stat_last_tier = 1;
    h->stat_last_two_pass = p.two_pass ? 1 : 0; h->stat_last_nq = static_cast<int64_t>(nq);
    size_t pn = static_cast<size_t>(p.a.slabs) * p.a.qpad * p.a.kprime;
    if (p.two_pass) pn = std::max(pn, static_cast<size_t>(p.b.slabs) * p.b.qpad * p.b.kprime);
    rc = ensure_dev(h, h->d_part_val, pn); if (rc) return rc;
    rc = ensure_dev(h, h->d_part_idx, pn); if (rc) return rc;
    if (p.two_pass) {
        rc = ensure_dev(h, h->d_flags, nq + p.b.qgroups); if (rc) return rc;
        CU_TRY(h, cudaMemsetAsync(h->d_flags.p, 0, (nq + p.b.qgroups) * sizeof(uint32_t), s));
    }
    uint32_t* d_incomplete = p.two_pass ? h->d_flags.p : nullptr;
    uint32_t* d_group_flags = p.two_pass ? h->d_flags.p + nq : nullptr;

    std::pair<cudaEvent_t, cudaEvent_t>* ev = nullptr;
    if (h->opt_profile && h->prof_used < 4096) {
        if (h->prof_used == h->prof_events.size()) {
            cudaEvent_t a, b;
            CU_TRY(h, cudaEventCreate(&a));
            CU_TRY(h, cudaEventCreate(&b));
            h->prof_events.emplace_back(a, b);
        }
        ev = &h->prof_events[h->prof_used++];
        CU_TRY(h, cudaEventRecord(ev->first, s));
    }
    FusedParams fp{};
    fill_fused_params(h, p.a, nq, &fp);
    fp.q_filter = sb.fv.q_filter; fp.filter_masks = sb.fv.masks; fp.filter_words = sb.fv.words;
    if (p.two_pass) fp.gthr = nullptr;  // a 16-entry list's last score is no bound for the k_need-th best
    rc = launch_fused_pass(h, p.a, fp, s); if (rc) return rc;
    if (ev && !p.two_pass) CU_TRY(h, cudaEventRecord(ev->second, s));

    MergeArgs m{};
    m.part_val = h->d_part_val.p; m.part_idx = h->d_part_idx.p; m.n_slabs = p.a.slabs; m.qpad = p.a.qpad;
    m.list_len = p.a.kprime; m.n_extract = std::min<uint32_t>(p.k_need, p.a.slabs * p.a.kprime);
    m.n_q = static_cast<uint32_t>(nq); m.k_out = k; m.ev = xv;
    m.metric = h->metric; m.n_valid = n; m.out_slots = sb.d_out_slots; m.out_dist = sb.d_out_dist;
    m.out_packed = sb.d_out_packed; m.id_offset = sb.id_offset;
    m.incomplete_out = d_incomplete; m.group_flags_out = d_group_flags;
    m.group_queries = p.two_pass ? p.b.ctas * kBlockM : kBlockM; m.incomplete_in = nullptr;
    m.noise = noise_bound(h); m.q_k = sb.d_qk; m.uncertain_out = h->opt_certainty ? sb.d_flags : nullptr;
    if (!h->opt_certainty && sb.d_flags) CU_TRY(h, cudaMemsetAsync(sb.d_flags, 0, nq * sizeof(uint32_t), s));
    CU_TRY(h, launch_merge_refine(m, s));
    h->stat_launches++;
    if (p.two_pass) {
        // pass b: full k_need lists, only for flagged query groups (unflagged groups cost a flag read per item)
        FusedParams fb{};
        fill_fused_params(h, p.b, nq, &fb);
        fb.q_filter = sb.fv.q_filter; fb.filter_masks = sb.fv.masks; fb.filter_words = sb.fv.words;
        fb.group_flags = d_group_flags;
        // thresholds were not touched by pass a (sharing off): still -inf from pack_queries
        rc = launch_fused_pass(h, p.b, fb, s); if (rc) return rc;
        if (ev) CU_TRY(h, cudaEventRecord(ev->second, s));
        MergeArgs m2 = m;
        m2.n_slabs = p.b.slabs; m2.qpad = p.b.qpad; m2.list_len = p.b.kprime;
        m2.n_extract = std::min<uint32_t>(p.k_need, p.b.slabs * p.b.kprime);
        m2.incomplete_out = nullptr; m2.group_flags_out = nullptr; m2.incomplete_in = d_incomplete;
        CU_TRY(h, launch_merge_refine(m2, s));
        h->stat_launches++;
    }
    h->stat_last_slabs = p.a.slabs; h->stat_last_stages = p.a.stages; h->stat_last_grid = p.a.grid;
    h->stat_last_ctas = p.a.ctas;
    return PWKNN_OK;
}

// Tier 2: per chunk of queries, one fused pass that dumps the fp32 ranking scores, then radix threshold -> gather ->
// f64 re-rank -> sort (launch_select_search).  Exact for any data given the noise bound; d_flags[q] = 1 when more
// than kSelectCap rows fall inside the 2-eps band of the threshold (masses of near-ties): those go to tier 3.
int run_select(pwknn_index* h, const SubBatch& sb, cudaStream_t s) {
    if (sb.d_out_packed) return fail(h, PWKNN_E_INVALID, "packed output needs the fused path (k + margin <= %u)", kFastMaxKPrime);
    if (!sb.d_flags) return fail(h, PWKNN_E_INVALID, "select path needs a flag buffer");
    const uint32_t n = static_cast<uint32_t>(h->current_size);
    const uint64_t ld = h->phys_rows;
    uint64_t q_chunk = std::max<uint64_t>(1, std::min<uint64_t>({sb.nq, static_cast<uint64_t>(h->opt_select_budget) / (ld * 4), 1024}));
    int rc = ensure_dev(h, h->d_scores, q_chunk * ld);
    while (rc == PWKNN_E_OOM && q_chunk > 1) {  // HBM is tight: smaller chunks (more corpus passes) instead of failing
        (void)cudaGetLastError();
        q_chunk = (q_chunk + 1) / 2;
        rc = ensure_dev(h, h->d_scores, q_chunk * ld);
    }
    if (rc) return rc;
    rc = ensure_dev(h, h->d_select_ws, select_workspace_bytes(static_cast<uint32_t>(q_chunk))); if (rc) return rc;
    h->stat_last_kernel = 3; h->stat_last_tier = 2; h->stat_last_kprime = 0; h->stat_last_two_pass = 0;
    for (uint64_t c0 = 0; c0 < sb.nq; c0 += q_chunk) {
        const uint64_t cn = std::min(q_chunk, sb.nq - c0);
        rc = pack_subbatch(h, sb, c0, cn, s); if (rc) return rc;
        const FusedPlan p = plan_fused(h, cn, 16);
        if (!p.ok) return fail(h, PWKNN_E_INVALID, "fused kernel unavailable");
        FusedParams fp{};
        fill_fused_params(h, p, cn, &fp);
        fp.part_val = nullptr; fp.part_idx = nullptr; fp.gthr = nullptr;  // dump launches keep no lists
        fp.dump = h->d_scores.p; fp.dump_ld = ld;
        rc = launch_fused_pass(h, p, fp, s); if (rc) return rc;
        SelectArgs a{};
        a.scores = h->d_scores.p; a.ld = ld; a.n_valid = n; a.nq = static_cast<uint32_t>(cn); a.k_out = sb.k;
        a.q_k = sb.d_qk ? sb.d_qk + c0 : nullptr;
        a.fv = sb.fv; if (a.fv.q_filter) a.fv.q_filter += c0;
        a.ev = exact_view(h, sb, c0); a.metric = h->metric; a.noise = noise_bound(h);
        a.workspace = h->d_select_ws.p;
        a.out_slots = sb.d_out_slots + c0 * sb.k; a.out_dist = sb.d_out_dist + c0 * sb.k;
        a.overflow_out = sb.d_flags + c0;
        CU_TRY(h, launch_select_search(a, s));
        h->stat_launches += 9;
        h->stat_last_slabs = p.slabs; h->stat_last_stages = p.stages; h->stat_last_grid = p.grid; h->stat_last_ctas = p.ctas;
    }
    return PWKNN_OK;
}

// Tier 3: f64 distances of every row + stable segmented sort (any k, any data).
int run_full(pwknn_index* h, const SubBatch& sb, cudaStream_t s) {
    if (sb.d_out_packed) return fail(h, PWKNN_E_INVALID, "packed output needs the fused path (k + margin <= %u)", kFastMaxKPrime);
    const uint32_t n = static_cast<uint32_t>(h->current_size);
    int rc = pack_subbatch(h, sb, 0, sb.nq, s); if (rc) return rc;
    const ExactView xv = exact_view(h, sb, 0);
    h->stat_last_kernel = 2; h->stat_last_tier = 3; h->stat_last_kprime = 0; h->stat_last_two_pass = 0;
    // chunk queries so the [q_chunk x N] sort workspace stays bounded
    const uint64_t budget_items = 96ull << 20;  // ~2.3 GB of keys+values
    uint32_t q_chunk = static_cast<uint32_t>(std::max<uint64_t>(1, std::min<uint64_t>(sb.nq, budget_items / std::max<uint32_t>(n, 1))));
    while (static_cast<uint64_t>(q_chunk) * n >= (1ull << 31) && q_chunk > 1) q_chunk /= 2;
    if (static_cast<uint64_t>(q_chunk) * n >= (1ull << 31))
        return fail(h, PWKNN_E_INVALID, "exact path supports < 2^31 rows per shard");
    const size_t ws = exact_workspace_bytes(n, q_chunk);
    rc = ensure_dev(h, h->d_exact_ws, ws); if (rc) return rc;
    for (uint64_t q0 = 0; q0 < sb.nq; q0 += q_chunk) {
        const uint32_t qc = static_cast<uint32_t>(std::min<uint64_t>(q_chunk, sb.nq - q0));
        CU_TRY(h, launch_exact_search(xv, sb.fv, n, h->metric, static_cast<uint32_t>(q0), qc, sb.k, h->d_exact_ws.p,
                                      h->d_exact_ws.cap, sb.d_out_slots, sb.d_out_dist, s));
        h->stat_launches += 4;
    }
    if (sb.d_flags) CU_TRY(h, cudaMemsetAsync(sb.d_flags, 0, sb.nq * sizeof(uint32_t), s));
    h->stat_last_slabs = 0; h->stat_last_stages = 0; h->stat_last_grid = 0;
    return PWKNN_OK;
}

int run_tier(pwknn_index* h, int tier, const SubBatch& sb, cudaStream_t s) {
    return tier == 1 ? run_fused(h, sb, s) : (tier == 2 ? run_select(h, sb, s) : run_full(h, sb, s));
}

// Finite test on the exponent bits (integer compare: auto-vectorised, one streaming pass).
// "usable" = finite AND |x| < 2^127: the tensor-core pass ranks on bf16, and a value that rounds to inf there
// (anything from the top binade up) would poison its row.
inline bool all_finite(const double* r, uint64_t n) {
    const uint64_t m = 0x7FF0000000000000ull, lim = 0x47E0000000000000ull;  // biased exponent 1150 = 2^127
    uint64_t bad = 0;
    for (uint64_t i = 0; i < n; ++i) {
        uint64_t b;
        std::memcpy(&b, r + i, 8);
        bad |= static_cast<uint64_t>((b & m) >= lim);
    }
    return bad == 0;
}
inline bool all_finite(const float* r, uint64_t n) {
    const uint32_t m = 0x7F800000u, lim = 0x7F000000u;  // biased exponent 254 = 2^127
    uint32_t bad = 0;
    for (uint64_t i = 0; i < n; ++i) {
        uint32_t b;
        std::memcpy(&b, r + i, 4);
        bad |= static_cast<uint32_t>((b & m) >= lim);
    }
    return bad == 0;
}
template <typename T>
bool row_finite(const T* r, uint32_t d) { return all_finite(r, d); }

// One streaming pass: copy `bytes` from src to the pinned staging buffer with NON-TEMPORAL stores (the lines
// go straight to DRAM; a DMA read of lines that sit dirty in several cores' caches was measured ~8x slower)
// while OR-ing the "exponent all ones" (inf/NaN) test of every element.  dst must be 16-byte aligned.
template <typename T>
bool stream_copy_check(uint8_t* dst, const uint8_t* src, size_t bytes) {
    // f64: biased exponent >= 1151 (inf, NaN or |x| >= 2^128, which overflows bf16/f32); f32: exponent all ones.
    const __m128i m = sizeof(T) == 8 ? _mm_set_epi32(0x7FF00000, 0, 0x7FF00000, 0) : _mm_set1_epi32(0x7F800000);
    const __m128i lim = sizeof(T) == 8 ? _mm_set1_epi32(0x47E00000 - 1) : _mm_set1_epi32(0x7F000000 - 1);
    const __m128i sel = sizeof(T) == 8 ? _mm_set_epi32(-1, 0, -1, 0) : _mm_set1_epi32(-1);
    __m128i acc = _mm_setzero_si128();
    size_t i = 0;
    for (; i + 64 <= bytes; i += 64) {
        const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i));
        const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i + 16));
        const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i + 32));
        const __m128i d = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i + 48));
        acc = _mm_or_si128(acc, _mm_or_si128(_mm_or_si128(_mm_cmpgt_epi32(_mm_and_si128(a, m), lim), _mm_cmpgt_epi32(_mm_and_si128(b, m), lim)),
                                             _mm_or_si128(_mm_cmpgt_epi32(_mm_and_si128(c, m), lim), _mm_cmpgt_epi32(_mm_and_si128(d, m), lim))));
        _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i), a);
        _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i + 16), b);
        _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i + 32), c);
        _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i + 48), d);
    }
    bool ok = _mm_movemask_epi8(_mm_and_si128(acc, sel)) == 0;
    if (i < bytes) {
        std::memcpy(dst + i, src + i, bytes - i);
        ok = ok && all_finite(reinterpret_cast<const T*>(src + i), (bytes - i) / sizeof(T));
    }
    _mm_sfence();
    return ok;
}

// dst (f32) <- src (f64) with non-temporal stores and the same usability test: the device would round
// f64 -> f32 -> bf16 anyway, so narrowing on the host halves the pinned-buffer writes and the PCIe bytes.
inline bool stream_narrow_check(float* dst, const double* src, size_t n) {
    const __m128i m = _mm_set_epi32(0x7FF00000, 0, 0x7FF00000, 0), lim = _mm_set1_epi32(0x47E00000 - 1);
    const __m128i sel = _mm_set_epi32(-1, 0, -1, 0);
    __m128i acc = _mm_setzero_si128();
    size_t i = 0;
    for (; i + 8 <= n; i += 8) {
        const __m128d a = _mm_loadu_pd(src + i), b = _mm_loadu_pd(src + i + 2), c = _mm_loadu_pd(src + i + 4),
                      d = _mm_loadu_pd(src + i + 6);
        acc = _mm_or_si128(acc, _mm_or_si128(_mm_or_si128(_mm_cmpgt_epi32(_mm_and_si128(_mm_castpd_si128(a), m), lim),
                                                          _mm_cmpgt_epi32(_mm_and_si128(_mm_castpd_si128(b), m), lim)),
                                             _mm_or_si128(_mm_cmpgt_epi32(_mm_and_si128(_mm_castpd_si128(c), m), lim),
                                                          _mm_cmpgt_epi32(_mm_and_si128(_mm_castpd_si128(d), m), lim))));
        _mm_stream_ps(dst + i, _mm_movelh_ps(_mm_cvtpd_ps(a), _mm_cvtpd_ps(b)));       // RNE, like __double2float_rn
        _mm_stream_ps(dst + i + 4, _mm_movelh_ps(_mm_cvtpd_ps(c), _mm_cvtpd_ps(d)));
    }
    bool ok = _mm_movemask_epi8(_mm_and_si128(acc, sel)) == 0;
    for (; i < n; ++i) { dst[i] = static_cast<float>(src[i]); }
    if (n % 8) ok = ok && all_finite(src + (n - n % 8), n % 8);
    _mm_sfence();
    return ok;
}
inline bool narrow_check_parallel(float* dst, const double* src, uint64_t n) {
    const int parts = n * sizeof(double) >= (1u << 20) ? HostPool::get().threads() : 1;
    std::vector<uint8_t> ok(parts, 1);
    HostPool::get().parallel(parts, [&](int part, int nparts) {
        const uint64_t b = (n * part / nparts) & ~15ull, e = part + 1 == nparts ? n : (n * (part + 1) / nparts) & ~15ull;
        ok[part] = stream_narrow_check(dst + b, src + b, e - b) ? 1 : 0;
    });
    for (uint8_t v : ok) if (!v) return false;
    return true;
}

// dst <- src (n elements) in parallel; returns true when every element is finite (checked on the fly).
template <typename T>
bool copy_check_finite(T* dst, const T* src, uint64_t n) {
    const uint64_t bytes = n * sizeof(T);
    const int parts = bytes >= (1u << 20) ? HostPool::get().threads() : 1;
    std::vector<uint8_t> ok(parts, 1);
    HostPool::get().parallel(parts, [&](int part, int nparts) {
        const uint64_t b = (bytes * part / nparts) & ~63ull, e = part + 1 == nparts ? bytes : (bytes * (part + 1) / nparts) & ~63ull;
        ok[part] = stream_copy_check<T>(reinterpret_cast<uint8_t*>(dst) + b, reinterpret_cast<const uint8_t*>(src) + b, e - b) ? 1 : 0;
    });
    for (uint8_t v : ok) if (!v) return false;
    return true;
}

// Fresh generation id for the per-slot stamp arrays (sized for slots < n_slots).
uint32_t next_generation(pwknn_index* h, uint64_t n_slots) {
    if (h->stamp.size() < n_slots) {
        const size_t ncap = std::max<size_t>(n_slots, h->stamp.size() + h->stamp.size() / 2);
        h->stamp.resize(ncap, SlotStamp{0, 0, 0, 0});
    }
    if (++h->generation == 0) {  // wrapped: clear
        std::fill(h->stamp.begin(), h->stamp.end(), SlotStamp{0, 0, 0, 0});
        h->generation = 1;
    }
    return h->generation;
}

// ---- delta staging ring
int ensure_ring(pwknn_index* h, size_t min_slot_bytes) {
    const size_t want = std::max<size_t>(round_up(min_slot_bytes, 4096), 8u << 20);
    if (h->ring_bytes >= want) return PWKNN_OK;
    CU_TRY(h, cudaStreamSynchronize(h->stream));
    if (h->copy_stream) CU_TRY(h, cudaStreamSynchronize(h->copy_stream));
    if (h->ring[0].h) { cudaFreeHost(h->ring[0].h); cudaFree(h->ring[0].d); h->ring[0].h = nullptr; h->ring[0].d = nullptr; h->ring_bytes = 0; }
    if (!h->copy_stream) CU_TRY(h, cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    uint8_t *hp = nullptr, *dp = nullptr;
    CU_TRY(h, cudaHostAlloc(reinterpret_cast<void**>(&hp), want * pwknn_index::kRingSlots, cudaHostAllocPortable));
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&dp), want * pwknn_index::kRingSlots);
    if (e != cudaSuccess) { cudaFreeHost(hp); return fail(h, PWKNN_E_OOM, "cudaMalloc(delta ring) failed: %s", cudaGetErrorString(e)); }
    for (int i = 0; i < pwknn_index::kRingSlots; ++i) {
        auto& r = h->ring[i];
        r.h = hp + i * want; r.d = dp + i * want; r.used = false;
        if (!r.copied) CU_TRY(h, cudaEventCreateWithFlags(&r.copied, cudaEventDisableTiming));
        if (!r.done) CU_TRY(h, cudaEventCreateWithFlags(&r.done, cudaEventDisableTiming));
    }
    h->ring_bytes = want;
    return PWKNN_OK;
}
// next ring slot, free for the host to fill (waits for the kernel that last read it)
int ring_acquire(pwknn_index* h, pwknn_index::RingSlot** out) {
    auto& r = h->ring[h->ring_next];
    h->ring_next = (h->ring_next + 1) % pwknn_index::kRingSlots;
    if (r.used) CU_TRY(h, cudaEventSynchronize(r.done));
    r.used = false;
    *out = &r;
    return PWKNN_OK;
}
// H2D of the filled slot on the copy stream; the index stream waits for it.  Call ring_release after the kernel launch.
int ring_upload(pwknn_index* h, pwknn_index::RingSlot* r, size_t bytes) {
    CU_TRY(h, cudaMemcpyAsync(r->d, r->h, bytes, cudaMemcpyHostToDevice, h->copy_stream));
    CU_TRY(h, cudaEventRecord(r->copied, h->copy_stream));
    CU_TRY(h, cudaStreamWaitEvent(h->stream, r->copied, 0));
    return PWKNN_OK;
}
int ring_release(pwknn_index* h, pwknn_index::RingSlot* r) {
    CU_TRY(h, cudaEventRecord(r->done, h->stream));
    r->used = true;
    return PWKNN_OK;
}

// add(): rs:114-144.  Slot assignment on the host, row conversion on the device.
template <typename T>
int add_impl(pwknn_index* h, uint64_t n, const uint64_t* keys, const T* rows, const void* d_rows_bf16,
             uint64_t dev_stride, cudaStream_t user_stream, int32_t* status) {
    if (!h) return PWKNN_E_INVALID;
    if (n == 0) return PWKNN_OK;
    if (!keys || (!rows && !d_rows_bf16)) return fail(h, PWKNN_E_INVALID, "pwknn_add: null keys/rows");
    CU_TRY(h, cudaSetDevice(h->device));
    cudaStream_t s = h->stream;
    const auto t_begin = std::chrono::steady_clock::now();
    auto ns_since = [](std::chrono::steady_clock::time_point t0) {
        return std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count();
    };
    std::vector<int32_t> st(n, PWKNN_ROW_OK);
    // f64 rows are narrowed to f32 on the host unless an f64 re-rank store needs every bit
    const bool narrow = rows && std::is_same<T, double>::value && h->shadow.kind != 2;
    const uint64_t out_row_bytes = static_cast<uint64_t>(h->dims) * (narrow ? sizeof(float) : sizeof(T));
    // Host rows: the usability test (finite, |x| < 2^127: the reference has no check) is fused
    // into the ONE pass that copies the rows into the pinned ring whenever the batch fits into the ring ("single
    // pass": the rows' H2D copies start while the key map is still being updated); larger batches (bulk loads) test
    // first, because the growth rule needs the number of accepted rows before anything can be written.
    struct Chunk { uint64_t c0, cn; pwknn_index::RingSlot* r; size_t rows_off; };
    std::vector<Chunk> chunks;
    auto copy_rows = [&](uint8_t* dst, uint64_t c0, uint64_t cn) -> bool {  // parallel copy + usability test of rows [c0, c0 + cn)
        if constexpr (std::is_same<T, double>::value) {
            if (narrow) return narrow_check_parallel(reinterpret_cast<float*>(dst), reinterpret_cast<const double*>(rows + c0 * h->dims), cn * h->dims);
            return copy_check_finite(reinterpret_cast<T*>(dst), rows + c0 * h->dims, cn * h->dims);
        } else {
            return copy_check_finite(reinterpret_cast<T*>(dst), rows + c0 * h->dims, cn * h->dims);  // streaming stores
        }
    };
    uint64_t accepted = n;
    bool single_pass = false, mapped = false;
    const uint64_t size_before = h->current_size;
    std::vector<uint32_t> slot_of(n, kInvalidSlot);
    int64_t map_ns = 0;
    // slot assignment == KeyToU64IdMapper::get_next_free_u64_id (mod.rs:224-233), rows with st != OK are skipped.
    // Two-stage software prefetch: the bucket of key i + 16, then (from a read-only probe of key i + 8) the stamp
    // record of the slot it maps to.  Host state only (no CUDA): may run on a pool thread next to the row copies.
    auto assign_slots = [&]() {
        const auto t_map = std::chrono::steady_clock::now();
        h->key_to_slot.reserve(h->current_size + n);
        const uint32_t gen = next_generation(h, h->current_size + n);  // stamp[slot].val = latest batch row
        constexpr uint64_t kAheadA = 16, kAheadB = 8;
        std::vector<size_t> hs(n);  // every key is hashed once (~15 ns for a u128 key); independent -> pipelined
        for (uint64_t i = 0; i < n; ++i) hs[i] = KeyHash()(Key{keys[2 * i], keys[2 * i + 1]});
        for (uint64_t i = 0; i < std::min<uint64_t>(kAheadA, n); ++i) h->key_to_slot.prefetch_h(hs[i]);
        for (uint64_t i = 0; i < n; ++i) {
            if (i + kAheadA < n) h->key_to_slot.prefetch_h(hs[i + kAheadA]);
            if (i + kAheadB < n) {
                uint32_t ps;
                if (h->key_to_slot.find_h(Key{keys[2 * (i + kAheadB)], keys[2 * (i + kAheadB) + 1]}, hs[i + kAheadB], &ps)) __builtin_prefetch(&h->stamp[ps]);
            }
            if (st[i] != PWKNN_ROW_OK) continue;
            const Key k{keys[2 * i], keys[2 * i + 1]};
            uint32_t slot;
            if (h->key_to_slot.find_h(k, hs[i], &slot)) {
                // live key: overwrite in place (quirk fenced off: size not bumped); duplicates in a batch: last wins
                if (h->stamp[slot].gen == gen) slot_of[h->stamp[slot].a] = kInvalidSlot;
            } else {
                slot = static_cast<uint32_t>(h->current_size++);
                h->key_to_slot.set_h(k, hs[i], slot);
                h->slot_to_key.push_back(k);
            }
            slot_of[i] = slot;
            h->stamp[slot] = SlotStamp{gen, static_cast<uint32_t>(i), 0, 0};
        }
        map_ns += ns_since(t_map);
    };
    if (rows) {
        int rc = ensure_ring(h, out_row_bytes + 512); if (rc) return rc;
        const uint64_t per_slot = std::max<uint64_t>(1, (h->ring_bytes - 256) / (out_row_bytes + sizeof(uint32_t)));
        single_pass = n <= per_slot * pwknn_index::kRingSlots;
        if (single_pass) {
            const uint64_t n_chunks = (n + per_slot - 1) / per_slot;
            const uint64_t per = (n + n_chunks - 1) / n_chunks;  // even chunks: the last H2D is not a straggler
            for (uint64_t c0 = 0; c0 < n; c0 += per) {
                Chunk c{c0, std::min(per, n - c0), nullptr, 0};
                rc = ring_acquire(h, &c.r); if (rc) return rc;
                c.rows_off = round_up(c.cn * sizeof(uint32_t), 256);
                chunks.push_back(c);
            }
            // slices of a chunk's rows for the pool (16-element granules keep the non-temporal stores 64-byte aligned)
            const int threads = HostPool::get().threads();
            const int slices = (threads > 1 && n * out_row_bytes >= (1u << 20)) ? threads : 1;
            struct Piece { uint32_t chunk; uint64_t e0, e1; };
            std::vector<Piece> pieces;
            for (uint32_t ci = 0; ci < chunks.size(); ++ci) {
                const uint64_t elems = chunks[ci].cn * h->dims;
                for (int p = 0; p < slices; ++p) {
                    const uint64_t e0 = (elems * p / slices) & ~15ull, e1 = p + 1 == slices ? elems : (elems * (p + 1) / slices) & ~15ull;
                    if (e1 > e0) pieces.push_back(Piece{ci, e0, e1});
                }
            }
            std::vector<uint8_t> piece_ok(pieces.size(), 1);
            auto copy_piece = [&](size_t pi) {
                const Piece& pc = pieces[pi];
                const Chunk& c = chunks[pc.chunk];
                const T* src = rows + c.c0 * h->dims + pc.e0;
                uint8_t* dst = c.r->h + c.rows_off;
                bool ok;
                if constexpr (std::is_same<T, double>::value) {
                    ok = narrow ? stream_narrow_check(reinterpret_cast<float*>(dst) + pc.e0, reinterpret_cast<const double*>(src), pc.e1 - pc.e0)
                                : stream_copy_check<T>(dst + pc.e0 * sizeof(T), reinterpret_cast<const uint8_t*>(src), (pc.e1 - pc.e0) * sizeof(T));
                } else {
                    ok = stream_copy_check<T>(dst + pc.e0 * sizeof(T), reinterpret_cast<const uint8_t*>(src), (pc.e1 - pc.e0) * sizeof(T));
                }
                piece_ok[pi] = ok ? 1 : 0;
            };
            // The key map is updated (optimistically: every row usable) on a helper thread while the pool copies and tests
            // the rows chunk by chunk; a chunk's H2D starts as soon as it is staged, so that by the time the call
            // returns only the last chunk is still in flight.
            struct Joiner { std::thread t; ~Joiner() { if (t.joinable()) t.join(); } } mapper;
            std::exception_ptr map_err;  // (an exception must not escape the helper thread: handed to the caller's)
            if (threads > 1 && n >= 512) {
                mapper.t = std::thread([&] { try { assign_slots(); } catch (...) { map_err = std::current_exception(); } });
                mapped = true;
            }
            size_t p0 = 0;
            cudaError_t copy_err = cudaSuccess;
            for (uint32_t ci = 0; ci < chunks.size(); ++ci) {
                size_t p1 = p0;
                while (p1 < pieces.size() && pieces[p1].chunk == ci) ++p1;
                if (p1 - p0 > 1) HostPool::get().parallel(static_cast<int>(p1 - p0), [&](int part, int) { copy_piece(p0 + static_cast<size_t>(part)); });
                else for (size_t pi = p0; pi < p1; ++pi) copy_piece(pi);
                const Chunk& c = chunks[ci];
                if (copy_err == cudaSuccess)
                    copy_err = cudaMemcpyAsync(c.r->d + c.rows_off, c.r->h + c.rows_off, c.cn * out_row_bytes, cudaMemcpyHostToDevice, h->copy_stream);
                p0 = p1;
            }
            if (mapper.t.joinable()) mapper.t.join();
            if (map_err) std::rethrow_exception(map_err);
            CU_TRY(h, copy_err);
            std::vector<uint8_t> chunk_bad(chunks.size(), 0);
            for (size_t pi = 0; pi < pieces.size(); ++pi) if (!piece_ok[pi]) chunk_bad[pieces[pi].chunk] = 1;
            for (size_t ci = 0; ci < chunks.size(); ++ci)
                if (chunk_bad[ci])
                    for (uint64_t i = chunks[ci].c0; i < chunks[ci].c0 + chunks[ci].cn; ++i)
                        if (!row_finite(rows + i * h->dims, h->dims)) { st[i] = PWKNN_ROW_BAD_VECTOR; --accepted; }
            if (mapped && accepted != n) {
                // rare: a row was unusable after all.  add() only ever APPENDS slots, so the optimistic update is undone
                // exactly by popping them again; the assignment is then redone with the rejected rows skipped.
                for (uint64_t sl = h->current_size; sl-- > size_before;) { uint32_t dummy; h->key_to_slot.erase(h->slot_to_key[sl], &dummy); }
                h->slot_to_key.resize(size_before);
                h->current_size = size_before;
                std::fill(slot_of.begin(), slot_of.end(), kInvalidSlot);
                mapped = false;
            }
        } else {
            const uint64_t total = n * h->dims;
            const int parts = total * sizeof(T) >= (1u << 20) ? HostPool::get().threads() : 1;
            std::vector<uint8_t> ok(parts, 1);
            HostPool::get().parallel(parts, [&](int part, int nparts) {
                const uint64_t b = total * part / nparts, e = total * (part + 1) / nparts;
                ok[part] = all_finite(rows + b, e - b) ? 1 : 0;
            });
            if (!std::all_of(ok.begin(), ok.end(), [](uint8_t v) { return v != 0; }))
                for (uint64_t i = 0; i < n; ++i)
                    if (!row_finite(rows + i * h->dims, h->dims)) { st[i] = PWKNN_ROW_BAD_VECTOR; --accepted; }
        }
    }
    h->stat_add_check_ns += ns_since(t_begin);
    if (accepted + size_before > 0xFFFFFFF0ull) return fail(h, PWKNN_E_INVALID, "index limited to 2^32-16 rows per shard");
    // growth rule rs:115-128 (on the size BEFORE this batch, like the reference)
    if (accepted + size_before > h->current_allocated) {
        const uint64_t na = std::max<uint64_t>(2 * h->current_allocated, accepted + size_before);
        int rc = realloc_corpus(h, na, size_before);
        if (rc) return rc;
    }
    if (!mapped) assign_slots();
    h->stat_add_map_ns += map_ns;
    const auto t_stage = std::chrono::steady_clock::now();
    auto launch_rows = [&](const Chunk& c) -> int {  // slot numbers -> ring, H2D, upsert kernel for one staged chunk
        std::memcpy(c.r->h, slot_of.data() + c.c0, c.cn * sizeof(uint32_t));
        CU_TRY(h, cudaMemcpyAsync(c.r->d, c.r->h, c.cn * sizeof(uint32_t), cudaMemcpyHostToDevice, h->copy_stream));
        CU_TRY(h, cudaEventRecord(c.r->copied, h->copy_stream));
        CU_TRY(h, cudaStreamWaitEvent(s, c.r->copied, 0));
        const uint32_t* dsl = reinterpret_cast<const uint32_t*>(c.r->d);
        if (narrow || std::is_same<T, float>::value) {
            CU_TRY(h, launch_upsert_f32(reinterpret_cast<const float*>(c.r->d + c.rows_off), dsl, c.cn, h->dims, h->dpad,
                                        h->metric, h->d_rows, h->d_aux, h->shadow, h->d_xmax_sq, s));
        } else {
            CU_TRY(h, launch_upsert_f64(reinterpret_cast<const double*>(c.r->d + c.rows_off), dsl, c.cn, h->dims, h->dpad,
                                        h->metric, h->d_rows, h->d_aux, h->shadow, h->d_xmax_sq, s));
        }
        h->stat_launches++;
        return ring_release(h, c.r);
    };
    if (d_rows_bf16) {
        // rows already in HBM: only the slot numbers travel.  The caller may reuse d_rows as soon as the call returns
        // ("inputs are consumed before the call returns"), so this entry point waits for its kernel.
        if (user_stream) CU_TRY(h, cudaStreamSynchronize(user_stream));  // producer of d_rows_bf16 done
        int rc = ensure_ring(h, 0); if (rc) return rc;
        const uint64_t chunk = h->ring_bytes / sizeof(uint32_t);
        for (uint64_t c0 = 0; c0 < n; c0 += chunk) {
            const uint64_t cn = std::min(chunk, n - c0);
            pwknn_index::RingSlot* r = nullptr;
            rc = ring_acquire(h, &r); if (rc) return rc;
            std::memcpy(r->h, slot_of.data() + c0, cn * sizeof(uint32_t));
            rc = ring_upload(h, r, cn * sizeof(uint32_t)); if (rc) return rc;
            CU_TRY(h, launch_upsert_bf16(static_cast<const __nv_bfloat16*>(d_rows_bf16) + c0 * dev_stride, dev_stride,
                                         reinterpret_cast<const uint32_t*>(r->d), cn, h->dims, h->dpad, h->metric,
                                         h->d_rows, h->d_aux, h->shadow, h->d_xmax_sq, s));
            h->stat_launches++;
            rc = ring_release(h, r); if (rc) return rc;
        }
        CU_TRY(h, cudaStreamSynchronize(s));
    } else if (single_pass) {
        for (const Chunk& c : chunks) { int rc = launch_rows(c); if (rc) return rc; }
    } else {
        // slot layout: [cn slot numbers, padded to 256 B][cn rows]
        const uint64_t per_slot = std::max<uint64_t>(1, (h->ring_bytes - 256) / (out_row_bytes + sizeof(uint32_t)));
        for (uint64_t c0 = 0; c0 < n; c0 += per_slot) {
            Chunk c{c0, std::min(per_slot, n - c0), nullptr, 0};
            int rc = ring_acquire(h, &c.r); if (rc) return rc;
            c.rows_off = round_up(c.cn * sizeof(uint32_t), 256);
            copy_rows(c.r->h + c.rows_off, c.c0, c.cn);
            CU_TRY(h, cudaMemcpyAsync(c.r->d + c.rows_off, c.r->h + c.rows_off, c.cn * out_row_bytes, cudaMemcpyHostToDevice, h->copy_stream));
            rc = launch_rows(c); if (rc) return rc;
        }
    }
    // no wait: the inputs are in the ring, the copies and upserts drain in stream order before any later search
    h->stat_add_stage_ns += ns_since(t_stage);
    h->stat_add_total_ns += ns_since(t_begin);
    h->stat_delta_calls++;
    if (status) std::memcpy(status, st.data(), n * sizeof(int32_t));
    return PWKNN_OK;
}

// Validates the filter ids of a batch and uploads dirty filter masks; fills `fv` (masks / words).  filter_ids may be null.
int prepare_filters(pwknn_index* h, uint64_t nq, const uint32_t* filter_ids, std::vector<int32_t>& st, FilterView* fv,
                    cudaStream_t s) {
    *fv = FilterView();
    if (!filter_ids) return PWKNN_OK;
    bool any = false;
    for (uint64_t q = 0; q < nq; ++q) {
        const uint32_t f = filter_ids[q];
        if (f == kNoFilter) continue;
        if (f >= h->filters.size() || !h->filters[f].live) { st[q] = PWKNN_ROW_BAD_VECTOR; continue; }
        any = true;
    }
    if (!any) return PWKNN_OK;
    const uint32_t words = static_cast<uint32_t>(h->phys_rows / 32);
    const size_t need = h->filters.size() * static_cast<size_t>(words);
    if (words != h->mask_words || need > h->d_masks.cap) {  // layout changed: re-upload everything
        int rc = ensure_dev(h, h->d_masks, need); if (rc) return rc;
        h->mask_words = words;
        for (auto& f : h->filters) f.dirty = f.live;
    }
    for (size_t f = 0; f < h->filters.size(); ++f) {
        auto& flt = h->filters[f];
        if (!flt.live || !flt.dirty) continue;
        CU_TRY(h, cudaMemcpyAsync(h->d_masks.p + f * words, flt.bits.data(), static_cast<size_t>(words) * 4,
                                  cudaMemcpyHostToDevice, s));
        flt.dirty = false;
    }
    fv->q_filter = nullptr; fv->masks = h->d_masks.p; fv->words = words;  // q_filter: per class, see enqueue_round
    return PWKNN_OK;
}

// pwknn_search runs in phases so that a multi-device handle can stage once, enqueue on every device and only then
// wait: search_stage (host: pinned copy + usability test), search_begin (H2D, classification by k, first round of
// kernels + D2H; asynchronous), search_finish (wait, harvest, escalate uncertain / overflowing queries, repeat).
template <typename T>
struct StagedQueries {
    const void* pinned = nullptr;  // staged copy (f32 when narrowed, else T)
    bool narrow = false;           // f64 input narrowed to f32 on the way
    uint64_t h2d_bytes = 0;
};

template <typename T>
int search_stage(pwknn_index* h, uint64_t nq, const T* queries, std::vector<int32_t>& st, StagedQueries<T>* out) {
    const uint64_t qbytes = nq * h->dims * sizeof(T);
    if (h->stage_pending) {  // an asynchronous pwknn_search_host_packed may still be copying out of the staging buffer
        CU_TRY(h, cudaEventSynchronize(h->stage_ev[0]));
        h->stage_pending = false;
    }
    int rc = ensure_pin(h, h->h_stage, qbytes); if (rc) return rc;
    // stage the queries into pinned memory (parallel non-temporal copy) and test usability in the same pass;
    // f64 queries are narrowed to f32 on the way unless an exact re-rank store wants the original values
    bool narrow = false, all_ok;
    if constexpr (std::is_same<T, double>::value) {
        narrow = h->shadow.kind == 0;
        all_ok = narrow ? narrow_check_parallel(reinterpret_cast<float*>(h->h_stage.p), queries, nq * h->dims)
                        : copy_check_finite(reinterpret_cast<T*>(h->h_stage.p), queries, nq * h->dims);
    } else {
        all_ok = copy_check_finite(reinterpret_cast<T*>(h->h_stage.p), queries, nq * h->dims);
    }
    if (!all_ok) {
        for (uint64_t q = 0; q < nq; ++q)
            if (!row_finite(queries + q * h->dims, h->dims)) st[q] = PWKNN_ROW_BAD_VECTOR;
    }
    out->pinned = h->h_stage.p;
    out->narrow = narrow;
    out->h2d_bytes = narrow ? nq * h->dims * sizeof(float) : qbytes;
    return PWKNN_OK;
}

// Launches every class of h->ss.round (kernels + D2H of results and flags) on the index stream.
int enqueue_round(pwknn_index* h) {
    auto& ss = h->ss;
    cudaStream_t s = h->stream;
    const bool filters = !ss.filter_ids.empty() && ss.masks.masks != nullptr;
    uint64_t out_total = 0, flag_total = 0, meta_total = 0;
    for (auto& c : ss.round) {
        const uint64_t nqc = c.idx.size();
        c.out_base = out_total; out_total += nqc * c.kmax;
        c.flag_base = flag_total; flag_total += nqc;
        c.meta_base = meta_total; meta_total += 2 * nqc + (filters ? round_up(nqc, 2 * kBlockM) : 0);
    }
    int rc = ensure_dev(h, h->d_out_slots, out_total); if (rc) return rc;
    rc = ensure_dev(h, h->d_out_dist, out_total); if (rc) return rc;
    rc = ensure_pin(h, h->h_out_slots, out_total); if (rc) return rc;
    rc = ensure_pin(h, h->h_out_dist, out_total); if (rc) return rc;
    rc = ensure_dev(h, h->d_cflags, flag_total); if (rc) return rc;
    rc = ensure_pin(h, h->h_cflags, flag_total); if (rc) return rc;
    rc = ensure_dev(h, h->d_meta, meta_total); if (rc) return rc;
    rc = ensure_pin(h, h->h_meta, meta_total); if (rc) return rc;
    for (auto& c : ss.round) {
        const uint64_t nqc = c.idx.size();
        uint32_t* m = h->h_meta.p + c.meta_base;
        for (uint64_t i = 0; i < nqc; ++i) { m[i] = c.idx[i]; m[nqc + i] = ss.kq[c.idx[i]]; }
        if (filters) {
            const uint64_t qpad = round_up(nqc, 2 * kBlockM);
            for (uint64_t i = 0; i < qpad; ++i) m[2 * nqc + i] = i < nqc ? ss.filter_ids[c.idx[i]] : kNoFilter;
        }
    }
    CU_TRY(h, cudaMemcpyAsync(h->d_meta.p, h->h_meta.p, meta_total * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
    for (auto& c : ss.round) {
        const uint64_t nqc = c.idx.size();
        SubBatch sb;
        sb.nq = nqc; sb.k = c.kmax;
        sb.src.p = ss.src_p; sb.src.kind = ss.src_kind; sb.src.stride = ss.src_stride;
        bool identity = nqc == ss.kq.size(), uniform = true;
        for (uint64_t i = 0; i < nqc; ++i) { identity = identity && c.idx[i] == i; uniform = uniform && ss.kq[c.idx[i]] == c.kmax; }
        sb.d_qindex = identity ? nullptr : h->d_meta.p + c.meta_base;
        sb.d_qk = uniform ? nullptr : h->d_meta.p + c.meta_base + nqc;
        if (filters) { sb.fv = ss.masks; sb.fv.q_filter = h->d_meta.p + c.meta_base + 2 * nqc; }
        sb.d_out_slots = h->d_out_slots.p + c.out_base; sb.d_out_dist = h->d_out_dist.p + c.out_base;
        sb.d_flags = h->d_cflags.p + c.flag_base;
        rc = run_tier(h, c.tier, sb, s); if (rc) return rc;
    }
    CU_TRY(h, cudaMemcpyAsync(h->h_out_slots.p, h->d_out_slots.p, out_total * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    CU_TRY(h, cudaMemcpyAsync(h->h_out_dist.p, h->d_out_dist.p, out_total * sizeof(double), cudaMemcpyDeviceToHost, s));
    CU_TRY(h, cudaMemcpyAsync(h->h_cflags.p, h->d_cflags.p, flag_total * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    return PWKNN_OK;
}

constexpr uint64_t kRoundOutBudget = 64ull << 20;  // result entries (12 bytes each) one round may hold

// Moves deferred classes into the (empty or freshly built) round while the output budget allows; at least one.
void fill_round(pwknn_index::SearchState& ss) {
    uint64_t used = 0;
    for (auto& c : ss.round) used += c.idx.size() * c.kmax;
    while (!ss.later.empty()) {
        const uint64_t need = ss.later.back().idx.size() * ss.later.back().kmax;
        if (!ss.round.empty() && used + need > kRoundOutBudget) break;
        used += need;
        ss.round.push_back(std::move(ss.later.back()));
        ss.later.pop_back();
    }
}

// Appends classes for the given batch positions (all of one tier); tier 3 is cut so that a class fits the budget.
void add_classes(pwknn_index::SearchState& ss, int tier, const std::vector<uint32_t>& qs, std::vector<pwknn_index::QueryClass>& out) {
    if (qs.empty()) return;
    if (tier < 3) {
        pwknn_index::QueryClass c;
        c.tier = tier; c.idx = qs;
        for (uint32_t q : qs) c.kmax = std::max(c.kmax, ss.kq[q]);
        out.push_back(std::move(c));
        return;
    }
    std::vector<uint32_t> order = qs;  // similar k together: a class pays kmax rows for each of its queries
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return ss.kq[a] < ss.kq[b]; });
    pwknn_index::QueryClass c;
    c.tier = 3;
    for (uint32_t q : order) {
        const uint32_t km = std::max(c.kmax, ss.kq[q]);
        if (!c.idx.empty() && (c.idx.size() + 1) * static_cast<uint64_t>(km) > kRoundOutBudget / 2) {
            out.push_back(std::move(c));
            c = pwknn_index::QueryClass();
            c.tier = 3;
        }
        c.idx.push_back(q);
        c.kmax = std::max(c.kmax, ss.kq[q]);
    }
    out.push_back(std::move(c));
}

// Queries are on the device (src); classifies them by k and enqueues the first round.  kq[q] == 0: nothing to do.
int search_begin(pwknn_index* h, uint64_t nq, const QuerySource& src, const uint32_t* limits, std::vector<int32_t>& st,
                 const uint32_t* filter_ids) {
    auto& ss = h->ss;
    ss.round.clear(); ss.later.clear();
    ss.kq.assign(nq, 0); ss.filter_ids.clear(); ss.masks = FilterView();
    ss.res_off.assign(nq + 1, 0); ss.res_cnt.assign(nq, 0);
    ss.escalated_uncertain = 0; ss.escalated_overflow = 0;
    ss.src_p = src.p; ss.src_kind = src.kind; ss.src_stride = src.stride;
    h->dev_flags_n = 0;
    if (h->current_size) { int rc = prepare_filters(h, nq, filter_ids, st, &ss.masks, h->stream); if (rc) return rc; }
    if (ss.masks.masks) ss.filter_ids.assign(filter_ids, filter_ids + nq);
    std::vector<uint32_t> by_tier[4];
    for (uint64_t q = 0; q < nq; ++q) {
        if (st[q] == PWKNN_ROW_OK) ss.kq[q] = static_cast<uint32_t>(std::min<uint64_t>(limits[q], h->current_size));
        ss.res_off[q + 1] = ss.res_off[q] + ss.kq[q];
        if (ss.kq[q]) {
            if (!ss.filter_ids.empty() && st[q] != PWKNN_ROW_OK) ss.filter_ids[q] = kNoFilter;
            by_tier[tier_of(h, ss.kq[q])].push_back(static_cast<uint32_t>(q));
        }
    }
    ss.res_slots.resize(ss.res_off[nq]); ss.res_dist.resize(ss.res_off[nq]);
    if (h->current_size == 0 || ss.res_off[nq] == 0) return PWKNN_OK;  // rs:193-198: empty index -> empty replies
    std::vector<pwknn_index::QueryClass> all;
    for (int t = 3; t >= 1; --t) add_classes(ss, t, by_tier[t], all);  // `later` is consumed from the back: tier 1 first
    ss.later = std::move(all);
    fill_round(ss);
    return enqueue_round(h);
}

// Waits for the round in flight, copies the finished queries' rows into ss.res_*, re-answers flagged queries on the
// next tier, and repeats until nothing is left.
int search_finish(pwknn_index* h) {
    auto& ss = h->ss;
    while (!ss.round.empty()) {
        CU_TRY(h, cudaSetDevice(h->device));
        CU_TRY(h, cudaStreamSynchronize(h->stream));
        std::vector<uint32_t> up[4];
        for (auto& c : ss.round) {
            for (uint64_t i = 0; i < c.idx.size(); ++i) {
                const uint32_t q = c.idx[i];
                if (c.tier < 3 && h->h_cflags.p[c.flag_base + i]) {
                    up[c.tier + 1].push_back(q);
                    (c.tier == 1 ? ss.escalated_uncertain : ss.escalated_overflow)++;
                    continue;
                }
                const uint32_t* sl = h->h_out_slots.p + c.out_base + i * c.kmax;
                const double* di = h->h_out_dist.p + c.out_base + i * c.kmax;
                uint32_t got = 0;
                for (; got < ss.kq[q]; ++got) {
                    if (sl[got] == kInvalidSlot || sl[got] >= h->current_size) break;
                    ss.res_slots[ss.res_off[q] + got] = sl[got];
                    ss.res_dist[ss.res_off[q] + got] = di[got];
                }
                ss.res_cnt[q] = got;
            }
        }
        ss.round.clear();
        std::sort(up[2].begin(), up[2].end());
        std::sort(up[3].begin(), up[3].end());
        add_classes(ss, 2, up[2], ss.round);
        add_classes(ss, 3, up[3], ss.later);
        fill_round(ss);
        if (!ss.round.empty()) { int rc = enqueue_round(h); if (rc) return rc; }
    }
    h->stat_last_uncertain = ss.escalated_uncertain; h->stat_last_overflow = ss.escalated_overflow;
    h->stat_total_uncertain += ss.escalated_uncertain;
    return PWKNN_OK;
}

// H2D of the staged batch + search_begin.
template <typename T>
int search_upload_begin(pwknn_index* h, uint64_t nq, const StagedQueries<T>& sq, const uint32_t* limits,
                        std::vector<int32_t>& st, const uint32_t* filter_ids) {
    CU_TRY(h, cudaSetDevice(h->device));
    h->stat_h2d_bytes = static_cast<int64_t>(sq.h2d_bytes);
    QuerySource src;
    src.kind = (sq.narrow || !std::is_same<T, double>::value) ? 1u : 2u;
    src.stride = h->dims;
    bool any = false;
    for (uint64_t q = 0; q < nq && !any; ++q) any = st[q] == PWKNN_ROW_OK && limits[q] > 0;
    if (h->current_size && any) {
        int rc = ensure_dev(h, h->d_stage, nq * h->dims * sizeof(T)); if (rc) return rc;
        CU_TRY(h, cudaMemcpyAsync(h->d_stage.p, sq.pinned, sq.h2d_bytes, cudaMemcpyHostToDevice, h->stream));
    }
    src.p = h->d_stage.p;
    return search_begin(h, nq, src, limits, st, filter_ids);
}

template <typename T>
int search_impl(pwknn_index* h, uint64_t nq, const T* queries, const uint32_t* limits, uint64_t* out_keys,
                double* out_scores, uint32_t* out_counts, int32_t* status, const uint32_t* filter_ids = nullptr) {
    if (!h) return PWKNN_E_INVALID;
    if (nq == 0) return PWKNN_OK;
    if (!queries || !limits || !out_counts) return fail(h, PWKNN_E_INVALID, "pwknn_search: null argument");
    CU_TRY(h, cudaSetDevice(h->device));
    const auto t_call = std::chrono::steady_clock::now();
    auto since = [&] { return std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t_call).count(); };
    std::vector<int32_t> st(nq, PWKNN_ROW_OK);
    StagedQueries<T> sq;
    int rc = search_stage(h, nq, queries, st, &sq); if (rc) return rc;
    h->stat_stage_ns = since();
    std::fill(out_counts, out_counts + nq, 0u);
    if (h->current_size && (!out_keys || !out_scores)) {
        bool any = false;
        for (uint64_t q = 0; q < nq && !any; ++q) any = st[q] == PWKNN_ROW_OK && limits[q] > 0;
        if (any) return fail(h, PWKNN_E_INVALID, "pwknn_search: null output");
    }
    rc = search_upload_begin(h, nq, sq, limits, st, filter_ids);
    if (status) std::memcpy(status, st.data(), nq * sizeof(int32_t));
    if (rc) return rc;
    h->stat_enqueue_ns = since();
    rc = search_finish(h); if (rc) return rc;
    h->stat_sync_ns = since();

    const auto& ss = h->ss;
    uint64_t off = 0;
    for (uint64_t q = 0; q < nq; ++q) {
        const uint32_t got = ss.res_cnt[q];
        for (uint32_t r = 0; r < got; ++r) {
            const Key& k = h->slot_to_key[ss.res_slots[ss.res_off[q] + r]];
            out_keys[2 * (off + r)] = k.lo;
            out_keys[2 * (off + r) + 1] = k.hi;
            out_scores[off + r] = -ss.res_dist[ss.res_off[q] + r];  // score = -distance  rs:230
        }
        out_counts[q] = got;
        off += limits[q];
    }
    h->stat_search_host_ns = since();
    return PWKNN_OK;
}

// ---------------------------------------------------------------------------------------------------------
// Multi-device handle: one NonFilteringExternalIndex whose rows are spread over several GPUs of this process.
// A key lives on the shard that owns its hash range (shard_bounds; no global map; each shard keeps the reference's
// dense-slot bookkeeping for its rows).  search: stage the queries once, enqueue on every device, wait, k-way merge on the host ordered by
// (distance, shard, position) -- inside a shard that is the reference's (distance, slot) order.
inline uint32_t shard_of(const pwknn_index* h, uint64_t lo, uint64_t hi) {
    const uint64_t x = (lo ^ (hi * 0x9E3779B97F4A7C15ull)) * 0xD6E8FEB86659FD93ull;
    const uint32_t h32 = static_cast<uint32_t>(x >> 32);
    const size_t G = h->shards.size();
    for (size_t g = 0; g + 1 < G; ++g)
        if (h32 < h->shard_bounds[g]) return static_cast<uint32_t>(g);
    return static_cast<uint32_t>(G - 1);
}

// hash-range boundaries proportional to `weights` (equal weights = equal ranges)
void set_shard_bounds(pwknn_index* h, const double* weights) {
    const size_t G = h->shards.size();
    double total = 0;
    for (size_t g = 0; g < G; ++g) total += weights ? weights[g] : 1.0;
    h->shard_bounds.assign(G, 0xFFFFFFFFu);
    double acc = 0;
    for (size_t g = 0; g + 1 < G; ++g) {
        acc += weights ? weights[g] : 1.0;
        h->shard_bounds[g] = static_cast<uint32_t>(std::min(4294967295.0, acc / total * 4294967296.0));
    }
}

void multi_refresh_size(pwknn_index* h) {
    uint64_t total = 0;
    for (auto* sh : h->shards) total += sh->current_size;
    h->current_size = total;
}

int multi_fail(pwknn_index* h, pwknn_index* sh, int rc) {
    h->last_error = "device " + std::to_string(sh->device) + ": " + sh->last_error;
    g_last_error = h->last_error;
    return rc;
}

template <typename T>
int multi_add(pwknn_index* h, uint64_t n, const uint64_t* keys, const T* rows, int32_t* status) {
    if (n == 0) return PWKNN_OK;
    if (!keys || !rows) return fail(h, PWKNN_E_INVALID, "pwknn_add: null keys/rows");
    const size_t G = h->shards.size();
    multi_refresh_size(h);
    const uint64_t before = h->current_size;
    const uint64_t chunk = std::max<uint64_t>(1, (256ull << 20) / (h->dims * sizeof(T)));  // <= 256 MiB gathered at a time
    std::vector<std::vector<uint64_t>> idx(G);
    std::vector<uint64_t> gk;
    std::vector<T> gr;
    std::vector<int32_t> gs;
    for (uint64_t c0 = 0; c0 < n; c0 += chunk) {
        const uint64_t cn = std::min(chunk, n - c0);
        for (auto& v : idx) v.clear();
        for (uint64_t i = c0; i < c0 + cn; ++i) idx[shard_of(h, keys[2 * i], keys[2 * i + 1])].push_back(i);
        for (size_t g = 0; g < G; ++g) {
            const auto& ix = idx[g];
            if (ix.empty()) continue;
            gk.resize(2 * ix.size()); gr.resize(ix.size() * h->dims); gs.assign(ix.size(), PWKNN_ROW_OK);
            for (size_t j = 0; j < ix.size(); ++j) {
                gk[2 * j] = keys[2 * ix[j]]; gk[2 * j + 1] = keys[2 * ix[j] + 1];
                std::memcpy(gr.data() + j * h->dims, rows + ix[j] * h->dims, h->dims * sizeof(T));
            }
            int rc = add_impl<T>(h->shards[g], ix.size(), gk.data(), gr.data(), nullptr, 0, nullptr, gs.data());
            if (rc) { multi_refresh_size(h); return multi_fail(h, h->shards[g], rc); }
            if (status) for (size_t j = 0; j < ix.size(); ++j) status[ix[j]] = gs[j];
        }
    }
    multi_refresh_size(h);
    if (before + n > h->current_allocated)  // growth rule rs:115-128 on the totals (what pwknn_allocated reports)
        h->current_allocated = std::max<uint64_t>(2 * h->current_allocated, before + n);
    return PWKNN_OK;
}

int multi_remove(pwknn_index* h, uint64_t n, const uint64_t* keys, int32_t* status) {
    if (n == 0) return PWKNN_OK;
    if (!keys) return fail(h, PWKNN_E_INVALID, "pwknn_remove: null keys");
    const size_t G = h->shards.size();
    std::vector<std::vector<uint64_t>> idx(G);
    for (uint64_t i = 0; i < n; ++i) idx[shard_of(h, keys[2 * i], keys[2 * i + 1])].push_back(i);
    std::vector<uint64_t> gk;
    std::vector<int32_t> gs;
    for (size_t g = 0; g < G; ++g) {
        const auto& ix = idx[g];
        if (ix.empty()) continue;
        gk.resize(2 * ix.size()); gs.assign(ix.size(), PWKNN_ROW_OK);
        for (size_t j = 0; j < ix.size(); ++j) { gk[2 * j] = keys[2 * ix[j]]; gk[2 * j + 1] = keys[2 * ix[j] + 1]; }
        int rc = pwknn_remove(h->shards[g], ix.size(), gk.data(), gs.data());
        if (rc) { multi_refresh_size(h); return multi_fail(h, h->shards[g], rc); }
        if (status) for (size_t j = 0; j < ix.size(); ++j) status[ix[j]] = gs[j];
    }
    multi_refresh_size(h);
    if (4 * h->current_size < h->current_allocated && h->current_allocated / 2 >= h->minimum_allocated)
        h->current_allocated /= 2;  // shrink rule rs:174-185 on the totals
    return PWKNN_OK;
}

template <typename T>
int multi_search(pwknn_index* h, uint64_t nq, const T* queries, const uint32_t* limits, uint64_t* out_keys,
                 double* out_scores, uint32_t* out_counts, int32_t* status, const uint32_t* filter_ids) {
    if (nq == 0) return PWKNN_OK;
    if (!queries || !limits || !out_counts) return fail(h, PWKNN_E_INVALID, "pwknn_search: null argument");
    const size_t G = h->shards.size();
    multi_refresh_size(h);
    const auto t_call = std::chrono::steady_clock::now();
    auto since = [&] { return std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t_call).count(); };
    std::vector<int32_t> st(nq, PWKNN_ROW_OK);
    pwknn_index* first = h->shards[0];
    {
        cudaError_t e = cudaSetDevice(first->device);
        if (e != cudaSuccess) return fail(h, PWKNN_E_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
    }
    StagedQueries<T> sq;
    int rc = search_stage(first, nq, queries, st, &sq);
    if (rc) return multi_fail(h, first, rc);
    h->stat_stage_ns = since();
    h->stat_h2d_bytes = static_cast<int64_t>(sq.h2d_bytes * G);
    std::fill(out_counts, out_counts + nq, 0u);
    if (h->current_size && (!out_keys || !out_scores)) {
        for (uint64_t q = 0; q < nq; ++q)
            if (st[q] == PWKNN_ROW_OK && limits[q] > 0) return fail(h, PWKNN_E_INVALID, "pwknn_search: null output");
    }
    std::vector<char> begun(G, 0);
    for (size_t g = 0; g < G && !rc; ++g) {
        rc = search_upload_begin(h->shards[g], nq, sq, limits, st, filter_ids);
        begun[g] = 1;
        if (rc) multi_fail(h, h->shards[g], rc);
    }
    h->stat_enqueue_ns = since();
    for (size_t g = 0; g < G; ++g) {  // wait for everything that was enqueued (also on the error path), escalate per shard
        pwknn_index* sh = h->shards[g];
        if (!begun[g]) continue;
        if (rc) { cudaSetDevice(sh->device); cudaStreamSynchronize(sh->stream); sh->ss.round.clear(); sh->ss.later.clear(); continue; }
        const int r2 = search_finish(sh);
        if (r2) rc = multi_fail(h, sh, r2);
    }
    h->stat_sync_ns = since();
    if (status) std::memcpy(status, st.data(), nq * sizeof(int32_t));
    if (rc) return rc;
    h->stat_last_uncertain = 0; h->stat_last_overflow = 0;
    for (auto* sh : h->shards) { h->stat_last_uncertain += sh->stat_last_uncertain; h->stat_last_overflow += sh->stat_last_overflow; }

    // k-way merge of the G ascending lists per query; ties: lower shard, then position (= slot order in the shard)
    std::vector<uint64_t> out_off(nq + 1, 0);
    for (uint64_t q = 0; q < nq; ++q) out_off[q + 1] = out_off[q] + limits[q];
    const int parts = nq * G >= 2048 ? HostPool::get().threads() : 1;
    HostPool::get().parallel(parts, [&](int part, int nparts) {
        std::vector<uint32_t> pos(G), avail(G);
        for (uint64_t q = nq * part / nparts; q < nq * (part + 1) / nparts; ++q) {
            if (st[q] != PWKNN_ROW_OK) continue;
            for (size_t g = 0; g < G; ++g) { pos[g] = 0; avail[g] = h->shards[g]->ss.res_cnt[q]; }
            const uint64_t off = out_off[q];
            uint32_t got = 0;
            while (got < limits[q]) {
                int best = -1;
                double bd = 0;
                for (size_t g = 0; g < G; ++g) {
                    if (pos[g] >= avail[g]) continue;
                    const auto& gs = h->shards[g]->ss;
                    const double d = gs.res_dist[gs.res_off[q] + pos[g]];
                    // (OrderedFloat(d), shard): NaN greater than every number
                    const bool better = best < 0 || (std::isnan(bd) ? !std::isnan(d) : d < bd);
                    if (better) { best = static_cast<int>(g); bd = d; }
                }
                if (best < 0) break;
                const pwknn_index* sh = h->shards[best];
                const Key& k = sh->slot_to_key[sh->ss.res_slots[sh->ss.res_off[q] + pos[best]]];
                out_keys[2 * (off + got)] = k.lo;
                out_keys[2 * (off + got) + 1] = k.hi;
                out_scores[off + got] = -bd;  // score = -distance  rs:230
                ++pos[best];
                ++got;
            }
            out_counts[q] = got;
        }
    });
    h->stat_search_host_ns = since();
    return PWKNN_OK;
}

// "never aborts the process": no C++ exception (std::bad_alloc of a host vector, std::system_error of a thread) may
// cross the C ABI
template <typename F>
int guarded(pwknn_index* h, const char* what, F&& body) {
    try {
        return body();
    } catch (const std::bad_alloc&) {
        return fail(h, PWKNN_E_OOM, "%s: out of host memory", what);
    } catch (const std::exception& e) {
        return fail(h, PWKNN_E_INVALID, "%s: %s", what, e.what());
    } catch (...) {
        return fail(h, PWKNN_E_INVALID, "%s: unknown exception", what);
    }
}

}  // namespace

extern "C" {

uint32_t pwknn_abi_version(void) { return 2; }
const char* pwknn_last_global_error(void) { return g_last_error.c_str(); }

int pwknn_create(uint32_t dimensions, uint64_t reserved_space, uint64_t auxiliary_space, int metric, int device,
                 pwknn_index** out) {
    uint32_t flags = 0;
    if (const char* e = std::getenv("PWKNN_EXACT_RERANK")) {
        if (!strcmp(e, "f32")) flags = PWKNN_EXACT_RERANK_F32;
        else if (!strcmp(e, "f64")) flags = PWKNN_EXACT_RERANK_F64;
    }
    return pwknn_create_ex(dimensions, reserved_space, auxiliary_space, metric, device, flags, out);
}

int pwknn_create_ex(uint32_t dimensions, uint64_t reserved_space, uint64_t auxiliary_space, int metric, int device,
                    uint32_t flags, pwknn_index** out) {
    if (!out) return fail(nullptr, PWKNN_E_INVALID, "pwknn_create: out is null");
    if ((flags & 3u) == 3u || (flags & ~3u)) return fail(nullptr, PWKNN_E_INVALID, "pwknn_create_ex: bad flags");
    *out = nullptr;
    if (dimensions == 0 || dimensions > 65536) return fail(nullptr, PWKNN_E_INVALID, "pwknn_create: dimensions must be 1..65536");
    if (metric != PWKNN_METRIC_L2SQ && metric != PWKNN_METRIC_COS && metric != PWKNN_METRIC_IP)
        return fail(nullptr, PWKNN_E_INVALID, "pwknn_create: bad metric");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, PWKNN_E_NOGPU, "no CUDA device: %s", e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(nullptr, PWKNN_E_INVALID, "pwknn_create: device %d out of range (%d devices)", device, ndev);
    cudaDeviceProp prop{};
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) return fail(nullptr, PWKNN_E_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    if (prop.major != 10)
        return fail(nullptr, PWKNN_E_NOGPU, "device %d is sm_%d%d; libpwknn is built for sm_100a (B200) only", device, prop.major, prop.minor);
    pwknn_index* h = new pwknn_index();
    h->device = device; h->dims = dimensions; h->dpad = static_cast<uint32_t>(round_up(dimensions, kBlockK));
    h->metric = metric; h->minimum_allocated = reserved_space; h->auxiliary_space = auxiliary_space;
    h->num_sms = prop.multiProcessorCount; h->smem_optin = prop.sharedMemPerBlockOptin;
    h->shadow.kind = flags & 3u;
    h->shadow.stride = h->shadow.kind == 1 ? static_cast<uint32_t>(round_up(dimensions, 4))
                                          : (h->shadow.kind == 2 ? static_cast<uint32_t>(round_up(dimensions, 2)) : 0);
    if (const char* e = std::getenv("PWKNN_CTA_PAIR")) h->auto_ctas = std::atoi(e) == 2 ? 2 : 1;
    if (const char* e = std::getenv("PWKNN_LEAD")) h->opt_lead_tiles = std::max(0, std::atoi(e));
    auto bail = [&](int code) { pwknn_destroy(h); return code; };
    if (cudaSetDevice(device) != cudaSuccess) return bail(fail(nullptr, PWKNN_E_CUDA, "cudaSetDevice failed"));
    if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) return bail(fail(nullptr, PWKNN_E_CUDA, "cudaStreamCreate failed"));
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || !fn || qres != cudaDriverEntryPointSuccess)
        return bail(fail(nullptr, PWKNN_E_NOGPU, "driver lacks cuTensorMapEncodeTiled (%s)", cudaGetErrorString(e)));
    h->encode = reinterpret_cast<EncodeTiledFn>(fn);
    if (cudaMalloc(reinterpret_cast<void**>(&h->d_xmax_sq), sizeof(float)) != cudaSuccess ||
        cudaMemset(h->d_xmax_sq, 0, sizeof(float)) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->order_ev, cudaEventDisableTiming) != cudaSuccess)
        return bail(fail(nullptr, PWKNN_E_OOM, "pwknn_create: device scratch allocation failed"));
    int rc = realloc_corpus(h, reserved_space, 0);
    if (rc) { g_last_error = h->last_error; return bail(rc); }
    *out = h;
    return PWKNN_OK;
}

int pwknn_create_multi(uint32_t dimensions, uint64_t reserved_space, uint64_t auxiliary_space, int metric,
                       const int* devices, uint32_t n_devices, uint32_t flags, pwknn_index** out) {
    if (!out) return fail(nullptr, PWKNN_E_INVALID, "pwknn_create_multi: out is null");
    *out = nullptr;
    if (!devices || n_devices == 0 || n_devices > 64)
        return fail(nullptr, PWKNN_E_INVALID, "pwknn_create_multi: need 1..64 devices");
    pwknn_index* h = new pwknn_index();
    h->multi = true;
    h->device = devices[0]; h->dims = dimensions; h->metric = metric;
    h->minimum_allocated = reserved_space; h->current_allocated = reserved_space; h->auxiliary_space = auxiliary_space;
    const uint64_t per = (reserved_space + n_devices - 1) / n_devices;
    for (uint32_t g = 0; g < n_devices; ++g) {
        pwknn_index* sh = nullptr;
        int rc = pwknn_create_ex(dimensions, per, auxiliary_space, metric, devices[g], flags, &sh);
        if (rc) { pwknn_destroy(h); return rc; }  // message already in pwknn_last_global_error()
        h->shards.push_back(sh);
    }
    h->dpad = h->shards[0]->dpad; h->num_sms = h->shards[0]->num_sms; h->shadow.kind = h->shards[0]->shadow.kind;
    set_shard_bounds(h, nullptr);
    *out = h;
    return PWKNN_OK;
}

int pwknn_multi_set_weights(pwknn_index* h, uint32_t n_devices, const double* weights) {
    if (!h || !weights) return PWKNN_E_INVALID;
    if (!h->is_multi() || n_devices != h->shards.size())
        return fail(h, PWKNN_E_INVALID, "pwknn_multi_set_weights: needs a multi-device handle and one weight per device");
    if (pwknn_size(h) != 0)  // a key's shard is a pure function of the key: it cannot change under live rows
        return fail(h, PWKNN_E_INVALID, "pwknn_multi_set_weights: the index must be empty");
    for (uint32_t g = 0; g < n_devices; ++g)
        if (!(weights[g] > 0.0) || !std::isfinite(weights[g])) return fail(h, PWKNN_E_INVALID, "pwknn_multi_set_weights: weights must be positive");
    set_shard_bounds(h, weights);
    return PWKNN_OK;
}

int pwknn_destroy(pwknn_index* h) {
    if (!h) return PWKNN_OK;
    if (h->is_multi()) {
        for (auto* sh : h->shards) pwknn_destroy(sh);
        delete h;
        return PWKNN_OK;
    }
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    cudaFree(h->d_rows); cudaFree(h->d_aux); cudaFree(h->shadow.rows); cudaFree(h->d_q.p); cudaFree(h->d_stage.p); cudaFree(h->d_slots.p);
    cudaFree(h->d_gthr.p); cudaFree(h->d_flags.p); cudaFree(h->d_progress.p); cudaFree(h->d_masks.p); cudaFree(h->d_qfilter.p);
    cudaFreeHost(h->h_qfilter.p); cudaFree(h->d_part_val.p); cudaFree(h->d_part_idx.p); cudaFree(h->d_out_slots.p); cudaFree(h->d_out_dist.p);
    cudaFree(h->d_exact_ws.p); cudaFree(h->d_meta.p); cudaFree(h->d_cflags.p); cudaFree(h->d_scores.p); cudaFree(h->d_select_ws.p);
    cudaFree(h->d_cta_times.p); cudaFree(h->d_xmax_sq); cudaFreeHost(h->h_meta.p); cudaFreeHost(h->h_cflags.p);
    if (h->ring[0].h) cudaFreeHost(h->ring[0].h);
    if (h->ring[0].d) cudaFree(h->ring[0].d);
    for (auto& r : h->ring) { if (r.copied) cudaEventDestroy(r.copied); if (r.done) cudaEventDestroy(r.done); }
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    if (h->order_ev) cudaEventDestroy(h->order_ev);
    if (h->stage_ev[0]) { cudaEventDestroy(h->stage_ev[0]); cudaEventDestroy(h->stage_ev[1]); }
    for (auto& e : h->prof_events) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
    cudaFreeHost(h->h_stage.p); cudaFreeHost(h->h_slots.p); cudaFreeHost(h->h_out_slots.p); cudaFreeHost(h->h_out_dist.p);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    (void)cudaGetLastError();  // teardown errors must not surface in a later launch check
    return PWKNN_OK;
}

int pwknn_add(pwknn_index* h, uint64_t n, const uint64_t* keys, const double* rows, int32_t* status) {
    return guarded(h, "pwknn_add", [&]() -> int {
        if (h && h->is_multi()) return multi_add<double>(h, n, keys, rows, status);
        return add_impl<double>(h, n, keys, rows, nullptr, 0, nullptr, status);
    });
}
int pwknn_add_f32(pwknn_index* h, uint64_t n, const uint64_t* keys, const float* rows, int32_t* status) {
    return guarded(h, "pwknn_add_f32", [&]() -> int {
        if (h && h->is_multi()) return multi_add<float>(h, n, keys, rows, status);
        return add_impl<float>(h, n, keys, rows, nullptr, 0, nullptr, status);
    });
}
int pwknn_add_device_bf16(pwknn_index* h, uint64_t n, const uint64_t* keys, const void* d_rows, uint64_t row_stride_elems,
                          void* stream, int32_t* status) {
    if (h && h->is_multi()) return fail(h, PWKNN_E_INVALID, "pwknn_add_device_bf16: not available on a multi-device handle");
    if (h && row_stride_elems < h->dims) return fail(h, PWKNN_E_INVALID, "row_stride_elems < dimensions");
    if (h && !stream) {
        CU_TRY(h, cudaSetDevice(h->device));
        int rc = order_after_default_stream(h, stream); if (rc) return rc;
    }
    return guarded(h, "pwknn_add_device_bf16", [&]() -> int {
        return add_impl<float>(h, n, keys, nullptr, d_rows, row_stride_elems, static_cast<cudaStream_t>(stream), status);
    });
}

// remove(): rs:146-187
static int remove_impl(pwknn_index* h, uint64_t n, const uint64_t* keys, int32_t* status);
int pwknn_remove(pwknn_index* h, uint64_t n, const uint64_t* keys, int32_t* status) {
    if (!h) return PWKNN_E_INVALID;
    return guarded(h, "pwknn_remove", [&]() -> int {
        if (h->is_multi()) return multi_remove(h, n, keys, status);
        return remove_impl(h, n, keys, status);
    });
}
static int remove_impl(pwknn_index* h, uint64_t n, const uint64_t* keys, int32_t* status) {
    if (n == 0) return PWKNN_OK;
    if (!keys) return fail(h, PWKNN_E_INVALID, "pwknn_remove: null keys");
    CU_TRY(h, cudaSetDevice(h->device));
    cudaStream_t s = h->stream;
    const uint64_t old_size = h->current_size;
    const auto t_begin = std::chrono::steady_clock::now();
    auto ns_since = [](std::chrono::steady_clock::time_point t0) {
        return std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count();
    };
    // Slot bookkeeping of the whole batch (slot_book.h): key -> slot probes and the net hash-table updates run on the
    // host pool, the order-dependent swap replay on slot-indexed records; result = per-key status, the (src, dst) row
    // moves (sources in the pre-batch layout, >= the new size; destinations < it) and the new size.
    const uint32_t gen = next_generation(h, old_size);
    std::vector<uint32_t> pairs;
    std::vector<int32_t> st_local;
    if (!status) { st_local.assign(n, 0); }
    const ParallelRunner runner = [](int nparts, const std::function<void(int)>& body) {
        HostPool::get().parallel(nparts, [&](int part, int) { body(part); });
    };
    h->current_size = remove_batch(h->key_to_slot, h->slot_to_key, h->stamp, gen, keys, n, status ? status : st_local.data(),
                                   &pairs, runner, HostPool::get().threads());
    h->stat_rm_map_ns += ns_since(t_begin);
    for (auto& f : h->filters) {  // filter bits follow their rows (sources >= new size, destinations < it)
        if (!f.live) continue;
        for (size_t i = 0; i + 1 < pairs.size(); i += 2) {
            const uint32_t src = pairs[i], dst = pairs[i + 1];
            const uint32_t bit = (f.bits[src >> 5] >> (src & 31)) & 1u;
            f.bits[dst >> 5] = (f.bits[dst >> 5] & ~(1u << (dst & 31))) | (bit << (dst & 31));
        }
        for (uint64_t r = h->current_size; r < old_size; ++r) f.bits[r >> 5] &= ~(1u << (r & 31));
        f.dirty = true;
    }
    const uint64_t npairs = pairs.size() / 2;
    if (npairs) {  // through the staging ring, like the rows of add(): no wait, the moves drain in stream order
        int rc = ensure_ring(h, 0); if (rc) return rc;
        const uint64_t chunk = h->ring_bytes / (2 * sizeof(uint32_t));
        for (uint64_t c0 = 0; c0 < npairs; c0 += chunk) {
            const uint64_t cn = std::min(chunk, npairs - c0);
            pwknn_index::RingSlot* r = nullptr;
            rc = ring_acquire(h, &r); if (rc) return rc;
            std::memcpy(r->h, pairs.data() + 2 * c0, cn * 2 * sizeof(uint32_t));
            rc = ring_upload(h, r, cn * 2 * sizeof(uint32_t)); if (rc) return rc;
            CU_TRY(h, launch_move_rows(reinterpret_cast<const uint32_t*>(r->d), cn, h->dpad, h->d_rows, h->d_aux, h->shadow, s));
            h->stat_launches++;
            rc = ring_release(h, r); if (rc) return rc;
        }
    }
    if (h->current_size < old_size) {  // keep the invariant: rows >= size are zero / never selected
        CU_TRY(h, launch_init_rows(h->d_rows, h->d_aux, h->current_size, old_size, h->dpad, s));
        h->stat_launches++;
    }
    // shrink rule rs:174-185
    if (4 * h->current_size < h->current_allocated && h->current_allocated / 2 >= h->minimum_allocated) {
        int rc = realloc_corpus(h, h->current_allocated / 2, h->current_size);
        if (rc) return rc;
    }
    h->stat_rm_total_ns += ns_since(t_begin);
    h->stat_delta_calls++;
    return PWKNN_OK;
}

int pwknn_search(pwknn_index* h, uint64_t nq, const double* queries, const uint32_t* limits, uint64_t* out_keys,
                 double* out_scores, uint32_t* out_counts, int32_t* status) {
    return guarded(h, "pwknn_search", [&]() -> int {
        if (h && h->is_multi()) return multi_search<double>(h, nq, queries, limits, out_keys, out_scores, out_counts, status, nullptr);
        return search_impl<double>(h, nq, queries, limits, out_keys, out_scores, out_counts, status);
    });
}
int pwknn_search_f32(pwknn_index* h, uint64_t nq, const float* queries, const uint32_t* limits, uint64_t* out_keys,
                     double* out_scores, uint32_t* out_counts, int32_t* status) {
    return guarded(h, "pwknn_search_f32", [&]() -> int {
        if (h && h->is_multi()) return multi_search<float>(h, nq, queries, limits, out_keys, out_scores, out_counts, status, nullptr);
        return search_impl<float>(h, nq, queries, limits, out_keys, out_scores, out_counts, status);
    });
}

int pwknn_search_filtered(pwknn_index* h, uint64_t nq, const double* queries, const uint32_t* limits,
                          const uint32_t* filter_ids, uint64_t* out_keys, double* out_scores, uint32_t* out_counts,
                          int32_t* status) {
    return guarded(h, "pwknn_search_filtered", [&]() -> int {
        if (h && h->is_multi()) return multi_search<double>(h, nq, queries, limits, out_keys, out_scores, out_counts, status, filter_ids);
        return search_impl<double>(h, nq, queries, limits, out_keys, out_scores, out_counts, status, filter_ids);
    });
}

int pwknn_filter_create(pwknn_index* h, uint32_t* out_filter_id) {
    if (!h || !out_filter_id) return PWKNN_E_INVALID;
    if (h->is_multi()) {  // the shards allocate filter ids in lock-step, so one id names the filter on every device
        for (size_t g = 0; g < h->shards.size(); ++g) {
            uint32_t id = 0;
            int rc = pwknn_filter_create(h->shards[g], &id);
            if (rc) return multi_fail(h, h->shards[g], rc);
            if (g && id != *out_filter_id) return fail(h, PWKNN_E_INVALID, "filter ids diverged across devices");
            *out_filter_id = id;
        }
        return PWKNN_OK;
    }
    size_t f = 0;
    while (f < h->filters.size() && h->filters[f].live) ++f;
    if (f == h->filters.size()) {
        if (h->filters.size() >= 1024) return fail(h, PWKNN_E_INVALID, "too many filters (1024)");
        h->filters.emplace_back();
    }
    h->filters[f].bits.assign(h->phys_rows / 32, 0u);
    h->filters[f].live = true; h->filters[f].dirty = true;
    *out_filter_id = static_cast<uint32_t>(f);
    return PWKNN_OK;
}

int pwknn_filter_destroy(pwknn_index* h, uint32_t filter_id) {
    if (h && h->is_multi()) {
        int rc = PWKNN_OK;
        for (auto* sh : h->shards) { int r = pwknn_filter_destroy(sh, filter_id); if (r) rc = r; }
        return rc;
    }
    if (!h || filter_id >= h->filters.size() || !h->filters[filter_id].live) return PWKNN_E_INVALID;
    h->filters[filter_id].live = false;
    std::vector<uint32_t>().swap(h->filters[filter_id].bits);
    return PWKNN_OK;
}

int pwknn_filter_set(pwknn_index* h, uint32_t filter_id, uint64_t n, const uint64_t* keys, const uint8_t* pass,
                     int32_t* status) {
    if (!h || (n && (!keys || !pass))) return PWKNN_E_INVALID;
    if (h->is_multi()) {
        for (uint64_t i = 0; i < n; ++i) {  // row by row: filter_set is host-only bit twiddling
            pwknn_index* sh = h->shards[shard_of(h, keys[2 * i], keys[2 * i + 1])];
            int rc = pwknn_filter_set(sh, filter_id, 1, keys + 2 * i, pass + i, status ? status + i : nullptr);
            if (rc) return multi_fail(h, sh, rc);
        }
        return PWKNN_OK;
    }
    if (filter_id >= h->filters.size() || !h->filters[filter_id].live) return fail(h, PWKNN_E_INVALID, "unknown filter %u", filter_id);
    auto& f = h->filters[filter_id];
    for (uint64_t i = 0; i < n; ++i) {
        uint32_t slot;
        if (!h->key_to_slot.find(Key{keys[2 * i], keys[2 * i + 1]}, &slot)) { if (status) status[i] = PWKNN_ROW_MISSING_KEY; continue; }
        if (pass[i]) f.bits[slot >> 5] |= (1u << (slot & 31)); else f.bits[slot >> 5] &= ~(1u << (slot & 31));
        if (status) status[i] = PWKNN_ROW_OK;
    }
    f.dirty = true;
    return PWKNN_OK;
}

// Device tier: one k for the whole batch, bf16 queries in HBM.  Tier 1 (k + margin <= 108) only enqueues; the
// per-query certainty flags stay on the device (stat "last_uncertain" reads them).  Larger k runs the select path and
// WAITS for it (candidate overflow must be re-answered by the full sort before the call returns).
static int search_device_common(pwknn_index* h, uint64_t nq, const void* d_queries, uint64_t row_stride_elems, uint32_t k,
                                uint32_t* d_out_slots, double* d_out_dist, long long* d_out_packed, int64_t id_offset,
                                void* stream) {
    CU_TRY(h, cudaSetDevice(h->device));
    int rc = order_after_default_stream(h, stream); if (rc) return rc;
    cudaStream_t s = stream ? static_cast<cudaStream_t>(stream) : h->stream;
    if (s != h->stream) {  // deltas drain asynchronously on the index stream: a caller's stream must see them applied
        CU_TRY(h, cudaEventRecord(h->order_ev, h->stream));
        CU_TRY(h, cudaStreamWaitEvent(s, h->order_ev, 0));
    }
    rc = ensure_dev(h, h->d_cflags, nq); if (rc) return rc;
    SubBatch sb;
    sb.nq = nq; sb.k = k;
    sb.src.p = d_queries; sb.src.kind = 0; sb.src.stride = row_stride_elems;
    sb.d_out_slots = d_out_slots; sb.d_out_dist = d_out_dist; sb.d_out_packed = d_out_packed; sb.id_offset = id_offset;
    sb.d_flags = h->d_cflags.p;
    const uint32_t k_eff = static_cast<uint32_t>(std::min<uint64_t>(k, h->current_size));
    const int tier = tier_of(h, std::max<uint32_t>(k_eff, 1));
    h->ss.round.clear(); h->ss.later.clear();
    rc = run_tier(h, tier, sb, s); if (rc) return rc;
    h->dev_flags_n = nq;
    if (tier == 2) {
        rc = ensure_pin(h, h->h_cflags, nq); if (rc) return rc;
        CU_TRY(h, cudaMemcpyAsync(h->h_cflags.p, h->d_cflags.p, nq * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
        CU_TRY(h, cudaStreamSynchronize(s));
        bool overflow = false;
        for (uint64_t q = 0; q < nq && !overflow; ++q) overflow = h->h_cflags.p[q] != 0;
        if (overflow) { rc = run_full(h, sb, s); if (rc) return rc; }
    }
    return PWKNN_OK;
}

int pwknn_search_device(pwknn_index* h, uint64_t nq, const void* d_queries, uint64_t row_stride_elems, uint32_t k,
                        void* d_out_slots, void* d_out_dist, void* stream) {
    if (!h) return PWKNN_E_INVALID;
    SINGLE_DEVICE_ONLY(h, "pwknn_search_device");
    if (nq == 0 || k == 0) return PWKNN_OK;
    if (!d_queries || !d_out_slots || !d_out_dist) return fail(h, PWKNN_E_INVALID, "pwknn_search_device: null pointer");
    if (row_stride_elems < h->dims) return fail(h, PWKNN_E_INVALID, "row_stride_elems < dimensions");
    CU_TRY(h, cudaSetDevice(h->device));
    cudaStream_t s = stream ? static_cast<cudaStream_t>(stream) : h->stream;
    if (h->current_size == 0) {
        CU_TRY(h, cudaMemsetAsync(d_out_slots, 0xFF, nq * k * sizeof(uint32_t), s));
        std::vector<double> inf(nq * k, INFINITY);
        CU_TRY(h, cudaMemcpyAsync(d_out_dist, inf.data(), nq * k * sizeof(double), cudaMemcpyHostToDevice, s));
        CU_TRY(h, cudaStreamSynchronize(s));
        h->dev_flags_n = 0;
        return PWKNN_OK;
    }
    return search_device_common(h, nq, d_queries, row_stride_elems, k, static_cast<uint32_t*>(d_out_slots),
                                static_cast<double*>(d_out_dist), nullptr, 0, stream);
}

int pwknn_search_device_packed(pwknn_index* h, uint64_t nq, const void* d_queries, uint64_t row_stride_elems, uint32_t k,
                               int64_t id_offset, void* d_out_packed, void* stream) {
    if (!h) return PWKNN_E_INVALID;
    SINGLE_DEVICE_ONLY(h, "pwknn_search_device_packed");
    if (nq == 0 || k == 0) return PWKNN_OK;
    if (!d_queries || !d_out_packed) return fail(h, PWKNN_E_INVALID, "pwknn_search_device_packed: null pointer");
    if (row_stride_elems < h->dims) return fail(h, PWKNN_E_INVALID, "row_stride_elems < dimensions");
    if (h->current_size == 0) return fail(h, PWKNN_E_INVALID, "pwknn_search_device_packed: empty shard (use pwknn_search_device)");
    return search_device_common(h, nq, d_queries, row_stride_elems, k, nullptr, nullptr,
                                static_cast<long long*>(d_out_packed), id_offset, stream);
}

// Row-sharded deployment, one process per GPU: HOST queries in (what the Rust trait hands over), this shard's packed
// top-k left in HBM for the all-gather -- no device -> host -> device round trip between the search and the exchange.
int pwknn_search_host_packed(pwknn_index* h, uint64_t nq, const double* queries, uint32_t k, int64_t id_offset,
                             void* d_out_packed, void* stream) {
    if (!h) return PWKNN_E_INVALID;
    SINGLE_DEVICE_ONLY(h, "pwknn_search_host_packed");
    if (nq == 0 || k == 0) return PWKNN_OK;
    if (!queries || !d_out_packed) return fail(h, PWKNN_E_INVALID, "pwknn_search_host_packed: null pointer");
    if (h->current_size == 0) return fail(h, PWKNN_E_INVALID, "pwknn_search_host_packed: empty shard");
    CU_TRY(h, cudaSetDevice(h->device));
    cudaStream_t s = stream ? static_cast<cudaStream_t>(stream) : h->stream;
    if (s != h->stream) {
        CU_TRY(h, cudaEventRecord(h->order_ev, h->stream));
        CU_TRY(h, cudaStreamWaitEvent(s, h->order_ev, 0));
    }
    if (!h->stage_ev[0]) {
        CU_TRY(h, cudaEventCreateWithFlags(&h->stage_ev[0], cudaEventDisableTiming));
        CU_TRY(h, cudaEventCreateWithFlags(&h->stage_ev[1], cudaEventDisableTiming));
    }
    std::vector<int32_t> st(nq, PWKNN_ROW_OK);  // (search_stage waits for a previous call's H2D out of the staging buffer)
    StagedQueries<double> sq;
    int rc = search_stage(h, nq, queries, st, &sq); if (rc) return rc;
    for (uint64_t q = 0; q < nq; ++q)
        if (st[q] != PWKNN_ROW_OK) return fail(h, PWKNN_E_INVALID, "pwknn_search_host_packed: query %llu has a non-finite component", (unsigned long long)q);
    h->stat_h2d_bytes = static_cast<int64_t>(sq.h2d_bytes);
    rc = ensure_dev(h, h->d_stage, nq * h->dims * sizeof(double)); if (rc) return rc;
    CU_TRY(h, cudaMemcpyAsync(h->d_stage.p, sq.pinned, sq.h2d_bytes, cudaMemcpyHostToDevice, s));
    CU_TRY(h, cudaEventRecord(h->stage_ev[0], s));
    h->stage_pending = true;
    rc = ensure_dev(h, h->d_cflags, nq); if (rc) return rc;
    SubBatch sb;
    sb.nq = nq; sb.k = k;
    sb.src.p = h->d_stage.p; sb.src.kind = sq.narrow ? 1u : 2u; sb.src.stride = h->dims;
    sb.d_out_packed = static_cast<long long*>(d_out_packed); sb.id_offset = id_offset;
    sb.d_flags = h->d_cflags.p;
    const uint32_t k_eff = static_cast<uint32_t>(std::min<uint64_t>(k, h->current_size));
    if (tier_of(h, std::max<uint32_t>(k_eff, 1)) != 1)
        return fail(h, PWKNN_E_INVALID, "packed output needs the fused path (k + margin <= %u)", kFastMaxKPrime);
    h->ss.round.clear(); h->ss.later.clear();
    rc = run_fused(h, sb, s); if (rc) return rc;
    h->dev_flags_n = nq;
    return PWKNN_OK;
}

int pwknn_merge_shards_device(pwknn_index* h, uint32_t world, uint64_t nq, uint32_t k, const void* d_gathered,
                              void* d_out_ids, void* d_out_dist, void* stream) {
    if (!h || !d_gathered || !d_out_ids || !d_out_dist || world == 0) return PWKNN_E_INVALID;
    SINGLE_DEVICE_ONLY(h, "pwknn_merge_shards_device");
    if (static_cast<uint64_t>(world) * k * 16 > 200 * 1024) return fail(h, PWKNN_E_INVALID, "world*k too large for the merge kernel");
    CU_TRY(h, cudaSetDevice(h->device));
    cudaStream_t s = stream ? static_cast<cudaStream_t>(stream) : h->stream;
    CU_TRY(h, launch_merge_shards(static_cast<const long long*>(d_gathered), world, static_cast<uint32_t>(nq), k,
                                  static_cast<long long*>(d_out_ids), static_cast<double*>(d_out_dist), s));
    h->stat_launches++;
    return PWKNN_OK;
}

// ---- snapshot: the reference rebuilds the index by replaying (and re-embedding) the whole input
// after a restart; here the HBM state can be written to / restored from a file instead.
// Files are written to `<path>.tmp` and renamed (a crash never leaves a half-written snapshot under the final name);
// a multi-device manifest and its shard files carry the same random generation id, so a stale `path.shardN` of an
// earlier save cannot pair with a new manifest.  Loading trusts nothing in the header: sizes are bounded and checked
// against the file length before anything is allocated, and no C++ exception crosses the C ABI.
namespace {
struct SnapHeader {
    char magic[8];           // "PWKNNSN2"
    uint32_t dims, dpad, metric, shadow_kind, shadow_stride, xmax_sq_bits;
    uint64_t size, allocated, minimum_allocated, auxiliary_space, generation;
};
struct MultiSnapHeader {
    char magic[8];           // "PWKNNMU2"
    uint32_t n_shards, dims, metric, reserved;
    uint64_t minimum_allocated, allocated, auxiliary_space, generation;
    uint32_t bounds[64];  // hash-range boundaries (placement weights are part of the layout)
};
struct FileCloser { FILE* f; ~FileCloser() { if (f) fclose(f); } };

uint64_t snapshot_generation() {
    const uint64_t t = static_cast<uint64_t>(std::chrono::steady_clock::now().time_since_epoch().count());
    const uint64_t w = static_cast<uint64_t>(std::chrono::system_clock::now().time_since_epoch().count());
    return KeyHash::mix(t ^ KeyHash::mix(w) ^ reinterpret_cast<uintptr_t>(&t)) | 1ull;
}

long long file_size_of(FILE* f) {
    const long cur = ftell(f);
    if (cur < 0 || fseek(f, 0, SEEK_END) != 0) return -1;
    const long end = ftell(f);
    if (fseek(f, cur, SEEK_SET) != 0) return -1;
    return end;
}

int snapshot_save_single(pwknn_index* h, const char* path, uint64_t generation) {
    CU_TRY(h, cudaSetDevice(h->device));
    CU_TRY(h, cudaStreamSynchronize(h->stream));
    const std::string tmp = std::string(path) + ".tmp";
    {
        FileCloser fc{fopen(tmp.c_str(), "wb")};
        if (!fc.f) return fail(h, PWKNN_E_INVALID, "cannot open %s for writing", tmp.c_str());
        SnapHeader hd{};
        std::memcpy(hd.magic, "PWKNNSN2", 8);
        hd.dims = h->dims; hd.dpad = h->dpad; hd.metric = static_cast<uint32_t>(h->metric); hd.shadow_kind = h->shadow.kind;
        hd.shadow_stride = h->shadow.stride; hd.size = h->current_size; hd.allocated = h->current_allocated;
        hd.minimum_allocated = h->minimum_allocated; hd.auxiliary_space = h->auxiliary_space; hd.generation = generation;
        float xm = 0.f;
        CU_TRY(h, cudaMemcpy(&xm, h->d_xmax_sq, sizeof(float), cudaMemcpyDeviceToHost));
        std::memcpy(&hd.xmax_sq_bits, &xm, 4);
        if (fwrite(&hd, sizeof(hd), 1, fc.f) != 1) return fail(h, PWKNN_E_INVALID, "write failed: %s", tmp.c_str());
        if (hd.size && fwrite(h->slot_to_key.data(), sizeof(Key), hd.size, fc.f) != hd.size) return fail(h, PWKNN_E_INVALID, "write failed: %s", tmp.c_str());
        // device arrays through a bounded pinned bounce buffer
        const size_t bounce = 32u << 20;
        int rc = ensure_pin(h, h->h_stage, bounce); if (rc) return rc;
        auto dump = [&](const void* dptr, size_t bytes) -> int {
            for (size_t off = 0; off < bytes; off += bounce) {
                const size_t n = std::min(bounce, bytes - off);
                CU_TRY(h, cudaMemcpy(h->h_stage.p, static_cast<const uint8_t*>(dptr) + off, n, cudaMemcpyDeviceToHost));
                if (fwrite(h->h_stage.p, 1, n, fc.f) != n) return fail(h, PWKNN_E_INVALID, "write failed: %s", tmp.c_str());
            }
            return PWKNN_OK;
        };
        rc = dump(h->d_rows, hd.size * h->dpad * sizeof(__nv_bfloat16)); if (rc) return rc;
        rc = dump(h->d_aux, hd.size * sizeof(RowAux)); if (rc) return rc;
        if (h->shadow.kind) { rc = dump(h->shadow.rows, hd.size * h->shadow.stride * (h->shadow.kind == 1 ? 4 : 8)); if (rc) return rc; }
        if (fflush(fc.f) != 0) return fail(h, PWKNN_E_INVALID, "write failed: %s", tmp.c_str());
    }
    if (std::rename(tmp.c_str(), path) != 0) return fail(h, PWKNN_E_INVALID, "cannot rename %s to %s", tmp.c_str(), path);
    return PWKNN_OK;
}

// expect_generation == 0: any.  expect_dims / expect_metric == ~0u: any.
int snapshot_load_single(const char* path, int device, uint64_t expect_generation, uint32_t expect_dims,
                         uint32_t expect_metric, pwknn_index** out) {
    *out = nullptr;
    FileCloser fc{fopen(path, "rb")};
    if (!fc.f) return fail(nullptr, PWKNN_E_INVALID, "cannot open %s", path);
    SnapHeader hd{};
    const bool got = fread(&hd, sizeof(hd), 1, fc.f) == 1;
    if (std::memcmp(hd.magic, "PWKNNMU", 7) == 0)
        return fail(nullptr, PWKNN_E_INVALID, "%s is a multi-device snapshot: use pwknn_snapshot_load_multi", path);
    if (!got || std::memcmp(hd.magic, "PWKNNSN2", 8) != 0)
        return fail(nullptr, PWKNN_E_INVALID, "%s is not a pwknn snapshot (version 2)", path);
    // ---- validate before allocating anything
    if (hd.dims == 0 || hd.dims > 65536 || hd.dpad != round_up(hd.dims, kBlockK))
        return fail(nullptr, PWKNN_E_INVALID, "%s: bad dimensions in header", path);
    if (hd.metric != PWKNN_METRIC_L2SQ && hd.metric != PWKNN_METRIC_COS && hd.metric != PWKNN_METRIC_IP)
        return fail(nullptr, PWKNN_E_INVALID, "%s: unknown metric %u", path, hd.metric);
    if (hd.shadow_kind > 2) return fail(nullptr, PWKNN_E_INVALID, "%s: unknown exact-store kind %u", path, hd.shadow_kind);
    const uint32_t want_stride = hd.shadow_kind == 1 ? static_cast<uint32_t>(round_up(hd.dims, 4))
                                                     : (hd.shadow_kind == 2 ? static_cast<uint32_t>(round_up(hd.dims, 2)) : 0u);
    if (hd.shadow_stride != want_stride) return fail(nullptr, PWKNN_E_INVALID, "%s: snapshot layout mismatch", path);
    if (hd.size > 0xFFFFFFF0ull || hd.size > hd.allocated || hd.allocated > (1ull << 40) || hd.minimum_allocated > (1ull << 40))
        return fail(nullptr, PWKNN_E_INVALID, "%s: inconsistent sizes in header (size %llu, allocated %llu)", path,
                    (unsigned long long)hd.size, (unsigned long long)hd.allocated);
    if (expect_generation && hd.generation != expect_generation)
        return fail(nullptr, PWKNN_E_INVALID, "%s belongs to a different save than its manifest (stale shard file)", path);
    if ((expect_dims != ~0u && hd.dims != expect_dims) || (expect_metric != ~0u && hd.metric != expect_metric))
        return fail(nullptr, PWKNN_E_INVALID, "%s: dimensions / metric differ from the manifest", path);
    const unsigned long long per_row = sizeof(Key) + static_cast<unsigned long long>(hd.dpad) * 2 + sizeof(RowAux) +
                                       static_cast<unsigned long long>(hd.shadow_stride) * (hd.shadow_kind == 1 ? 4 : (hd.shadow_kind == 2 ? 8 : 0));
    const long long fsz = file_size_of(fc.f);
    if (fsz < 0 || static_cast<unsigned long long>(fsz) != sizeof(SnapHeader) + hd.size * per_row)
        return fail(nullptr, PWKNN_E_INVALID, "%s: file length does not match its header (truncated or corrupted)", path);
    pwknn_index* h = nullptr;
    int rc = pwknn_create_ex(hd.dims, hd.minimum_allocated, hd.auxiliary_space, static_cast<int>(hd.metric), device,
                             hd.shadow_kind, &h);
    if (rc) return rc;
    auto bail = [&](int code) { g_last_error = h->last_error; pwknn_destroy(h); return code; };
    rc = realloc_corpus(h, hd.allocated, 0); if (rc) return bail(rc);
    try {
        h->slot_to_key.resize(hd.size);
        if (hd.size && fread(h->slot_to_key.data(), sizeof(Key), hd.size, fc.f) != hd.size) return bail(fail(h, PWKNN_E_INVALID, "truncated snapshot"));
        h->key_to_slot.reserve(hd.size);
        for (uint64_t i = 0; i < hd.size; ++i) h->key_to_slot.set(h->slot_to_key[i], static_cast<uint32_t>(i));
        if (h->key_to_slot.size() != hd.size) return bail(fail(h, PWKNN_E_INVALID, "%s: duplicate keys in the slot table", path));
    } catch (const std::bad_alloc&) {
        return bail(fail(h, PWKNN_E_OOM, "out of host memory restoring %llu keys", (unsigned long long)hd.size));
    } catch (const std::exception& e) {
        return bail(fail(h, PWKNN_E_INVALID, "restoring the key table failed: %s", e.what()));
    }
    h->current_size = hd.size;
    const size_t bounce = 32u << 20;
    rc = ensure_pin(h, h->h_stage, bounce); if (rc) return bail(rc);
    auto restore = [&](void* dptr, size_t bytes) -> int {
        for (size_t off = 0; off < bytes; off += bounce) {
            const size_t n = std::min(bounce, bytes - off);
            if (fread(h->h_stage.p, 1, n, fc.f) != n) return fail(h, PWKNN_E_INVALID, "truncated snapshot");
            if (cudaMemcpy(static_cast<uint8_t*>(dptr) + off, h->h_stage.p, n, cudaMemcpyHostToDevice) != cudaSuccess)
                return fail(h, PWKNN_E_CUDA, "snapshot H2D copy failed");
        }
        return PWKNN_OK;
    };
    rc = restore(h->d_rows, hd.size * h->dpad * sizeof(__nv_bfloat16)); if (rc) return bail(rc);
    rc = restore(h->d_aux, hd.size * sizeof(RowAux)); if (rc) return bail(rc);
    if (h->shadow.kind) { rc = restore(h->shadow.rows, hd.size * h->shadow.stride * (h->shadow.kind == 1 ? 4 : 8)); if (rc) return bail(rc); }
    float xm;
    std::memcpy(&xm, &hd.xmax_sq_bits, 4);
    if (!(xm >= 0.f) || !std::isfinite(xm)) return bail(fail(h, PWKNN_E_INVALID, "%s: bad norm bound in header", path));
    if (cudaMemcpy(h->d_xmax_sq, &xm, sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) return bail(fail(h, PWKNN_E_CUDA, "snapshot H2D copy failed"));
    *out = h;
    return PWKNN_OK;
}
}  // namespace

int pwknn_snapshot_save(pwknn_index* h, const char* path) {
    if (!h || !path) return PWKNN_E_INVALID;
    try {
        if (!h->is_multi()) return snapshot_save_single(h, path, snapshot_generation());
        // shard files first, the manifest (renamed into place last) names their generation
        multi_refresh_size(h);
        MultiSnapHeader mh{};
        std::memcpy(mh.magic, "PWKNNMU2", 8);
        mh.n_shards = static_cast<uint32_t>(h->shards.size()); mh.dims = h->dims; mh.metric = static_cast<uint32_t>(h->metric);
        mh.minimum_allocated = h->minimum_allocated; mh.allocated = h->current_allocated; mh.auxiliary_space = h->auxiliary_space;
        mh.generation = snapshot_generation();
        for (size_t g = 0; g < h->shards.size(); ++g) mh.bounds[g] = h->shard_bounds[g];
        for (size_t g = 0; g < h->shards.size(); ++g) {
            const std::string part = std::string(path) + ".shard" + std::to_string(g);
            int rc = snapshot_save_single(h->shards[g], part.c_str(), mh.generation);
            if (rc) return multi_fail(h, h->shards[g], rc);
        }
        const std::string tmp = std::string(path) + ".tmp";
        {
            FileCloser fc{fopen(tmp.c_str(), "wb")};
            if (!fc.f) return fail(h, PWKNN_E_INVALID, "cannot open %s for writing", tmp.c_str());
            if (fwrite(&mh, sizeof(mh), 1, fc.f) != 1 || fflush(fc.f) != 0) return fail(h, PWKNN_E_INVALID, "write failed: %s", tmp.c_str());
        }
        if (std::rename(tmp.c_str(), path) != 0) return fail(h, PWKNN_E_INVALID, "cannot rename %s to %s", tmp.c_str(), path);
        return PWKNN_OK;
    } catch (const std::bad_alloc&) {
        return fail(h, PWKNN_E_OOM, "pwknn_snapshot_save: out of host memory");
    } catch (const std::exception& e) {
        return fail(h, PWKNN_E_INVALID, "pwknn_snapshot_save: %s", e.what());
    }
}

int pwknn_snapshot_load(const char* path, int device, pwknn_index** out) {
    if (!path || !out) return fail(nullptr, PWKNN_E_INVALID, "pwknn_snapshot_load: null argument");
    try {
        return snapshot_load_single(path, device, 0, ~0u, ~0u, out);
    } catch (const std::bad_alloc&) {
        return fail(nullptr, PWKNN_E_OOM, "pwknn_snapshot_load: out of host memory");
    } catch (const std::exception& e) {
        return fail(nullptr, PWKNN_E_INVALID, "pwknn_snapshot_load: %s", e.what());
    }
}

int pwknn_snapshot_load_multi(const char* path, const int* devices, uint32_t n_devices, pwknn_index** out) {
    if (!path || !out || !devices) return fail(nullptr, PWKNN_E_INVALID, "pwknn_snapshot_load_multi: null argument");
    *out = nullptr;
    MultiSnapHeader mh{};
    {
        FileCloser fc{fopen(path, "rb")};
        if (!fc.f) return fail(nullptr, PWKNN_E_INVALID, "cannot open %s", path);
        if (fread(&mh, sizeof(mh), 1, fc.f) != 1 || std::memcmp(mh.magic, "PWKNNMU2", 8) != 0)
            return fail(nullptr, PWKNN_E_INVALID, "%s is not a multi-device pwknn snapshot (version 2)", path);
    }
    if (mh.n_shards == 0 || mh.n_shards > 64) return fail(nullptr, PWKNN_E_INVALID, "%s: bad shard count %u", path, mh.n_shards);
    if (mh.n_shards != n_devices)  // keys are placed by hash range per shard: the count is part of the layout
        return fail(nullptr, PWKNN_E_INVALID, "snapshot has %u shards, %u devices given", mh.n_shards, n_devices);
    for (uint32_t g = 0; g + 1 < mh.n_shards; ++g)
        if (mh.bounds[g] > mh.bounds[g + 1]) return fail(nullptr, PWKNN_E_INVALID, "%s: hash ranges out of order", path);
    pwknn_index* h = nullptr;
    try {
        h = new pwknn_index();
        h->multi = true;
        h->device = devices[0]; h->dims = mh.dims; h->metric = static_cast<int>(mh.metric);
        h->minimum_allocated = mh.minimum_allocated; h->current_allocated = mh.allocated; h->auxiliary_space = mh.auxiliary_space;
        for (uint32_t g = 0; g < n_devices; ++g) {
            const std::string part = std::string(path) + ".shard" + std::to_string(g);
            pwknn_index* sh = nullptr;
            int rc = snapshot_load_single(part.c_str(), devices[g], mh.generation, mh.dims, mh.metric, &sh);
            if (rc) { pwknn_destroy(h); return rc; }
            h->shards.push_back(sh);
        }
        h->dpad = h->shards[0]->dpad; h->num_sms = h->shards[0]->num_sms; h->shadow.kind = h->shards[0]->shadow.kind;
        for (auto* sh : h->shards)
            if (sh->shadow.kind != h->shadow.kind) { pwknn_destroy(h); return fail(nullptr, PWKNN_E_INVALID, "%s: shards disagree on the exact-store kind", path); }
        h->shard_bounds.assign(mh.bounds, mh.bounds + n_devices);
        multi_refresh_size(h);
    } catch (const std::bad_alloc&) {
        if (h) pwknn_destroy(h);
        return fail(nullptr, PWKNN_E_OOM, "pwknn_snapshot_load_multi: out of host memory");
    } catch (const std::exception& e) {
        if (h) pwknn_destroy(h);
        return fail(nullptr, PWKNN_E_INVALID, "pwknn_snapshot_load_multi: %s", e.what());
    }
    *out = h;
    return PWKNN_OK;
}

int pwknn_slots_to_keys(pwknn_index* h, uint64_t n, const uint32_t* slots, uint64_t* out_keys) {
    if (!h || !slots || !out_keys) return PWKNN_E_INVALID;
    SINGLE_DEVICE_ONLY(h, "pwknn_slots_to_keys");
    for (uint64_t i = 0; i < n; ++i) {
        if (slots[i] == kInvalidSlot || slots[i] >= h->current_size) { out_keys[2 * i] = ~0ull; out_keys[2 * i + 1] = ~0ull; }
        else { out_keys[2 * i] = h->slot_to_key[slots[i]].lo; out_keys[2 * i + 1] = h->slot_to_key[slots[i]].hi; }
    }
    return PWKNN_OK;
}

// Reciprocal-rank fusion (hybrid_index.py:43-120) as one pass per query over a small open-addressing table:
// what the reference spreads over flatten -> concat -> groupby(query, match).sum -> groupby(query, sort_by) -> limit.
int pwknn_rrf_fuse(uint32_t n_lists, uint64_t nq, const uint64_t* const* list_keys, const uint32_t* const* list_counts,
                   const uint32_t* const* list_limits, double k_rrf, const uint32_t* limits, uint64_t* out_keys,
                   double* out_scores, uint32_t* out_counts) {
    if (n_lists == 0 || !list_keys || !list_counts || !limits || !out_keys || !out_scores || !out_counts)
        return fail(nullptr, PWKNN_E_INVALID, "pwknn_rrf_fuse: null argument");
    for (uint32_t l = 0; l < n_lists; ++l)
        if (!list_keys[l] || !list_counts[l]) return fail(nullptr, PWKNN_E_INVALID, "pwknn_rrf_fuse: null list %u", l);
    // prefix sums first (cheap, sequential), then the queries are independent: split them over the host pool
    std::vector<uint64_t> in_off(size_t(n_lists) * (nq + 1), 0), out_off(nq + 1, 0);
    uint64_t work = 0;
    for (uint32_t l = 0; l < n_lists; ++l) {
        uint64_t* off = in_off.data() + size_t(l) * (nq + 1);
        const uint32_t* lim = list_limits ? list_limits[l] : nullptr;
        for (uint64_t q = 0; q < nq; ++q) {
            if (lim && list_counts[l][q] > lim[q])
                return fail(nullptr, PWKNN_E_INVALID, "pwknn_rrf_fuse: list %u query %llu: count > limit", l,
                            (unsigned long long)q);
            off[q + 1] = off[q] + (lim ? lim[q] : list_counts[l][q]);
            work += list_counts[l][q];
        }
    }
    for (uint64_t q = 0; q < nq; ++q) out_off[q + 1] = out_off[q] + limits[q];
    struct Cand { uint64_t lo, hi; double score; };
    struct Ord { uint64_t inv_score_bits; uint32_t idx; };  // positive doubles order like their bit patterns
    const int parts = work >= 8192 ? int(std::min<uint64_t>(HostPool::get().threads(), (nq + 15) / 16)) : 1;
    HostPool::get().parallel(std::max(1, parts), [&](int part, int nparts) {
        std::vector<Cand> cands;
        std::vector<Ord> order;
        std::vector<int32_t> table;
        const uint64_t q0 = nq * part / nparts, q1 = nq * (part + 1) / nparts;
        for (uint64_t q = q0; q < q1; ++q) {
            size_t total = 0;
            for (uint32_t l = 0; l < n_lists; ++l) total += list_counts[l][q];
            size_t cap = 16;
            while (cap < 2 * total) cap <<= 1;
            table.assign(cap, -1);
            cands.clear();
            for (uint32_t l = 0; l < n_lists; ++l) {
                const uint64_t* keys = list_keys[l] + 2 * in_off[size_t(l) * (nq + 1) + q];
                const uint32_t cnt = list_counts[l][q];
                for (uint32_t r = 0; r < cnt; ++r) {
                    const uint64_t lo = keys[2 * r], hi = keys[2 * r + 1];
                    const double s = 1.0 / (k_rrf + double(r + 1));
                    size_t pos = size_t((lo ^ (hi * 0x9E3779B97F4A7C15ull)) * 0xD6E8FEB86659FD93ull >> 17) & (cap - 1);
                    for (;;) {
                        const int32_t at = table[pos];
                        if (at < 0) { table[pos] = int32_t(cands.size()); cands.push_back({lo, hi, s}); break; }
                        if (cands[at].lo == lo && cands[at].hi == hi) { cands[at].score += s; break; }
                        pos = (pos + 1) & (cap - 1);
                    }
                }
            }
            const size_t keep = std::min<size_t>(limits[q], cands.size());
            order.resize(cands.size());
            bool plain = true;  // every score a positive finite double (k_rrf > -1): compare bit patterns
            for (size_t i = 0; i < cands.size(); ++i) {
                uint64_t bits;
                memcpy(&bits, &cands[i].score, 8);
                plain &= cands[i].score > 0.0 && std::isfinite(cands[i].score);
                order[i] = {~bits, uint32_t(i)};
            }
            auto better = [&](const Ord& a, const Ord& b) {
                if (plain ? a.inv_score_bits != b.inv_score_bits : cands[a.idx].score != cands[b.idx].score)
                    return plain ? a.inv_score_bits < b.inv_score_bits : cands[a.idx].score > cands[b.idx].score;
                const Cand &x = cands[a.idx], &y = cands[b.idx];
                return x.hi != y.hi ? x.hi < y.hi : x.lo < y.lo;
            };
            if (keep * 2 >= order.size()) std::sort(order.begin(), order.end(), better);
            else std::partial_sort(order.begin(), order.begin() + keep, order.end(), better);
            const uint64_t o = out_off[q];
            for (size_t r = 0; r < keep; ++r) {
                const Cand& c = cands[order[r].idx];
                out_keys[2 * (o + r)] = c.lo;
                out_keys[2 * (o + r) + 1] = c.hi;
                out_scores[o + r] = c.score;
            }
            out_counts[q] = uint32_t(keep);
        }
    });
    return PWKNN_OK;
}

uint64_t pwknn_size(const pwknn_index* h) {
    if (!h) return 0;
    if (!h->is_multi()) return h->current_size;
    uint64_t total = 0;  // summed on demand: rows may have been bulk-loaded straight into a shard (pwknn_shard)
    for (auto* sh : h->shards) total += sh->current_size;
    return total;
}

pwknn_index* pwknn_shard(pwknn_index* h, uint32_t i) {
    if (!h || !h->is_multi() || i >= h->shards.size()) return nullptr;
    return h->shards[i];
}

int pwknn_shards_of_keys(const pwknn_index* h, uint64_t n, const uint64_t* keys, uint32_t* out_shards) {
    if (!h || (n && (!keys || !out_shards))) return PWKNN_E_INVALID;
    for (uint64_t i = 0; i < n; ++i) out_shards[i] = h->is_multi() ? shard_of(h, keys[2 * i], keys[2 * i + 1]) : 0;
    return PWKNN_OK;
}
uint64_t pwknn_allocated(const pwknn_index* h) { return h ? h->current_allocated : 0; }
uint32_t pwknn_dimensions(const pwknn_index* h) { return h ? h->dims : 0; }
int pwknn_device(const pwknn_index* h) { return h ? h->device : -1; }
void* pwknn_stream(const pwknn_index* h) { return h ? static_cast<void*>(h->stream) : nullptr; }
const char* pwknn_last_error(const pwknn_index* h) { return h ? h->last_error.c_str() : g_last_error.c_str(); }

int pwknn_read_rows(pwknn_index* h, uint64_t slot0, uint64_t n, double* out_rows) {
    if (!h || !out_rows) return PWKNN_E_INVALID;
    SINGLE_DEVICE_ONLY(h, "pwknn_read_rows");
    if (slot0 + n > h->phys_rows) return fail(h, PWKNN_E_INVALID, "pwknn_read_rows: range beyond allocation");
    if (n == 0) return PWKNN_OK;
    CU_TRY(h, cudaSetDevice(h->device));
    const size_t bytes = n * h->dims * sizeof(double);
    int rc = ensure_dev(h, h->d_stage, bytes); if (rc) return rc;
    CU_TRY(h, launch_read_rows(h->d_rows, slot0, n, h->dims, h->dpad, reinterpret_cast<double*>(h->d_stage.p), h->stream));
    h->stat_launches++;
    CU_TRY(h, cudaMemcpyAsync(out_rows, h->d_stage.p, bytes, cudaMemcpyDeviceToHost, h->stream));
    CU_TRY(h, cudaStreamSynchronize(h->stream));
    return PWKNN_OK;
}

int pwknn_set_option(pwknn_index* h, const char* name, int64_t value) {
    if (!h || !name) return PWKNN_E_INVALID;
    if (h->is_multi()) {
        for (auto* sh : h->shards) { int rc = pwknn_set_option(sh, name, value); if (rc) return multi_fail(h, sh, rc); }
        return PWKNN_OK;
    }
    if (!strcmp(name, "kernel")) { if (value < 0 || value > 2) return fail(h, PWKNN_E_INVALID, "kernel must be 0..2"); h->opt_kernel = static_cast<int>(value); return PWKNN_OK; }
    if (!strcmp(name, "slabs")) { if (value < 0) return fail(h, PWKNN_E_INVALID, "slabs must be >= 0"); h->opt_slabs = static_cast<int>(value); return PWKNN_OK; }
    if (!strcmp(name, "margin")) { h->opt_margin = static_cast<int>(value); return PWKNN_OK; }
    if (!strcmp(name, "lead_tiles")) { h->opt_lead_tiles = static_cast<int>(value); return PWKNN_OK; }
    if (!strcmp(name, "max_stages")) { h->opt_max_stages = static_cast<int>(value); return PWKNN_OK; }
    if (!strcmp(name, "debug_flags")) { h->opt_debug_flags = static_cast<int>(value); return PWKNN_OK; }
    if (!strcmp(name, "optimistic")) { h->opt_optimistic = value != 0; return PWKNN_OK; }
    if (!strcmp(name, "certainty")) { h->opt_certainty = value != 0; return PWKNN_OK; }
    if (!strcmp(name, "delta_stats")) {  // resets the host-side delta timers
        h->stat_add_check_ns = h->stat_add_map_ns = h->stat_add_stage_ns = h->stat_add_total_ns = 0;
        h->stat_rm_map_ns = h->stat_rm_total_ns = h->stat_delta_calls = 0;
        return PWKNN_OK;
    }
    if (!strcmp(name, "noise_sigmas")) { if (value < 0) return fail(h, PWKNN_E_INVALID, "noise_sigmas must be >= 0"); h->opt_noise_sigmas = static_cast<int>(value); return PWKNN_OK; }
    if (!strcmp(name, "select_budget")) { if (value < (1 << 20)) return fail(h, PWKNN_E_INVALID, "select_budget must be >= 1 MiB"); h->opt_select_budget = value; return PWKNN_OK; }
    if (!strcmp(name, "cta_times")) { h->opt_cta_times = value != 0; return PWKNN_OK; }
    if (!strcmp(name, "force_tier")) { if (value != 0 && value != 2 && value != 3) return fail(h, PWKNN_E_INVALID, "force_tier must be 0, 2 or 3"); h->opt_force_tier = static_cast<int>(value); return PWKNN_OK; }
    if (!strcmp(name, "cta_pair")) { if (value < 0 || value > 2) return fail(h, PWKNN_E_INVALID, "cta_pair must be 0..2"); h->opt_cta_pair = static_cast<int>(value); return PWKNN_OK; }
    if (!strcmp(name, "profile")) { h->opt_profile = value != 0; h->prof_used = 0; return PWKNN_OK; }  // also resets samples
    return fail(h, PWKNN_E_INVALID, "unknown option '%s'", name);
}

int64_t pwknn_get_stat(const pwknn_index* h, const char* name) {
    if (!h || !name) return -1;
    if (!strcmp(name, "shards")) return static_cast<int64_t>(h->shards.size());  // 0: single-device handle
    if (h->is_multi()) {
        // totals for the additive counters, the router's own clocks for the last search, else the first shard's value
        if (!strcmp(name, "last_uncertain")) return h->stat_last_uncertain;
        if (!strcmp(name, "last_overflow")) return h->stat_last_overflow;
        if (!strcmp(name, "launches") || !strcmp(name, "phys_rows") || !strcmp(name, "fused_ns_sum") || !strcmp(name, "fused_samples") ||
            !strcmp(name, "total_uncertain")) {
            int64_t total = 0;
            for (auto* sh : h->shards) { const int64_t v = pwknn_get_stat(sh, name); if (v < 0) return v; total += v; }
            return total;
        }
        if (!strcmp(name, "stage_ns")) return h->stat_stage_ns;
        if (!strcmp(name, "h2d_bytes")) return h->stat_h2d_bytes;
        if (!strcmp(name, "enqueue_ns")) return h->stat_enqueue_ns;
        if (!strcmp(name, "sync_ns")) return h->stat_sync_ns;
        if (!strcmp(name, "search_host_ns")) return h->stat_search_host_ns;
        return pwknn_get_stat(h->shards[0], name);
    }
    if (!strcmp(name, "launches")) return h->stat_launches;
    if (!strcmp(name, "last_kernel")) return h->stat_last_kernel;
    if (!strcmp(name, "last_slabs")) return h->stat_last_slabs;
    if (!strcmp(name, "last_kprime")) return h->stat_last_kprime;
    if (!strcmp(name, "last_stages")) return h->stat_last_stages;
    if (!strcmp(name, "last_grid")) return h->stat_last_grid;
    if (!strcmp(name, "last_ctas")) return h->stat_last_ctas;
    if (!strcmp(name, "last_two_pass")) return h->stat_last_two_pass;
    if (!strcmp(name, "last_incomplete")) {  // queries the optimistic pass flagged in the last two-pass search (syncs)
        if (!h->stat_last_two_pass || !h->d_flags.p || !h->stat_last_nq) return 0;
        std::vector<uint32_t> f(h->stat_last_nq);
        if (cudaStreamSynchronize(h->stream) != cudaSuccess) return -1;
        if (cudaMemcpy(f.data(), h->d_flags.p, f.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
        int64_t c = 0; for (uint32_t v : f) c += v != 0; return c;
    }
    if (!strcmp(name, "add_check_ns")) return h->stat_add_check_ns;
    if (!strcmp(name, "add_map_ns")) return h->stat_add_map_ns;
    if (!strcmp(name, "add_stage_ns")) return h->stat_add_stage_ns;
    if (!strcmp(name, "add_total_ns")) return h->stat_add_total_ns;
    if (!strcmp(name, "remove_map_ns")) return h->stat_rm_map_ns;
    if (!strcmp(name, "remove_total_ns")) return h->stat_rm_total_ns;
    if (!strcmp(name, "delta_calls")) return h->stat_delta_calls;
    if (!strncmp(name, "cta_time_", 9)) {  // diagnostics: "cta_time_<i>" = end - start of CTA i of the last fused launch (ns), "cta_start_<i>" below
        const long i = std::atol(name + 9);
        if (!h->d_cta_times.p || i < 0 || i >= 1024) return -1;
        unsigned long long t[2];
        if (cudaStreamSynchronize(h->stream) != cudaSuccess || cudaMemcpy(t, h->d_cta_times.p + 2 * i, sizeof(t), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
        return static_cast<int64_t>(t[1] - t[0]);
    }
    if (!strncmp(name, "cta_start_", 10)) {
        const long i = std::atol(name + 10);
        if (!h->d_cta_times.p || i < 0 || i >= 1024) return -1;
        unsigned long long t[2], t0[2];
        if (cudaStreamSynchronize(h->stream) != cudaSuccess || cudaMemcpy(t, h->d_cta_times.p + 2 * i, sizeof(t), cudaMemcpyDeviceToHost) != cudaSuccess ||
            cudaMemcpy(t0, h->d_cta_times.p, sizeof(t0), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
        return static_cast<int64_t>(t[0]) - static_cast<int64_t>(t0[0]);
    }
    if (!strcmp(name, "last_tier")) return h->stat_last_tier;
    if (!strcmp(name, "last_overflow")) return h->stat_last_overflow;
    if (!strcmp(name, "total_uncertain")) return h->stat_total_uncertain;
    if (!strcmp(name, "last_uncertain")) {  // queries of the last search the certainty test sent to the select path
        if (!h->dev_flags_n) return h->stat_last_uncertain;  // host tier: already escalated and re-answered
        // device tier: the flags are still on the device (waits for the search)
        std::vector<uint32_t> f(h->dev_flags_n);
        if (cudaSetDevice(h->device) != cudaSuccess || cudaStreamSynchronize(h->stream) != cudaSuccess) return -1;
        if (cudaMemcpy(f.data(), h->d_cflags.p, f.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
        int64_t c = 0; for (uint32_t v : f) c += v != 0; return c;
    }
    if (!strcmp(name, "phys_rows")) return static_cast<int64_t>(h->phys_rows);
    if (!strcmp(name, "num_sms")) return h->num_sms;
    if (!strcmp(name, "exact_rerank")) return h->shadow.kind;
    if (!strcmp(name, "stage_ns")) return h->stat_stage_ns;
    if (!strcmp(name, "h2d_bytes")) return h->stat_h2d_bytes;  // query bytes the last pwknn_search copied to the device
    if (!strcmp(name, "enqueue_ns")) return h->stat_enqueue_ns;
    if (!strcmp(name, "sync_ns")) return h->stat_sync_ns;
    if (!strcmp(name, "search_host_ns")) return h->stat_search_host_ns;
    if (!strcmp(name, "fused_samples")) return static_cast<int64_t>(h->prof_used);
    if (!strcmp(name, "fused_ns_sum")) {  // device time of the recorded fused-kernel launches (waits for them)
        double total_ms = 0;
        for (size_t i = 0; i < h->prof_used; ++i) {
            float ms = 0;
            if (cudaEventSynchronize(h->prof_events[i].second) != cudaSuccess) return -1;
            if (cudaEventElapsedTime(&ms, h->prof_events[i].first, h->prof_events[i].second) != cudaSuccess) return -1;
            total_ms += ms;
        }
        return static_cast<int64_t>(total_ms * 1e6);
    }
    return -1;
}

int pwknn_debug_scores_device(pwknn_index* h, uint64_t nq, const void* d_queries, uint64_t row_stride_elems,
                              void* d_out_scores, uint64_t ld_out, void* stream) {
    if (!h || !d_queries || !d_out_scores) return PWKNN_E_INVALID;
    SINGLE_DEVICE_ONLY(h, "pwknn_debug_scores_device");
    if (h->current_size == 0 || nq == 0) return PWKNN_OK;
    CU_TRY(h, cudaSetDevice(h->device));
    cudaStream_t s = stream ? static_cast<cudaStream_t>(stream) : h->stream;
    int rc = ensure_queries(h, nq); if (rc) return rc;
    const uint64_t qpad = round_up(nq, 2 * kBlockM);
    CU_TRY(h, launch_pack_queries_bf16(static_cast<const __nv_bfloat16*>(d_queries), row_stride_elems, nullptr, nq, h->dims,
                                       h->dpad, qpad, h->d_q.p, h->d_gthr.p, s));
    const FusedPlan p = plan_fused(h, nq, 16);
    if (!p.ok) return fail(h, PWKNN_E_INVALID, "fused kernel unavailable");
    FusedParams fp{};
    fill_fused_params(h, p, nq, &fp);
    fp.part_val = nullptr; fp.part_idx = nullptr; fp.gthr = nullptr;  // dump launches keep no lists
    fp.dump = static_cast<float*>(d_out_scores); fp.dump_ld = ld_out;
    rc = launch_fused_pass(h, p, fp, s); if (rc) return rc;
    h->stat_last_ctas = p.ctas;
    return PWKNN_OK;
}

}  // extern "C"
