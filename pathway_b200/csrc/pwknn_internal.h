// Internal (non-ABI) interfaces between the host layer and the CUDA kernels.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace pwknn {

constexpr uint32_t kBlockM = 128;      // queries per tile  (TMEM lanes, UMMA M)
constexpr uint32_t kBlockN = 256;      // corpus rows per tile (TMEM columns, UMMA N)
constexpr uint32_t kBlockK = 64;       // bf16 elements per smem stage row = one 128-B swizzle atom
constexpr uint32_t kUmmaK = 16;        // K per tcgen05.mma kind::f16
constexpr uint32_t kFastMaxKPrime = 108;    // candidates per (query, slab) the fused kernel can keep (smem lists)
constexpr uint32_t kMergeMaxCand = 8192;    // slabs * kprime the merge kernel sorts in smem
constexpr uint32_t kInvalidSlot = 0xFFFFFFFFu;
constexpr uint32_t kNoFilter = 0xFFFFFFFFu;
constexpr uint32_t kSelectMaxK = 4096;      // largest k the select path answers (score dump -> radix threshold -> f64 re-rank)
constexpr uint32_t kSelectCap = 8192;       // candidates per query the select path can re-rank and sort in one block

// Per corpus row epilogue constants of the fp32 ranking pass:  v = fma(dot, scale, bias).
//   COS : scale = 1/|x|, bias = 0        (zero row: scale 0, bias -FLT_MAX  -> "NaN distance, sorted last")
//   L2SQ: scale = 1,     bias = -|x|^2/2 (ranking by dot - |x|^2/2  ==  ranking by -l2sq)
//   rows >= size: scale 0, bias -inf (never selected)
struct RowAux {
    float scale;
    float bias;
};

// What the f64 re-rank and the exact path read.  kind: 0 = bf16, 1 = f32, 2 = f64.
// Default: rows = the bf16 corpus (stride dpad), queries = the packed bf16 batch, len = dpad.
// With an exact re-rank store (pwknn_create_ex): rows = the f32/f64 shadow copy, queries = the caller's
// original f32/f64 (or bf16) vectors, len = dims -> distances of the RAW inputs, as the reference computes them.
struct ExactView {
    const void* rows;
    uint32_t row_kind, row_stride;
    const void* queries;
    uint32_t q_kind, q_stride;
    uint32_t len;
    const uint32_t* q_index;  // optional: query i of the (sub-)batch is row q_index[i] of `queries` (per-query routing
                              // runs subsets of a staged batch); nullptr = identity
};

// Noise bound of the tensor-core ranking pass, in ranking-score units (COS: dot/|x|, L2SQ: dot - |x|^2/2, IP: dot):
//   eps(q) = (a0 + a1 * Xmax) * |q| + b * Xmax^2,   Xmax = largest row norm ever stored (device-resident, monotone).
// Default store: worst case of the fp32 accumulation (validated by scripts/noise_probe.py); exact re-rank store: plus
// an 8-sigma bound of the bf16 rounding noise.  A row outside the extracted candidates has fp32 score <= s_min, hence
// true score <= s_min + eps: the k-th result is CERTAIN when its exact score exceeds that (see merge_refine_kernel).
struct NoiseBound {
    float a0 = 0.f, a1 = 0.f, b = 0.f;
    const float* xmax_sq = nullptr;  // device: max |x|^2 over every row ever upserted
};

struct ShadowStore {          // optional exact copy of the corpus, written by the upsert / move kernels
    void* rows = nullptr;
    uint32_t kind = 0;        // 0 none, 1 f32, 2 f64
    uint32_t stride = 0;      // elements per row
};

struct FilterView {            // all null / 0 when the batch has no filtered query
    const uint32_t* q_filter = nullptr;
    const uint32_t* masks = nullptr;
    uint32_t words = 0;
};

struct FusedParams {
    const RowAux* aux;        // [phys_rows]
    uint32_t n_valid;         // current_size
    uint32_t n_q;             // real queries
    uint32_t n_qtiles;        // ceil(n_q / 128)
    uint32_t n_qgroups;       // ceil(n_qtiles / ctas): query tiles are handled in groups of `ctas`
    uint32_t ctas;            // 1, or 2 = CTA pair (cluster of 2, cta_group::2)
    uint32_t n_slabs;         // corpus slabs
    uint32_t n_tiles;         // ceil(n_valid / 256)
    uint32_t kprime;          // candidates kept per (query, slab)
    uint32_t k_chunks;        // dpad / 64
    uint32_t stages;          // smem pipeline depth actually used
    float* part_val;          // [n_slabs][n_qgroups*ctas*128][kprime]
    uint32_t* part_idx;       // same shape, corpus slot
    float* gthr;              // [qpad] cross-slab lower bound of the k'-th best score, init -inf; nullptr = no sharing
                              // (optimistic pass: lists are shorter than the needed k', a slab's k'-th best is no bound)
    // filter push-down (optional): q_filter[q] = filter index or kNoFilter; filter_masks[f][row / 32] bit (row % 32)
    const uint32_t* q_filter;      // [qpad] or nullptr (no filtered query in this batch)
    const uint32_t* filter_masks;  // [n_filters][filter_words]
    uint32_t filter_words;
    uint32_t* slab_progress;  // optional [n_slabs][n_qgroups]: tiles issued by each work item (0 = not started,
                              // 0x7FFFFFFF = done); keeps the CTAs sharing a slab within `lead_tiles` of each other so a
                              // corpus tile is still in L2 when the last of them reads it
    uint32_t lead_tiles;      // 0 = no limit
    uint32_t debug_flags;     // bit 0: epilogue only does the barrier handshakes (diagnostics; results invalid)
    const uint32_t* group_flags;  // optional [n_qgroups]: groups whose flag is 0 are skipped (second pass)
    float* dump;              // optional [n_q][dump_ld] raw ranking scores (debug), else nullptr
    uint64_t dump_ld;
    unsigned long long* cta_times;  // optional [grid][2] %globaltimer at CTA start / end (diagnostics), else nullptr
};

// knn_tcgen05.cu
size_t fused_smem_bytes(uint32_t kprime, uint32_t stages, int ctas);
uint32_t fused_pick_stages(uint32_t kprime, size_t smem_optin, int ctas);
uint32_t fused_round_kprime(uint32_t kprime);   // k' actually kept: 16 / 32 (register lists) or as given
cudaError_t launch_fused_knn(const CUtensorMap& tmap_q, const CUtensorMap& tmap_c, const FusedParams& p,
                             uint32_t grid, cudaStream_t stream);

// pwknn_kernels.cu
cudaError_t launch_init_rows(__nv_bfloat16* rows, RowAux* aux, uint64_t row0, uint64_t row1, uint32_t dpad,
                             cudaStream_t s);
// xmax_sq: device float, raised to the largest |x|^2 written (bound of the ranking-pass noise for L2SQ / IP)
cudaError_t launch_upsert_f64(const double* src, const uint32_t* slots, uint64_t n, uint32_t dims, uint32_t dpad,
                              int metric, __nv_bfloat16* rows, RowAux* aux, ShadowStore sh, float* xmax_sq,
                              cudaStream_t s);
cudaError_t launch_upsert_f32(const float* src, const uint32_t* slots, uint64_t n, uint32_t dims, uint32_t dpad,
                              int metric, __nv_bfloat16* rows, RowAux* aux, ShadowStore sh, float* xmax_sq,
                              cudaStream_t s);
cudaError_t launch_upsert_bf16(const __nv_bfloat16* src, uint64_t src_stride, const uint32_t* slots, uint64_t n,
                               uint32_t dims, uint32_t dpad, int metric, __nv_bfloat16* rows, RowAux* aux,
                               ShadowStore sh, float* xmax_sq, cudaStream_t s);
cudaError_t launch_move_rows(const uint32_t* src_dst_pairs, uint64_t n, uint32_t dpad, __nv_bfloat16* rows,
                             RowAux* aux, ShadowStore sh, cudaStream_t s);
// q_index (optional): packed query i is row q_index[i] of src (a subset of a staged batch)
cudaError_t launch_pack_queries_f64(const double* src, const uint32_t* q_index, uint64_t nq, uint32_t dims,
                                    uint32_t dpad, uint64_t qpad, __nv_bfloat16* dst, float* gthr, cudaStream_t s);
cudaError_t launch_pack_queries_f32(const float* src, const uint32_t* q_index, uint64_t nq, uint32_t dims,
                                    uint32_t dpad, uint64_t qpad, __nv_bfloat16* dst, float* gthr, cudaStream_t s);
cudaError_t launch_pack_queries_bf16(const __nv_bfloat16* src, uint64_t src_stride, const uint32_t* q_index,
                                     uint64_t nq, uint32_t dims, uint32_t dpad, uint64_t qpad, __nv_bfloat16* dst,
                                     float* gthr, cudaStream_t s);
// merge per-slab candidates, re-rank in f64, write ascending (distance, slot) top-k.
struct MergeArgs {
    const float* part_val;       // [n_slabs][qpad][list_len]
    const uint32_t* part_idx;
    uint32_t n_slabs;
    uint64_t qpad;
    uint32_t list_len;           // entries per (slab, query) list
    uint32_t n_extract;          // candidates pulled out of the k-way merge and re-ranked in f64 (<= n_slabs*list_len)
    uint32_t n_q, k_out;
    ExactView ev;
    int metric;
    uint32_t n_valid;
    uint32_t* out_slots;         // [n_q][k_out]   (or out_packed)
    double* out_dist;
    long long* out_packed;       // [n_q][k_out][2] {slot + id_offset | -1, distance bits}
    long long id_offset;
    // optimistic first pass: a query whose merge exhausted a FULL slab list may miss candidates -> flag it
    uint32_t* incomplete_out;    // [n_q] (written 0/1) or nullptr
    uint32_t* group_flags_out;   // [ceil(n_q / group_queries)] OR-ed, or nullptr
    uint32_t group_queries;      // queries per flag group (128 * ctas)
    // second pass: only queries with incomplete_in[q] != 0 are processed
    const uint32_t* incomplete_in;
    // certainty test: uncertain_out[q] = 1 when a row outside the re-ranked candidates could still belong to the
    // top q_k[q] (or k_out when q_k is null) given the noise bound; such queries are re-answered by the select path
    NoiseBound noise;
    const uint32_t* q_k;         // optional per-query k (<= k_out)
    uint32_t* uncertain_out;     // [n_q] or nullptr
};
cudaError_t launch_merge_refine(const MergeArgs& a, cudaStream_t s);
// cross-shard merge of all-gathered packed {id, distance bits} top-k lists
cudaError_t launch_merge_shards(const long long* gathered, uint32_t world, uint32_t nq, uint32_t k, long long* out_ids,
                                double* out_dist, cudaStream_t s);
// exact path: f64 distances for a chunk of queries against all rows, then a stable segmented sort.
size_t exact_workspace_bytes(uint64_t n_rows, uint32_t q_chunk);
cudaError_t launch_exact_search(const ExactView& ev, const FilterView& fv, uint32_t n_valid, int metric, uint32_t q0,
                                uint32_t q_chunk, uint32_t k_out, void* workspace, size_t workspace_bytes,
                                uint32_t* out_slots, double* out_dist, cudaStream_t s);
// Select path (108 < k <= kSelectMaxK, and re-answering queries the certainty test flagged): `scores` is the fp32
// ranking-score dump of the fused kernel for `nq` queries ([nq][ld], rows < n_valid); per query: 2 x 12-bit radix
// histograms find (a lower bound of) the m-th largest score, every row scoring >= that minus 2 eps is gathered
// (<= kSelectCap, else overflow[q] = 1), re-scored in f64 with the reference formula, sorted by (distance, slot)
// and the first k_out emitted.  m = q_k[q] (or k_out).  Workspace: select_workspace_bytes(nq).
size_t select_workspace_bytes(uint32_t nq);
struct SelectArgs {
    const float* scores;
    uint64_t ld;
    uint32_t n_valid, nq, k_out;
    const uint32_t* q_k;         // optional per-query m
    FilterView fv;               // q_filter indexed by the sub-batch position
    ExactView ev;                // exact rows / queries for the f64 re-rank (q_index applies)
    int metric;
    NoiseBound noise;
    void* workspace;
    uint32_t* out_slots;         // [nq][k_out]
    double* out_dist;
    uint32_t* overflow_out;      // [nq]
};
cudaError_t launch_select_search(const SelectArgs& a, cudaStream_t s);
cudaError_t launch_read_rows(const __nv_bfloat16* rows, uint64_t slot0, uint64_t n, uint32_t dims, uint32_t dpad,
                             double* out, cudaStream_t s);

}  // namespace pwknn
