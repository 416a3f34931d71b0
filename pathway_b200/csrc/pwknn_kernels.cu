// Delta path (upsert / swap-remove), query packing, cross-slab merge + f64 re-rank, and the exact
// (any-k) search path.  All HBM-bound byte/row work: one warp per row, 16-byte vector accesses.
//
// Reference semantics restated here (paths relative to /root/reference):
//   add     src/external_integration/brute_force_knn_integration.rs:131-142  (row copy into slot)
//   remove  :146-170  (swap-with-last)
//   cos     :70-90    1 - dot / sqrt(|x|^2 |q|^2)      l2sq  :92-111   |x|^2 + |q|^2 - 2 dot
//   order   :220-224  ascending (OrderedFloat(distance), slot), NaN last
#include <cfloat>
#include <algorithm>
#include <cmath>
#include <cub/device/device_segmented_radix_sort.cuh>

#include "pwknn_internal.h"

namespace pwknn {

namespace {

constexpr int kWarpsPerBlock = 8;
inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ RowAux make_aux(int metric, double sq_norm) {
    RowAux a;
    if (metric == 1) {  // COS
        if (sq_norm > 0.0) {
            a.scale = static_cast<float>(1.0 / sqrt(sq_norm));
            a.bias = 0.f;
        } else {  // zero row: cosine distance is NaN in the reference -> ranked last, by slot
            a.scale = 0.f;
            a.bias = -FLT_MAX;
        }
    } else if (metric == 2) {  // IP (usearch MetricKind::IP): distance = 1 - x.q, rank by the dot product itself
        a.scale = 1.f;
        a.bias = 0.f;
    } else {  // L2SQ: rank by dot - |x|^2/2
        a.scale = 1.f;
        a.bias = static_cast<float>(-0.5 * sq_norm);
    }
    return a;
}

__global__ void init_rows_kernel(__nv_bfloat16* rows, RowAux* aux, uint64_t row0, uint64_t row1, uint32_t dpad) {
    const uint64_t vec_per_row = dpad / 8;  // uint4 = 8 bf16
    const uint64_t total = (row1 - row0) * vec_per_row;
    uint4* base = reinterpret_cast<uint4*>(rows + row0 * dpad);
    for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<uint64_t>(gridDim.x) * blockDim.x)
        base[i] = make_uint4(0, 0, 0, 0);
    for (uint64_t r = row0 + blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; r < row1;
         r += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
        RowAux a;
        a.scale = 0.f;
        a.bias = -INFINITY;
        aux[r] = a;
    }
}

template <typename T>
__device__ __forceinline__ __nv_bfloat16 to_bf16(T x);
template <>
__device__ __forceinline__ __nv_bfloat16 to_bf16<double>(double x) {
    return __float2bfloat16_rn(__double2float_rn(x));  // f64 -> f32 -> bf16, RNE twice (documented)
}
template <>
__device__ __forceinline__ __nv_bfloat16 to_bf16<float>(float x) {
    return __float2bfloat16_rn(x);
}
template <>
__device__ __forceinline__ __nv_bfloat16 to_bf16<__nv_bfloat16>(__nv_bfloat16 x) {
    return x;
}

// One warp per delta row: convert, write the padded bf16 row into its slot, compute |x|^2 in f64
// from the ROUNDED values (what the search kernels will see) and the per-row epilogue constants.
template <typename T>
__device__ __forceinline__ double to_f64(T x);
template <>
__device__ __forceinline__ double to_f64<double>(double x) { return x; }
template <>
__device__ __forceinline__ double to_f64<float>(float x) { return static_cast<double>(x); }
template <>
__device__ __forceinline__ double to_f64<__nv_bfloat16>(__nv_bfloat16 x) { return static_cast<double>(__bfloat162float(x)); }

template <typename T>
__global__ void upsert_kernel(const T* __restrict__ src, uint64_t src_stride, const uint32_t* __restrict__ slots,
                              uint64_t n, uint32_t dims, uint32_t dpad, int metric, __nv_bfloat16* rows, RowAux* aux,
                              ShadowStore sh) {
    const uint64_t w = (blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    if (w >= n) return;
    const uint32_t slot = slots[w];
    if (slot == kInvalidSlot) return;  // row skipped by the host (superseded duplicate)
    const T* in = src + w * src_stride;
    __nv_bfloat16* out = rows + static_cast<uint64_t>(slot) * dpad;
    double acc = 0.0;
    for (uint32_t d = lane; d < dpad; d += 32) {
        __nv_bfloat16 b = __float2bfloat16_rn(0.f);
        if (d < dims) {
            const T raw = in[d];
            b = to_bf16<T>(raw);
            // exact re-rank store: the unrounded value (f64 inputs into an f32 store are rounded once, RNE)
            if (sh.kind == 1) static_cast<float*>(sh.rows)[static_cast<uint64_t>(slot) * sh.stride + d] = static_cast<float>(to_f64<T>(raw));
            else if (sh.kind == 2) static_cast<double*>(sh.rows)[static_cast<uint64_t>(slot) * sh.stride + d] = to_f64<T>(raw);
        }
        out[d] = b;
        const double x = static_cast<double>(__bfloat162float(b));
        acc = fma(x, x, acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) aux[slot] = make_aux(metric, acc);
}

// (src, dst) pairs; sources are all >= the final size and destinations < it, so moves never overlap.
__global__ void move_rows_kernel(const uint32_t* __restrict__ pairs, uint64_t n, uint32_t dpad, __nv_bfloat16* rows,
                                 RowAux* aux, ShadowStore sh) {
    const uint64_t w = (blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    if (w >= n) return;
    const uint32_t s = pairs[2 * w], d = pairs[2 * w + 1];
    const uint4* in = reinterpret_cast<const uint4*>(rows + static_cast<uint64_t>(s) * dpad);
    uint4* out = reinterpret_cast<uint4*>(rows + static_cast<uint64_t>(d) * dpad);
    for (uint32_t i = lane; i < dpad / 8; i += 32) out[i] = in[i];
    if (lane == 0) aux[d] = aux[s];
    if (sh.kind) {  // rows of the exact store are 16-byte multiples (stride % 4 == 0 for f32, % 2 == 0 for f64)
        const uint32_t vecs = sh.stride * (sh.kind == 1 ? 4u : 8u) / 16u;
        const uint4* sin = reinterpret_cast<const uint4*>(static_cast<const uint8_t*>(sh.rows) + static_cast<uint64_t>(s) * vecs * 16);
        uint4* sout = reinterpret_cast<uint4*>(static_cast<uint8_t*>(sh.rows) + static_cast<uint64_t>(d) * vecs * 16);
        for (uint32_t i = lane; i < vecs; i += 32) sout[i] = sin[i];
    }
}

template <typename T>
__global__ void pack_queries_kernel(const T* __restrict__ src, uint64_t src_stride, uint64_t nq, uint32_t dims,
                                    uint32_t dpad, uint64_t qpad, __nv_bfloat16* dst, float* gthr) {
    const uint64_t total = qpad * dpad;
    for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
        const uint64_t q = i / dpad;
        const uint32_t d = static_cast<uint32_t>(i % dpad);
        __nv_bfloat16 b = __float2bfloat16_rn(0.f);
        if (q < nq && d < dims) b = to_bf16<T>(src[q * src_stride + d]);
        dst[i] = b;
        if (d == 0) gthr[q] = -INFINITY;  // reset the cross-slab threshold for this batch
    }
}

__global__ void read_rows_kernel(const __nv_bfloat16* rows, uint64_t slot0, uint64_t n, uint32_t dims, uint32_t dpad,
                                 double* out) {
    const uint64_t total = n * dims;
    for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
        const uint64_t r = i / dims;
        const uint32_t d = static_cast<uint32_t>(i % dims);
        out[i] = static_cast<double>(__bfloat162float(rows[(slot0 + r) * dpad + d]));
    }
}

// ------------------------------------------------------------------------------------------------
// kElems consecutive elements of T starting at an address aligned to kElems * sizeof(T), widened to f64, with the
// widest load that covers them (16 / 8 / 4 bytes; 32 bytes = two 16-byte loads).
template <typename T, int kElems>
__device__ __forceinline__ void load_elems_f64(const T* __restrict__ p, double (&out)[kElems]) {
    constexpr int kBytes = kElems * static_cast<int>(sizeof(T));
    static_assert(kBytes == 4 || kBytes == 8 || kBytes == 16 || kBytes == 32, "unsupported vector width");
    alignas(16) uint32_t raw[kBytes / 4];
    if constexpr (kBytes >= 16) {
#pragma unroll
        for (int v = 0; v < kBytes / 16; ++v) {
            const uint4 t = __ldg(reinterpret_cast<const uint4*>(p) + v);
            raw[4 * v] = t.x; raw[4 * v + 1] = t.y; raw[4 * v + 2] = t.z; raw[4 * v + 3] = t.w;
        }
    } else if constexpr (kBytes == 8) {
        const uint2 t = __ldg(reinterpret_cast<const uint2*>(p));
        raw[0] = t.x; raw[1] = t.y;
    } else {
        raw[0] = __ldg(reinterpret_cast<const uint32_t*>(p));
    }
    const T* e = reinterpret_cast<const T*>(raw);
#pragma unroll
    for (int i = 0; i < kElems; ++i) out[i] = to_f64<T>(e[i]);
}

// f64 dot / squared norm of one (row, query) pair by a group of LANES consecutive lanes: deterministic
// accumulation (per lane in index order, then an xor-shuffle tree).  RT / QT: bf16, float or double (see ExactView).
// Rows are read in 16-byte pieces (8 bf16 / 4 f32 / 2 f64), lane-strided, so that all of a lane's loads are in
// flight together -- the scalar 2-byte loop this replaces was load-latency bound (k = 100: ~0.5 ms per batch).
// `vec` (uniform in the group): x and q are aligned for that (checked by the caller); else the scalar loop.
template <typename RT, typename QT, int LANES>
__device__ __forceinline__ void group_dot_f64(const RT* __restrict__ x, const QT* __restrict__ q, uint32_t len,
                                              uint32_t sl, bool active, double& dot, double& nx, bool vec = false) {
    double a = 0.0, b = 0.0;
    if (active) {
        constexpr int E = 16 / static_cast<int>(sizeof(RT));
        uint32_t done = 0;
        if (vec) {
            const uint32_t chunks = len / E;
#pragma unroll 4
            for (uint32_t c = sl; c < chunks; c += LANES) {
                double xe[E], qe[E];
                load_elems_f64<RT, E>(x + c * E, xe);
                load_elems_f64<QT, E>(q + c * E, qe);
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    a = fma(xe[e], qe[e], a);
                    b = fma(xe[e], xe[e], b);
                }
            }
            done = chunks * E;
        }
        for (uint32_t i = done + sl; i < len; i += LANES) {
            const double xv = to_f64<RT>(x[i]);
            a = fma(xv, to_f64<QT>(q[i]), a);
            b = fma(xv, xv, b);
        }
    }
#pragma unroll
    for (int o = LANES / 2; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    dot = a;
    nx = b;
}

// alignment test for the 16-byte path of group_dot_f64 (rows: base and stride; query: base)
template <typename RT, typename QT>
__device__ __forceinline__ bool dot_vec_ok(const void* rows, uint64_t row_stride_elems, const void* qv) {
    constexpr uint64_t E = 16 / sizeof(RT);
    return (reinterpret_cast<uintptr_t>(rows) % 16 == 0) && ((row_stride_elems * sizeof(RT)) % 16 == 0) &&
           (reinterpret_cast<uintptr_t>(qv) % (E * sizeof(QT)) == 0);
}

__device__ __forceinline__ double distance_f64(int metric, double dot, double nx, double nq) {
    // brute_force_knn_integration.rs:87 / :108
    if (metric == 1) return 1.0 - dot / sqrt(nx * nq);
    if (metric == 2) return 1.0 - dot;  // usearch inner-product distance (usearch_integration.rs: metric passed through)
    return nx + nq - 2.0 * dot;
}

// (OrderedFloat(d), slot) ascending, NaN greatest  (rs:220-224)
__device__ __forceinline__ bool cand_less(double da, uint32_t sa, double db, uint32_t sb) {
    const bool an = isnan(da), bn = isnan(db);
    if (an != bn) return bn;
    if (!an && da != db) return da < db;
    return sa < sb;
}

// order-preserving float -> u32 (larger float -> larger key)
__device__ __forceinline__ uint32_t float_key(float v) {
    const uint32_t u = __float_as_uint(v);
    return u ^ ((u >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}

// One WARP per query (several queries per block).  The per-slab candidate lists are already sorted by
// (score desc, slot asc), so the best k' of their union come out of a k-way merge over the list heads
// (warp arg-max per extracted element) -- no full sort.  Those k' rows are then re-scored in f64 exactly
// as the reference formula does and ordered by (distance, slot), NaN last; the first k_out are emitted.
constexpr int kMergeMaxSlabsPerLane = 5;  // n_slabs <= 160

template <typename RT, typename QT>
__global__ void merge_refine_kernel(const MergeArgs a, uint32_t warp_smem_bytes) {
    extern __shared__ __align__(16) uint8_t sm[];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t q = blockIdx.x * (blockDim.x >> 5) + warp;
    if (q >= a.n_q) return;
    if (a.incomplete_in && a.incomplete_in[q] == 0) return;  // second pass: this query was already complete
    const uint32_t list_len = a.list_len, n_slabs = a.n_slabs, n_extract = a.n_extract, k_out = a.k_out;
    const uint32_t m = n_slabs * list_len;
    const uint32_t lstride = list_len | 1u;  // odd stride (in 8-byte words): the 32 list heads hit distinct banks
    uint8_t* base = sm + static_cast<size_t>(warp) * warp_smem_bytes;
    double* cd = reinterpret_cast<double*>(base);                                      // [n_extract]
    unsigned long long* lists = reinterpret_cast<unsigned long long*>(base + static_cast<size_t>(n_extract) * 8);
    uint32_t* cs = reinterpret_cast<uint32_t*>(base + static_cast<size_t>(n_extract) * 8 +
                                               static_cast<size_t>(n_slabs) * lstride * 8);

    // stage all candidate lists of this query as sortable 64-bit keys: (score key << 32) | ~slot
    for (uint32_t i = lane; i < m; i += 32) {
        const uint32_t slab = i / list_len, j = i % list_len;
        const uint64_t o = (static_cast<uint64_t>(slab) * a.qpad + q) * list_len + j;
        const uint32_t slot = a.part_idx[o];
        unsigned long long key = 0;  // invalid entries sort last
        if (slot != kInvalidSlot && slot < a.n_valid)
            key = (static_cast<unsigned long long>(float_key(a.part_val[o])) << 32) | (0xFFFFFFFFu - slot);
        lists[slab * lstride + j] = key;
    }
    __syncwarp();
    uint32_t head[kMergeMaxSlabsPerLane];
#pragma unroll
    for (int t = 0; t < kMergeMaxSlabsPerLane; ++t) head[t] = 0;
    uint32_t ncand = 0;
    for (uint32_t out = 0; out < n_extract; ++out) {
        unsigned long long best = 0;
        int best_t = -1;
#pragma unroll
        for (int t = 0; t < kMergeMaxSlabsPerLane; ++t) {
            const uint32_t slab = lane + 32u * t;
            if (slab < n_slabs && head[t] < list_len) {
                const unsigned long long key = lists[slab * lstride + head[t]];
                if (key > best) { best = key; best_t = t; }
            }
        }
        unsigned long long wmax = best;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long other = __shfl_xor_sync(0xffffffffu, wmax, o);
            wmax = other > wmax ? other : wmax;
        }
        if (wmax == 0) break;  // every list exhausted
        if (best == wmax) {     // unique winner: slots are distinct
#pragma unroll
            for (int t = 0; t < kMergeMaxSlabsPerLane; ++t)
                if (t == best_t) head[t]++;
            cs[out] = 0xFFFFFFFFu - static_cast<uint32_t>(wmax & 0xFFFFFFFFull);
        }
        ncand = out + 1;
    }
    if (a.incomplete_out) {
        // A slab whose FULL list was entirely consumed may hold further rows that outrank what is left elsewhere.
        bool exhausted = false;
#pragma unroll
        for (int t = 0; t < kMergeMaxSlabsPerLane; ++t) {
            const uint32_t slab = lane + 32u * t;
            if (slab < n_slabs && head[t] == list_len && lists[slab * lstride + list_len - 1] != 0) exhausted = true;
        }
        const bool any = __any_sync(0xffffffffu, exhausted);
        if (lane == 0) {
            a.incomplete_out[q] = any ? 1u : 0u;
            if (any && a.group_flags_out) atomicOr(a.group_flags_out + q / a.group_queries, 1u);
        }
    }
    __syncwarp();
    const RT* rows = static_cast<const RT*>(a.ev.rows);
    const QT* qv = static_cast<const QT*>(a.ev.queries) + static_cast<uint64_t>(q) * a.ev.q_stride;
    double d0, nq;
    group_dot_f64<QT, QT, 32>(qv, qv, a.ev.len, lane, true, d0, nq);
    // f64 re-rank: 4 candidates at a time, 8 lanes each, 16-byte row loads
    const bool vec = dot_vec_ok<RT, QT>(rows, a.ev.row_stride, qv);
    const uint32_t sub = lane >> 3, sl = lane & 7;
    for (uint32_t c0 = 0; c0 < ncand; c0 += 4) {
        const uint32_t c = c0 + sub;
        const bool active = c < ncand;
        const RT* x = rows + static_cast<uint64_t>(active ? cs[c] : 0u) * a.ev.row_stride;
        double dot, nx;
        group_dot_f64<RT, QT, 8>(x, qv, a.ev.len, sl, active, dot, nx, vec);
        if (active && sl == 0) cd[c] = distance_f64(a.metric, dot, nx, nq);
    }
    __syncwarp();
    for (uint32_t c = lane; c < ncand; c += 32) {
        const double d = cd[c];
        const uint32_t s = cs[c];
        uint32_t rank = 0;
        for (uint32_t o = 0; o < ncand; ++o) rank += cand_less(cd[o], cs[o], d, s) ? 1u : 0u;
        if (rank < k_out) {
            const uint64_t o = static_cast<uint64_t>(q) * k_out + rank;
            if (a.out_packed) {  // {global id, distance bits}: the payload of the cross-shard all-gather
                a.out_packed[2 * o] = static_cast<long long>(s) + a.id_offset;
                a.out_packed[2 * o + 1] = __double_as_longlong(d);
            } else {
                a.out_slots[o] = s;
                a.out_dist[o] = d;
            }
        }
    }
    for (uint32_t r = ncand + lane; r < k_out; r += 32) {
        const uint64_t o = static_cast<uint64_t>(q) * k_out + r;
        if (a.out_packed) {
            a.out_packed[2 * o] = -1;
            a.out_packed[2 * o + 1] = __double_as_longlong(static_cast<double>(INFINITY));
        } else {
            a.out_slots[o] = kInvalidSlot;
            a.out_dist[o] = INFINITY;
        }
    }
}

// Cross-shard k-way merge after the all-gather: gathered [world][nq][k][2] (id, distance bits), every shard's
// k entries ascending; output the k best by (distance, shard, position) -- NaN last, id -1 = padding last.
__global__ void __launch_bounds__(128)
merge_shards_kernel(const long long* __restrict__ gathered, uint32_t world, uint32_t nq, uint32_t k,
                    long long* out_ids, double* out_dist) {
    extern __shared__ uint8_t sm[];
    double* cd = reinterpret_cast<double*>(sm);
    long long* ci = reinterpret_cast<long long*>(sm + static_cast<size_t>(world) * k * 8);
    const uint32_t q = blockIdx.x, tid = threadIdx.x, m = world * k;
    for (uint32_t c = tid; c < m; c += blockDim.x) {
        const uint32_t w = c / k, r = c % k;
        const uint64_t o = ((static_cast<uint64_t>(w) * nq + q) * k + r) * 2;
        ci[c] = gathered[o];
        cd[c] = __longlong_as_double(gathered[o + 1]);
    }
    __syncthreads();
    for (uint32_t c = tid; c < m; c += blockDim.x) {
        const double d = cd[c];
        const bool pad = ci[c] < 0, nan = isnan(d);
        uint32_t rank = 0;
        for (uint32_t o = 0; o < m; ++o) {
            const double e = cd[o];
            const bool opad = ci[o] < 0, onan = isnan(e);
            bool before;
            if (opad != pad) before = pad;                  // real entries before padding
            else if (onan != nan) before = nan;             // numbers before NaN
            else if (!nan && e != d) before = e < d;
            else before = o < c;                            // (shard, position): stable
            rank += before ? 1u : 0u;
        }
        if (rank < k) {
            out_ids[static_cast<uint64_t>(q) * k + rank] = ci[c];
            out_dist[static_cast<uint64_t>(q) * k + rank] = d;
        }
    }
}

// ------------------------------------------------------------------------------------------------ exact path
constexpr unsigned long long kKeyExcluded = ~0ull;      // row filtered out: sorts after everything, never emitted
constexpr unsigned long long kKeyNaN = ~0ull - 1ull;    // NaN distance: greatest real key (ordered-float)
__device__ __forceinline__ uint64_t sortable_key(double d) {
    if (isnan(d)) return kKeyNaN;
    const uint64_t u = static_cast<uint64_t>(__double_as_longlong(d + 0.0));  // -0.0 -> +0.0
    return (u >> 63) ? ~u : (u | (1ull << 63));
}
__device__ __forceinline__ double key_to_double(uint64_t k) {
    if (k == kKeyNaN) return NAN;
    const uint64_t u = (k >> 63) ? (k ^ (1ull << 63)) : ~k;
    return __longlong_as_double(static_cast<long long>(u));
}

template <typename QT>
__global__ void query_norms_kernel(const ExactView ev, uint32_t q0, uint32_t q_chunk, double* nq_out) {
    const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= q_chunk) return;
    const QT* qv = static_cast<const QT*>(ev.queries) + static_cast<uint64_t>(q0 + w) * ev.q_stride;
    double d0, nq;
    group_dot_f64<QT, QT, 32>(qv, qv, ev.len, lane, true, d0, nq);
    if (lane == 0) nq_out[w] = nq;
}

// one warp per corpus row, looping over the query chunk; keys laid out [q][row] for the segmented sort
template <typename RT, typename QT>
__global__ void exact_dist_kernel(const ExactView ev, const FilterView fv, uint32_t n_valid, int metric, uint32_t q0,
                                  uint32_t q_chunk, const double* __restrict__ nq, uint64_t* keys, uint32_t* vals,
                                  int* offsets) {
    const uint64_t w = (blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    if (w == 0 && lane == 0)
        for (uint32_t i = 0; i <= q_chunk; ++i) offsets[i] = static_cast<int>(static_cast<uint64_t>(i) * n_valid);
    if (w >= n_valid) return;
    const RT* x = static_cast<const RT*>(ev.rows) + w * ev.row_stride;
    for (uint32_t qi = 0; qi < q_chunk; ++qi) {
        const QT* qv = static_cast<const QT*>(ev.queries) + static_cast<uint64_t>(q0 + qi) * ev.q_stride;
        double dot, nx;
        group_dot_f64<RT, QT, 32>(x, qv, ev.len, lane, true, dot, nx);
        if (lane == 0) {
            uint64_t key = sortable_key(distance_f64(metric, dot, nx, nq[qi]));
            if (fv.q_filter) {
                const uint32_t fid = fv.q_filter[q0 + qi];
                if (fid != kNoFilter && !((fv.masks[static_cast<uint64_t>(fid) * fv.words + (w >> 5)] >> (w & 31)) & 1u))
                    key = kKeyExcluded;
            }
            keys[static_cast<uint64_t>(qi) * n_valid + w] = key;
            vals[static_cast<uint64_t>(qi) * n_valid + w] = static_cast<uint32_t>(w);
        }
    }
}

__global__ void exact_emit_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals,
                                  uint32_t n_valid, uint32_t q0, uint32_t q_chunk, uint32_t k_out, uint32_t* out_slots,
                                  double* out_dist) {
    const uint64_t total = static_cast<uint64_t>(q_chunk) * k_out;
    for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
        const uint32_t qi = static_cast<uint32_t>(i / k_out), r = static_cast<uint32_t>(i % k_out);
        const uint64_t o = static_cast<uint64_t>(q0 + qi) * k_out + r;
        if (r < n_valid && keys[static_cast<uint64_t>(qi) * n_valid + r] != kKeyExcluded) {
            out_slots[o] = vals[static_cast<uint64_t>(qi) * n_valid + r];
            out_dist[o] = key_to_double(keys[static_cast<uint64_t>(qi) * n_valid + r]);
        } else {
            out_slots[o] = kInvalidSlot;
            out_dist[o] = INFINITY;
        }
    }
}

inline uint32_t blocks_for(uint64_t work_items, uint32_t per_block, uint32_t cap = 148 * 16) {
    uint64_t b = (work_items + per_block - 1) / per_block;
    if (b < 1) b = 1;
    if (b > cap) b = cap;
    return static_cast<uint32_t>(b);
}
struct ExactLayout {
    size_t keys_in, keys_out, vals_in, vals_out, nq, offsets, temp, temp_bytes, total;
};
ExactLayout exact_layout(uint64_t n_rows, uint32_t q_chunk) {
    ExactLayout L{};
    const size_t items = static_cast<size_t>(n_rows) * q_chunk;
    size_t off = 0;
    L.keys_in = off; off = align_up(off + items * 8, 256);
    L.keys_out = off; off = align_up(off + items * 8, 256);
    L.vals_in = off; off = align_up(off + items * 4, 256);
    L.vals_out = off; off = align_up(off + items * 4, 256);
    L.nq = off; off = align_up(off + static_cast<size_t>(q_chunk) * 8, 256);
    L.offsets = off; off = align_up(off + (static_cast<size_t>(q_chunk) + 1) * 4, 256);
    size_t temp = 0;
    cub::DeviceSegmentedRadixSort::SortPairs(nullptr, temp, static_cast<const uint64_t*>(nullptr),
                                             static_cast<uint64_t*>(nullptr), static_cast<const uint32_t*>(nullptr),
                                             static_cast<uint32_t*>(nullptr), static_cast<int>(items),
                                             static_cast<int>(q_chunk), static_cast<const int*>(nullptr),
                                             static_cast<const int*>(nullptr));
    L.temp = off; L.temp_bytes = temp; off = align_up(off + temp, 256);
    L.total = off;
    return L;
}

}  // namespace

cudaError_t launch_init_rows(__nv_bfloat16* rows, RowAux* aux, uint64_t row0, uint64_t row1, uint32_t dpad,
                             cudaStream_t s) {
    if (row1 <= row0) return cudaSuccess;
    init_rows_kernel<<<blocks_for((row1 - row0) * (dpad / 8), 256), 256, 0, s>>>(rows, aux, row0, row1, dpad);
    return cudaGetLastError();
}

template <typename T>
static cudaError_t launch_upsert_t(const T* src, uint64_t stride, const uint32_t* slots, uint64_t n, uint32_t dims,
                                   uint32_t dpad, int metric, __nv_bfloat16* rows, RowAux* aux, ShadowStore sh,
                                   cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    const uint64_t blocks = (n + kWarpsPerBlock - 1) / kWarpsPerBlock;
    upsert_kernel<T><<<static_cast<uint32_t>(blocks), kWarpsPerBlock * 32, 0, s>>>(src, stride, slots, n, dims, dpad,
                                                                                   metric, rows, aux, sh);
    return cudaGetLastError();
}
cudaError_t launch_upsert_f64(const double* src, const uint32_t* slots, uint64_t n, uint32_t dims, uint32_t dpad,
                              int metric, __nv_bfloat16* rows, RowAux* aux, ShadowStore sh, cudaStream_t s) {
    return launch_upsert_t<double>(src, dims, slots, n, dims, dpad, metric, rows, aux, sh, s);
}
cudaError_t launch_upsert_f32(const float* src, const uint32_t* slots, uint64_t n, uint32_t dims, uint32_t dpad,
                              int metric, __nv_bfloat16* rows, RowAux* aux, ShadowStore sh, cudaStream_t s) {
    return launch_upsert_t<float>(src, dims, slots, n, dims, dpad, metric, rows, aux, sh, s);
}
cudaError_t launch_upsert_bf16(const __nv_bfloat16* src, uint64_t src_stride, const uint32_t* slots, uint64_t n,
                               uint32_t dims, uint32_t dpad, int metric, __nv_bfloat16* rows, RowAux* aux,
                               ShadowStore sh, cudaStream_t s) {
    return launch_upsert_t<__nv_bfloat16>(src, src_stride, slots, n, dims, dpad, metric, rows, aux, sh, s);
}

cudaError_t launch_move_rows(const uint32_t* pairs, uint64_t n, uint32_t dpad, __nv_bfloat16* rows, RowAux* aux,
                             ShadowStore sh, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    const uint64_t blocks = (n + kWarpsPerBlock - 1) / kWarpsPerBlock;
    move_rows_kernel<<<static_cast<uint32_t>(blocks), kWarpsPerBlock * 32, 0, s>>>(pairs, n, dpad, rows, aux, sh);
    return cudaGetLastError();
}

cudaError_t launch_pack_queries_f64(const double* src, uint64_t nq, uint32_t dims, uint32_t dpad, uint64_t qpad,
                                    __nv_bfloat16* dst, float* gthr, cudaStream_t s) {
    pack_queries_kernel<double><<<blocks_for(qpad * dpad, 256), 256, 0, s>>>(src, dims, nq, dims, dpad, qpad, dst, gthr);
    return cudaGetLastError();
}
cudaError_t launch_pack_queries_f32(const float* src, uint64_t nq, uint32_t dims, uint32_t dpad, uint64_t qpad,
                                    __nv_bfloat16* dst, float* gthr, cudaStream_t s) {
    pack_queries_kernel<float><<<blocks_for(qpad * dpad, 256), 256, 0, s>>>(src, dims, nq, dims, dpad, qpad, dst, gthr);
    return cudaGetLastError();
}
cudaError_t launch_pack_queries_bf16(const __nv_bfloat16* src, uint64_t src_stride, uint64_t nq, uint32_t dims,
                                     uint32_t dpad, uint64_t qpad, __nv_bfloat16* dst, float* gthr, cudaStream_t s) {
    pack_queries_kernel<__nv_bfloat16><<<blocks_for(qpad * dpad, 256), 256, 0, s>>>(src, src_stride, nq, dims, dpad,
                                                                                    qpad, dst, gthr);
    return cudaGetLastError();
}

cudaError_t launch_merge_refine(const MergeArgs& a, cudaStream_t s) {
    if (a.n_q == 0) return cudaSuccess;
    if (a.n_slabs > 32u * kMergeMaxSlabsPerLane) return cudaErrorInvalidValue;
    const size_t per_warp = align_up(static_cast<size_t>(a.n_extract) * 8 +
                                         static_cast<size_t>(a.n_slabs) * (a.list_len | 1u) * 8 +
                                         static_cast<size_t>(a.n_extract) * 4, 16);
    uint32_t warps = static_cast<uint32_t>(std::min<size_t>(4, std::max<size_t>(1, (48 * 1024) / per_warp)));
    const size_t smem = per_warp * warps;
    const dim3 grid((a.n_q + warps - 1) / warps), block(warps * 32);
    const uint32_t pw = static_cast<uint32_t>(per_warp);
#define PWKNN_MERGE_CASE(RT, QT)                                                                                   \
    {                                                                                                              \
        cudaError_t e = cudaFuncSetAttribute(merge_refine_kernel<RT, QT>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                             static_cast<int>(smem));                                              \
        if (e != cudaSuccess) return e;                                                                            \
        merge_refine_kernel<RT, QT><<<grid, block, smem, s>>>(a, pw);                                              \
        return cudaGetLastError();                                                                                 \
    }
    const uint32_t combo = a.ev.row_kind * 3 + a.ev.q_kind;
    switch (combo) {
        case 0: PWKNN_MERGE_CASE(__nv_bfloat16, __nv_bfloat16)
        case 3: PWKNN_MERGE_CASE(float, __nv_bfloat16)
        case 4: PWKNN_MERGE_CASE(float, float)
        case 5: PWKNN_MERGE_CASE(float, double)
        case 6: PWKNN_MERGE_CASE(double, __nv_bfloat16)
        case 7: PWKNN_MERGE_CASE(double, float)
        case 8: PWKNN_MERGE_CASE(double, double)
        default: return cudaErrorInvalidValue;  // bf16 corpus is only ever re-ranked with the packed bf16 queries
    }
#undef PWKNN_MERGE_CASE
}

cudaError_t launch_merge_shards(const long long* gathered, uint32_t world, uint32_t nq, uint32_t k, long long* out_ids,
                                double* out_dist, cudaStream_t s) {
    if (nq == 0 || k == 0) return cudaSuccess;
    const size_t smem = static_cast<size_t>(world) * k * 16;
    cudaError_t e = cudaFuncSetAttribute(merge_shards_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    merge_shards_kernel<<<nq, 128, smem, s>>>(gathered, world, nq, k, out_ids, out_dist);
    return cudaGetLastError();
}

size_t exact_workspace_bytes(uint64_t n_rows, uint32_t q_chunk) { return exact_layout(n_rows, q_chunk).total; }

template <typename RT, typename QT>
static cudaError_t exact_dist_launch(const ExactView& ev, const FilterView& fv, uint32_t n_valid, int metric, uint32_t q0,
                                     uint32_t q_chunk, double* nq, uint64_t* keys_in, uint32_t* vals_in, int* offsets,
                                     cudaStream_t s) {
    query_norms_kernel<QT><<<(q_chunk + kWarpsPerBlock - 1) / kWarpsPerBlock, kWarpsPerBlock * 32, 0, s>>>(ev, q0, q_chunk, nq);
    const uint64_t blocks = (static_cast<uint64_t>(n_valid) + kWarpsPerBlock - 1) / kWarpsPerBlock;
    exact_dist_kernel<RT, QT><<<static_cast<uint32_t>(blocks), kWarpsPerBlock * 32, 0, s>>>(ev, fv, n_valid, metric, q0,
                                                                                           q_chunk, nq, keys_in, vals_in,
                                                                                           offsets);
    return cudaGetLastError();
}

cudaError_t launch_exact_search(const ExactView& ev, const FilterView& fv, uint32_t n_valid, int metric, uint32_t q0,
                                uint32_t q_chunk, uint32_t k_out, void* workspace, size_t workspace_bytes,
                                uint32_t* out_slots, double* out_dist, cudaStream_t s) {
    const ExactLayout L = exact_layout(n_valid, q_chunk);
    if (L.total > workspace_bytes) return cudaErrorInvalidValue;
    uint8_t* ws = static_cast<uint8_t*>(workspace);
    uint64_t* keys_in = reinterpret_cast<uint64_t*>(ws + L.keys_in);
    uint64_t* keys_out = reinterpret_cast<uint64_t*>(ws + L.keys_out);
    uint32_t* vals_in = reinterpret_cast<uint32_t*>(ws + L.vals_in);
    uint32_t* vals_out = reinterpret_cast<uint32_t*>(ws + L.vals_out);
    double* nq = reinterpret_cast<double*>(ws + L.nq);
    int* offsets = reinterpret_cast<int*>(ws + L.offsets);
    cudaError_t e;
    switch (ev.row_kind * 3 + ev.q_kind) {
        case 0: e = exact_dist_launch<__nv_bfloat16, __nv_bfloat16>(ev, fv, n_valid, metric, q0, q_chunk, nq, keys_in, vals_in, offsets, s); break;
        case 3: e = exact_dist_launch<float, __nv_bfloat16>(ev, fv, n_valid, metric, q0, q_chunk, nq, keys_in, vals_in, offsets, s); break;
        case 4: e = exact_dist_launch<float, float>(ev, fv, n_valid, metric, q0, q_chunk, nq, keys_in, vals_in, offsets, s); break;
        case 5: e = exact_dist_launch<float, double>(ev, fv, n_valid, metric, q0, q_chunk, nq, keys_in, vals_in, offsets, s); break;
        case 6: e = exact_dist_launch<double, __nv_bfloat16>(ev, fv, n_valid, metric, q0, q_chunk, nq, keys_in, vals_in, offsets, s); break;
        case 7: e = exact_dist_launch<double, float>(ev, fv, n_valid, metric, q0, q_chunk, nq, keys_in, vals_in, offsets, s); break;
        case 8: e = exact_dist_launch<double, double>(ev, fv, n_valid, metric, q0, q_chunk, nq, keys_in, vals_in, offsets, s); break;
        default: return cudaErrorInvalidValue;
    }
    if (e != cudaSuccess) return e;
    size_t temp = L.temp_bytes;
    // LSD radix sort is stable: equal distances keep ascending slot order == (OrderedFloat(d), slot)
    e = cub::DeviceSegmentedRadixSort::SortPairs(ws + L.temp, temp, keys_in, keys_out, vals_in, vals_out,
                                                 static_cast<int>(static_cast<uint64_t>(n_valid) * q_chunk),
                                                 static_cast<int>(q_chunk), offsets, offsets + 1, 0, 64, s);
    if (e != cudaSuccess) return e;
    exact_emit_kernel<<<blocks_for(static_cast<uint64_t>(q_chunk) * k_out, 256), 256, 0, s>>>(
        keys_out, vals_out, n_valid, q0, q_chunk, k_out, out_slots, out_dist);
    return cudaGetLastError();
}

cudaError_t launch_read_rows(const __nv_bfloat16* rows, uint64_t slot0, uint64_t n, uint32_t dims, uint32_t dpad,
                             double* out, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    read_rows_kernel<<<blocks_for(n * dims, 256), 256, 0, s>>>(rows, slot0, n, dims, dpad, out);
    return cudaGetLastError();
}

}  // namespace pwknn
