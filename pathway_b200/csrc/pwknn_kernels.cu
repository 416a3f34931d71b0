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
                              ShadowStore sh, float* xmax_sq) {
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
    if (lane == 0) {
        aux[slot] = make_aux(metric, acc);
        // largest row norm ever stored (monotone; non-negative floats order like their bit patterns): the noise
        // bound of the ranking pass scales with it for L2SQ / IP
        if (xmax_sq) atomicMax(reinterpret_cast<int*>(xmax_sq), __float_as_int(__double2float_ru(acc)));
    }
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
__global__ void pack_queries_kernel(const T* __restrict__ src, uint64_t src_stride, const uint32_t* __restrict__ q_index,
                                    uint64_t nq, uint32_t dims, uint32_t dpad, uint64_t qpad, __nv_bfloat16* dst,
                                    float* gthr) {
    const uint64_t total = qpad * dpad;
    for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
        const uint64_t q = i / dpad;
        const uint32_t d = static_cast<uint32_t>(i % dpad);
        __nv_bfloat16 b = __float2bfloat16_rn(0.f);
        if (q < nq && d < dims) b = to_bf16<T>(src[(q_index ? q_index[q] : q) * src_stride + d]);
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
// kBytes (4 / 8 / 16 / 32) raw bytes from an address aligned to min(kBytes, 16), as 32-bit words.  Loading raw
// words first and converting later keeps a whole batch of loads in flight (see group_dot_f64).
template <int kBytes>
struct RawVec { uint32_t w[kBytes / 4]; };

template <int kBytes>
__device__ __forceinline__ RawVec<kBytes> load_raw(const void* p) {
    static_assert(kBytes == 4 || kBytes == 8 || kBytes == 16 || kBytes == 32, "unsupported vector width");
    RawVec<kBytes> r;
    if constexpr (kBytes >= 16) {
#pragma unroll
        for (int v = 0; v < kBytes / 16; ++v) {
            const uint4 t = __ldg(static_cast<const uint4*>(p) + v);
            r.w[4 * v] = t.x; r.w[4 * v + 1] = t.y; r.w[4 * v + 2] = t.z; r.w[4 * v + 3] = t.w;
        }
    } else if constexpr (kBytes == 8) {
        const uint2 t = __ldg(static_cast<const uint2*>(p));
        r.w[0] = t.x; r.w[1] = t.y;
    } else {
        r.w[0] = __ldg(static_cast<const uint32_t*>(p));
    }
    return r;
}

template <typename T>
__device__ __forceinline__ double raw_elem_f64(const uint32_t* w, int i);
template <>
__device__ __forceinline__ double raw_elem_f64<__nv_bfloat16>(const uint32_t* w, int i) {
    const uint32_t bits = (i & 1) ? (w[i >> 1] & 0xFFFF0000u) : (w[i >> 1] << 16);  // bf16 -> f32 is a shift
    return static_cast<double>(__uint_as_float(bits));
}
template <>
__device__ __forceinline__ double raw_elem_f64<float>(const uint32_t* w, int i) {
    return static_cast<double>(__uint_as_float(w[i]));
}
template <>
__device__ __forceinline__ double raw_elem_f64<double>(const uint32_t* w, int i) {
    return __hiloint2double(static_cast<int>(w[2 * i + 1]), static_cast<int>(w[2 * i]));
}

// f64 dot / squared norm of one (row, query) pair by a group of LANES consecutive lanes: deterministic
// accumulation (per lane in index order, then an xor-shuffle tree).  RT / QT: bf16, float or double (see ExactView).
// Rows are read in 16-byte pieces (8 bf16 / 4 f32 / 2 f64), lane-strided, kBatch pieces per lane in flight at a time
// -- the scalar 2-byte loop this replaces was load-latency bound (k = 100: ~0.5 ms per batch).
// `vec` (uniform in the group): x and q are aligned for that (checked by the caller); else the scalar loop.
template <typename RT, typename QT, int LANES>
__device__ __forceinline__ void group_dot_f64(const RT* __restrict__ x, const QT* __restrict__ q, uint32_t len,
                                              uint32_t sl, bool active, double& dot, double& nx, bool vec = false) {
    double a = 0.0, b = 0.0;
    if (active) {
        constexpr int E = 16 / static_cast<int>(sizeof(RT));
        constexpr int QB = E * static_cast<int>(sizeof(QT));
        constexpr int kBatch = (LANES <= 8 && QB <= 16) ? 12 : 6;  // pieces per lane in flight
        uint32_t done = 0;
        if (vec) {
            const uint32_t chunks = len / E;
            for (uint32_t c0 = sl; c0 < chunks; c0 += LANES * kBatch) {
                RawVec<16> xr[kBatch];
                RawVec<QB> qr[kBatch];
#pragma unroll
                for (int u = 0; u < kBatch; ++u) {
                    const uint32_t c = c0 + u * LANES;
                    if (c < chunks) {
                        xr[u] = load_raw<16>(x + c * E);
                        qr[u] = load_raw<QB>(q + c * E);
                    }
                }
#pragma unroll
                for (int u = 0; u < kBatch; ++u) {
                    if (c0 + u * LANES < chunks) {
#pragma unroll
                        for (int e = 0; e < E; ++e) {
                            const double xv = raw_elem_f64<RT>(xr[u].w, e);
                            a = fma(xv, raw_elem_f64<QT>(qr[u].w, e), a);
                            b = fma(xv, xv, b);
                        }
                    }
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

__device__ __forceinline__ float key_to_float(uint32_t k) {
    return __uint_as_float(k ^ ((k >> 31) ? 0x80000000u : 0xFFFFFFFFu));
}
// the ranking score the tensor-core pass approximates (pwknn_internal.h RowAux), recovered exactly from the f64
// distance: COS (1 - d) |q| = dot/|x|;  L2SQ (|q|^2 - d) / 2 = dot - |x|^2/2;  IP 1 - d = dot
__device__ __forceinline__ double score_from_distance(int metric, double d, double nq, double sqrt_nq) {
    if (metric == 1) return (1.0 - d) * sqrt_nq;
    if (metric == 2) return 1.0 - d;
    return 0.5 * (nq - d);
}
__device__ __forceinline__ double noise_eps(const NoiseBound& nb, double nq) {
    const double xm2 = (nb.xmax_sq && (nb.a1 != 0.f || nb.b != 0.f)) ? static_cast<double>(*nb.xmax_sq) : 0.0;
    const double xm = nb.a1 != 0.f ? sqrt(xm2) : 0.0;
    return (static_cast<double>(nb.a0) + static_cast<double>(nb.a1) * xm) * sqrt(nq) + static_cast<double>(nb.b) * xm2;
}

// One WARP per query (several queries per block).  The best k' of the union of the per-slab candidate lists are
// the k' largest (score, ~slot) keys: a bitwise threshold search over the staged keys, no sort (their order is
// irrelevant here).  Those k' rows are then re-scored in f64 exactly as the reference formula does and ordered by
// (distance, slot), NaN last; the first k_out are emitted.
// (<= 4 warps per block; 4 blocks per SM must stay resident -- 1024 queries are 342 blocks of 3 warps on 148 SMs, and a
// second wave doubles the kernel: cap the registers at 128)
template <typename RT, typename QT>
__global__ void __launch_bounds__(128, 4) merge_refine_kernel(const MergeArgs a, uint32_t warp_smem_bytes) {
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
    float* cf = reinterpret_cast<float*>(cs + n_extract);  // fp32 ranking score of each candidate (certainty test)

    // stage all candidate lists of this query as sortable 64-bit keys: (score key << 32) | ~slot; 0 = invalid
    // (also the odd-stride padding word of every list)
#pragma unroll 8
    for (uint32_t i = lane; i < m; i += 32) {
        const uint32_t slab = i / list_len, j = i % list_len;
        const uint64_t o = (static_cast<uint64_t>(slab) * a.qpad + q) * list_len + j;
        const uint32_t slot = a.part_idx[o];
        unsigned long long key = 0;
        if (slot != kInvalidSlot && slot < a.n_valid)
            key = (static_cast<unsigned long long>(float_key(a.part_val[o])) << 32) | (0xFFFFFFFFu - slot);
        lists[slab * lstride + j] = key;
    }
    if (lstride != list_len)
        for (uint32_t slab = lane; slab < n_slabs; slab += 32) lists[slab * lstride + list_len] = 0;
    __syncwarp();
    // The n_extract best of the union = the n_extract largest keys (valid keys are distinct: the slot is part of the
    // key).  Their order does not matter here -- the f64 re-rank below orders the result -- so instead of a serial
    // k-way merge (one warp arg-max per extracted element: 50 us of a 165 us launch at k = 100) find the
    // n_extract-th largest key T by a bitwise search (64 counting passes over the staged keys) and take every key >= T.
    const uint32_t n_words = n_slabs * lstride;
    unsigned long long thr = 0;
    for (int bit = 63; bit >= 0; --bit) {
        const unsigned long long cand = thr | (1ull << bit);
        uint32_t cnt = 0;
        for (uint32_t i = lane; i < n_words; i += 32) cnt += lists[i] >= cand ? 1u : 0u;
        cnt = __reduce_add_sync(0xffffffffu, cnt);
        if (cnt >= n_extract) thr = cand;
    }
    if (thr == 0) thr = 1;  // fewer than n_extract valid keys: take them all
    uint32_t ncand = 0;
    for (uint32_t i0 = 0; i0 < n_words; i0 += 32) {
        const uint32_t i = i0 + lane;
        const unsigned long long key = i < n_words ? lists[i] : 0ull;
        const bool take = key >= thr;
        const uint32_t mask = __ballot_sync(0xffffffffu, take);
        if (take) {
            const uint32_t pos = ncand + __popc(mask & ((1u << lane) - 1u));
            cs[pos] = 0xFFFFFFFFu - static_cast<uint32_t>(key & 0xFFFFFFFFull);
            cf[pos] = key_to_float(static_cast<uint32_t>(key >> 32));
        }
        ncand += __popc(mask);
    }
    if (a.incomplete_out) {
        // A slab whose FULL list was entirely taken may hold further rows that outrank what was taken elsewhere
        // (the lists are sorted, so "entirely taken" == its last entry is valid and >= T).
        bool exhausted = false;
        for (uint32_t slab = lane; slab < n_slabs; slab += 32) exhausted |= lists[slab * lstride + list_len - 1] >= thr;
        const bool any = __any_sync(0xffffffffu, exhausted);
        if (lane == 0) {
            a.incomplete_out[q] = any ? 1u : 0u;
            if (any && a.group_flags_out) atomicOr(a.group_flags_out + q / a.group_queries, 1u);
        }
    }
    __syncwarp();
    {   // the selected rows are re-scored in f64 below: start pulling them into L2 now
        const uint32_t row_bytes = a.ev.len * static_cast<uint32_t>(sizeof(RT));
        const uint64_t row_stride_bytes = static_cast<uint64_t>(a.ev.row_stride) * sizeof(RT);
        const uint32_t lines = (row_bytes + 127u) / 128u;
        for (uint32_t i = lane; i < ncand * lines; i += 32) {
            const char* row = reinterpret_cast<const char*>(a.ev.rows) + static_cast<uint64_t>(cs[i / lines]) * row_stride_bytes;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(row + static_cast<uint64_t>(i % lines) * 128u));
        }
    }
    const RT* rows = static_cast<const RT*>(a.ev.rows);
    const QT* qv = static_cast<const QT*>(a.ev.queries) +
                   static_cast<uint64_t>(a.ev.q_index ? a.ev.q_index[q] : q) * a.ev.q_stride;
    double d0, nq;
    group_dot_f64<QT, QT, 32>(qv, qv, a.ev.len, lane, true, d0, nq, dot_vec_ok<QT, QT>(qv, 0, qv));
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
    // rank = number of candidates that precede mine in (OrderedFloat(distance), slot): one pass over the candidate
    // set against this lane's (up to 4) candidates held in registers
    constexpr int kMine = 4;  // n_extract <= 128
    double dm[kMine];
    uint32_t smn[kMine], rank[kMine];
#pragma unroll
    for (int j = 0; j < kMine; ++j) {
        const uint32_t c = lane + 32u * j;
        dm[j] = c < ncand ? cd[c] : 0.0;
        smn[j] = c < ncand ? cs[c] : 0u;
        rank[j] = 0;
    }
    for (uint32_t o = 0; o < ncand; ++o) {
        const double d = cd[o];
        const uint32_t sl2 = cs[o];
#pragma unroll
        for (int j = 0; j < kMine; ++j) rank[j] += cand_less(d, sl2, dm[j], smn[j]) ? 1u : 0u;
    }
    if (a.uncertain_out) {
        // Certainty test.  Every row that was NOT re-ranked has fp32 score <= s_min (the smallest extracted one), so its
        // exact score is <= s_min + eps; the answer is certain when the kq-th exact score lies above that.  Also certain:
        // nothing was left out (fewer valid rows than n_extract), the leftovers are zero rows (COS: score -FLT_MAX, NaN
        // distance, ranked last by slot, which the lists preserve) or the query itself is the zero vector (all NaN).
        const uint32_t kq = a.q_k ? min(a.q_k[q], k_out) : k_out;
        bool uncertain = false;
        if (kq > 0 && ncand >= n_extract && a.n_valid > n_extract && !(a.metric == 1 && nq == 0.0)) {
            const float s_min = key_to_float(static_cast<uint32_t>(thr >> 32));
            const double sqrt_nq = sqrt(nq);  // once per query: f64 square roots are slow
            double dk = 0.0, dobs = 0.0;
            bool mine = false;
#pragma unroll
            for (int j = 0; j < kMine; ++j) {
                const uint32_t c = lane + 32u * j;
                if (c < ncand) {
                    if (rank[j] == kq - 1) { dk = dm[j]; mine = true; }
                    const double t = score_from_distance(a.metric, dm[j], nq, sqrt_nq);
                    const float f = cf[c];
                    if (!isnan(t) && f > -FLT_MAX) dobs = fmax(dobs, fabs(t - static_cast<double>(f)));
                }
            }
            const uint32_t who = __ballot_sync(0xffffffffu, mine);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) dobs = fmax(dobs, __shfl_xor_sync(0xffffffffu, dobs, o));
            if (who) {
                dk = __shfl_sync(0xffffffffu, dk, __ffs(who) - 1);
                const double tk = score_from_distance(a.metric, dk, nq, sqrt_nq);
                const double eps = fmax(noise_eps(a.noise, nq), 3.0 * dobs);  // (one more sqrt, on one path)  // model bound, widened by what was observed
                uncertain = s_min > -FLT_MAX && !(tk - eps > static_cast<double>(s_min));
            }
        }
        if (lane == 0) a.uncertain_out[q] = uncertain ? 1u : 0u;
    }
#pragma unroll
    for (int j = 0; j < kMine; ++j) {
        const uint32_t c = lane + 32u * j;
        if (c < ncand && rank[j] < k_out) {
            const uint64_t o = static_cast<uint64_t>(q) * k_out + rank[j];
            if (a.out_packed) {  // {global id, distance bits}: the payload of the cross-shard all-gather
                a.out_packed[2 * o] = static_cast<long long>(smn[j]) + a.id_offset;
                a.out_packed[2 * o + 1] = __double_as_longlong(dm[j]);
            } else {
                a.out_slots[o] = smn[j];
                a.out_dist[o] = dm[j];
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
    const QT* qv = static_cast<const QT*>(ev.queries) +
                   static_cast<uint64_t>(ev.q_index ? ev.q_index[q0 + w] : q0 + w) * ev.q_stride;
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
        const QT* qv = static_cast<const QT*>(ev.queries) +
                       static_cast<uint64_t>(ev.q_index ? ev.q_index[q0 + qi] : q0 + qi) * ev.q_stride;
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

// ------------------------------------------------------------------------------------------------ select path
// k too large for the per-slab lists (and queries the certainty test flagged): the fused kernel dumps the fp32
// ranking scores [nq][ld]; per query a two-level radix histogram (2 x 12 bits of the order-preserving key) brackets the
// m-th largest score from below, every row within 2 eps of it is gathered, re-scored in f64 and sorted.
// Why 2 eps: with |fp32 - exact| <= eps, the m-th largest fp32 score F satisfies  exact_m >= F - eps, and every row of
// the exact top-m has fp32 >= exact_m - eps >= F - 2 eps.
constexpr uint32_t kSelBins = 4096;
struct SelectWs {
    uint32_t* hist;      // [nq][4096]
    uint32_t* sel_bin;   // [nq] bin chosen by pass 0 (0xFFFFFFFF: fewer than m eligible rows -> take everything)
    uint32_t* sel_rem;   // [nq] rank still to find inside that bin
    float* thr;          // [nq] gather threshold
    uint32_t* count;     // [nq] rows gathered
    double* qnorm;       // [nq] |q|^2
    uint32_t* cand;      // [nq][kSelectCap]
    uint64_t* dkey;      // [nq][kSelectCap] sortable f64 distance
    size_t total;
};
inline SelectWs select_layout(uint8_t* base, uint32_t nq) {
    SelectWs w{};
    size_t off = 0;
    auto take = [&](size_t bytes) { uint8_t* p = base ? base + off : nullptr; off = align_up(off + bytes, 256); return p; };
    w.hist = reinterpret_cast<uint32_t*>(take(static_cast<size_t>(nq) * kSelBins * 4));
    w.sel_bin = reinterpret_cast<uint32_t*>(take(static_cast<size_t>(nq) * 4));
    w.sel_rem = reinterpret_cast<uint32_t*>(take(static_cast<size_t>(nq) * 4));
    w.thr = reinterpret_cast<float*>(take(static_cast<size_t>(nq) * 4));
    w.count = reinterpret_cast<uint32_t*>(take(static_cast<size_t>(nq) * 4));
    w.qnorm = reinterpret_cast<double*>(take(static_cast<size_t>(nq) * 8));
    w.cand = reinterpret_cast<uint32_t*>(take(static_cast<size_t>(nq) * kSelectCap * 4));
    w.dkey = reinterpret_cast<uint64_t*>(take(static_cast<size_t>(nq) * kSelectCap * 8));
    w.total = off;
    return w;
}

__device__ __forceinline__ bool row_passes(const FilterView& fv, uint32_t q, uint32_t row) {
    if (!fv.q_filter) return true;
    const uint32_t fid = fv.q_filter[q];
    if (fid == kNoFilter) return true;
    return (fv.masks[static_cast<uint64_t>(fid) * fv.words + (row >> 5)] >> (row & 31)) & 1u;
}

// pass 0: bins = key bits 31..20; pass 1: rows whose key bits 31..20 equal sel_bin[q], bins = key bits 19..8
__global__ void __launch_bounds__(256)
select_hist_kernel(const float* __restrict__ scores, uint64_t ld, uint32_t n_valid, FilterView fv, int pass,
                   const uint32_t* __restrict__ sel_bin, uint32_t* hist) {
    __shared__ uint32_t h[kSelBins];
    const uint32_t q = blockIdx.y;
    for (uint32_t b = threadIdx.x; b < kSelBins; b += blockDim.x) h[b] = 0;
    __syncthreads();
    const uint32_t want = pass ? sel_bin[q] : 0u;
    if (!(pass && want == 0xFFFFFFFFu)) {
        const float* row = scores + static_cast<uint64_t>(q) * ld;
        const uint32_t lane = threadIdx.x & 31;
        // warp-uniform trip count; scores of one query crowd into a few dozen bins (cosines of random rows share their
        // exponent), so lanes that hit the same bin are merged (match.any) into ONE shared-memory atomic
        for (uint32_t base = blockIdx.x * blockDim.x + (threadIdx.x & ~31u); base < n_valid; base += gridDim.x * blockDim.x) {
            const uint32_t i = base + lane;
            bool take = false;
            uint32_t bin = 0;
            if (i < n_valid) {
                const float v = __ldg(row + i);
                if (v == v && row_passes(fv, q, i)) {
                    const uint32_t key = float_key(v);
                    if (!pass) { take = true; bin = key >> 20; }
                    else if ((key >> 20) == want) { take = true; bin = (key >> 8) & 0xFFFu; }
                }
            }
            const uint32_t mask = __ballot_sync(0xffffffffu, take);
            if (take) {
                const uint32_t peers = __match_any_sync(mask, bin);
                if (lane == static_cast<uint32_t>(__ffs(peers)) - 1u) atomicAdd(&h[bin], static_cast<uint32_t>(__popc(peers)));
            }
        }
    }
    __syncthreads();
    for (uint32_t b = threadIdx.x; b < kSelBins; b += blockDim.x)
        if (h[b]) atomicAdd(&hist[static_cast<uint64_t>(q) * kSelBins + b], h[b]);
}

// one block of 256 threads per query: the highest bin b with  count(bins >= b) >= m
__global__ void __launch_bounds__(256)
select_pick_kernel(const uint32_t* __restrict__ hist, int pass, uint32_t k_out, const uint32_t* __restrict__ q_k,
                   uint32_t* sel_bin, uint32_t* sel_rem, float* thr, uint32_t* count, NoiseBound nb,
                   const double* __restrict__ qnorm) {
    __shared__ uint32_t seg[256];
    __shared__ uint32_t res[2];
    const uint32_t q = blockIdx.x, t = threadIdx.x;
    const uint32_t* h = hist + static_cast<uint64_t>(q) * kSelBins;
    uint32_t m = pass ? sel_rem[q] : (q_k ? min(q_k[q], k_out) : k_out);
    if (pass && sel_bin[q] == 0xFFFFFFFFu) {  // pass 0 found fewer than m eligible rows: everything is a candidate
        const uint32_t m0 = q_k ? min(q_k[q], k_out) : k_out;
        if (t == 0) { thr[q] = m0 ? -INFINITY : INFINITY; count[q] = 0; }
        return;
    }
    uint32_t local = 0;
    for (uint32_t b = 0; b < 16; ++b) local += h[t * 16 + b];
    seg[t] = local;
    if (t == 0) { res[0] = 0xFFFFFFFFu; res[1] = 0; }
    __syncthreads();
    if (t == 0) {
        uint32_t above = 0;
        for (int sgi = 255; sgi >= 0; --sgi) {
            if (above + seg[sgi] >= m && m > 0) {
                for (int b = 15; b >= 0; --b) {
                    const uint32_t c = h[sgi * 16 + b];
                    if (above + c >= m) { res[0] = static_cast<uint32_t>(sgi * 16 + b); res[1] = m - above; break; }
                    above += c;
                }
                break;
            }
            above += seg[sgi];
        }
        if (!pass) {
            sel_bin[q] = res[0];
            sel_rem[q] = res[1];
        } else {
            // lower edge of the 24-bit prefix bin <= the m-th largest score; minus 2 eps (see the header comment)
            float T = -INFINITY;
            if (res[0] != 0xFFFFFFFFu) {
                const uint32_t prefix = (sel_bin[q] << 20) | (res[0] << 8);
                const float edge = key_to_float(prefix);
                T = static_cast<float>(static_cast<double>(edge) - 2.0 * noise_eps(nb, qnorm[q]));
                T = nextafterf(T, -INFINITY);  // the f64 -> f32 rounding must not raise the threshold
            }
            thr[q] = T;
            count[q] = 0;
        }
    }
}

__global__ void __launch_bounds__(256)
select_gather_kernel(const float* __restrict__ scores, uint64_t ld, uint32_t n_valid, FilterView fv,
                     const float* __restrict__ thr, uint32_t* count, uint32_t* cand) {
    const uint32_t q = blockIdx.y;
    const float T = thr[q];
    const float* row = scores + static_cast<uint64_t>(q) * ld;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_valid; i += gridDim.x * blockDim.x) {
        const float v = __ldg(row + i);
        if (!(v >= T) || !row_passes(fv, q, i)) continue;
        const uint32_t pos = atomicAdd(&count[q], 1u);
        if (pos < kSelectCap) cand[static_cast<uint64_t>(q) * kSelectCap + pos] = i;
    }
}

// 8 lanes per candidate: f64 distance with the reference formula
template <typename RT, typename QT>
__global__ void __launch_bounds__(256)
select_rerank_kernel(const ExactView ev, int metric, const uint32_t* __restrict__ count,
                     const uint32_t* __restrict__ cand, const double* __restrict__ qnorm, uint64_t* dkey) {
    const uint32_t q = blockIdx.y;
    const uint32_t n = min(count[q], kSelectCap);
    const uint32_t lane = threadIdx.x & 31, sub = lane >> 3, sl = lane & 7;
    const uint32_t groups_per_block = blockDim.x / 8;
    const QT* qv = static_cast<const QT*>(ev.queries) + static_cast<uint64_t>(ev.q_index ? ev.q_index[q] : q) * ev.q_stride;
    const RT* rows = static_cast<const RT*>(ev.rows);
    const bool vec = dot_vec_ok<RT, QT>(rows, ev.row_stride, qv);
    const double nq = qnorm[q];
    for (uint32_t c0 = blockIdx.x * groups_per_block; c0 < n; c0 += gridDim.x * groups_per_block) {  // warp-uniform trip count
        const uint32_t c = c0 + (threadIdx.x >> 5) * 4 + sub;
        const bool active = c < n;
        const uint32_t slot = active ? cand[static_cast<uint64_t>(q) * kSelectCap + c] : 0u;
        double dot, nx;
        group_dot_f64<RT, QT, 8>(rows + static_cast<uint64_t>(slot) * ev.row_stride, qv, ev.len, sl, active, dot, nx, vec);
        if (active && sl == 0) dkey[static_cast<uint64_t>(q) * kSelectCap + c] = sortable_key(distance_f64(metric, dot, nx, nq));
    }
}

// one block per query: bitonic sort of the (distance key, slot) pairs in shared memory, first k_out emitted
__global__ void __launch_bounds__(1024)
select_sort_emit_kernel(const uint32_t* __restrict__ count, const uint32_t* __restrict__ cand,
                        const uint64_t* __restrict__ dkey, uint32_t k_out, uint32_t* out_slots, double* out_dist,
                        uint32_t* overflow) {
    extern __shared__ __align__(16) uint8_t sm[];
    const uint32_t q = blockIdx.x, t = threadIdx.x;
    const uint32_t total = count[q];
    const uint32_t n = min(total, kSelectCap);
    uint32_t P = 1;
    while (P < n) P <<= 1;
    uint64_t* kk = reinterpret_cast<uint64_t*>(sm);
    uint32_t* vv = reinterpret_cast<uint32_t*>(sm + static_cast<size_t>(kSelectCap) * 8);
    for (uint32_t i = t; i < P; i += blockDim.x) {
        kk[i] = i < n ? dkey[static_cast<uint64_t>(q) * kSelectCap + i] : ~0ull;
        vv[i] = i < n ? cand[static_cast<uint64_t>(q) * kSelectCap + i] : 0xFFFFFFFFu;
    }
    __syncthreads();
    for (uint32_t size = 2; size <= P; size <<= 1) {
        for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
            for (uint32_t i = t; i < P / 2; i += blockDim.x) {
                const uint32_t lo = 2 * i - (i & (stride - 1)), hi = lo + stride;
                const bool up = (lo & size) == 0;
                const uint64_t ka = kk[lo], kb = kk[hi];
                const uint32_t va = vv[lo], vb = vv[hi];
                const bool a_gt_b = ka > kb || (ka == kb && va > vb);  // (distance, slot) order: rs:220-224
                if (a_gt_b == up) { kk[lo] = kb; kk[hi] = ka; vv[lo] = vb; vv[hi] = va; }
            }
            __syncthreads();
        }
    }
    for (uint32_t r = t; r < k_out; r += blockDim.x) {
        const uint64_t o = static_cast<uint64_t>(q) * k_out + r;
        if (r < n) { out_slots[o] = vv[r]; out_dist[o] = key_to_double(kk[r]); }
        else { out_slots[o] = kInvalidSlot; out_dist[o] = INFINITY; }
    }
    if (t == 0) overflow[q] = total > kSelectCap ? 1u : 0u;
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
                                   float* xmax_sq, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    const uint64_t blocks = (n + kWarpsPerBlock - 1) / kWarpsPerBlock;
    upsert_kernel<T><<<static_cast<uint32_t>(blocks), kWarpsPerBlock * 32, 0, s>>>(src, stride, slots, n, dims, dpad,
                                                                                   metric, rows, aux, sh, xmax_sq);
    return cudaGetLastError();
}
cudaError_t launch_upsert_f64(const double* src, const uint32_t* slots, uint64_t n, uint32_t dims, uint32_t dpad,
                              int metric, __nv_bfloat16* rows, RowAux* aux, ShadowStore sh, float* xmax_sq,
                              cudaStream_t s) {
    return launch_upsert_t<double>(src, dims, slots, n, dims, dpad, metric, rows, aux, sh, xmax_sq, s);
}
cudaError_t launch_upsert_f32(const float* src, const uint32_t* slots, uint64_t n, uint32_t dims, uint32_t dpad,
                              int metric, __nv_bfloat16* rows, RowAux* aux, ShadowStore sh, float* xmax_sq,
                              cudaStream_t s) {
    return launch_upsert_t<float>(src, dims, slots, n, dims, dpad, metric, rows, aux, sh, xmax_sq, s);
}
cudaError_t launch_upsert_bf16(const __nv_bfloat16* src, uint64_t src_stride, const uint32_t* slots, uint64_t n,
                               uint32_t dims, uint32_t dpad, int metric, __nv_bfloat16* rows, RowAux* aux,
                               ShadowStore sh, float* xmax_sq, cudaStream_t s) {
    return launch_upsert_t<__nv_bfloat16>(src, src_stride, slots, n, dims, dpad, metric, rows, aux, sh, xmax_sq, s);
}

cudaError_t launch_move_rows(const uint32_t* pairs, uint64_t n, uint32_t dpad, __nv_bfloat16* rows, RowAux* aux,
                             ShadowStore sh, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    const uint64_t blocks = (n + kWarpsPerBlock - 1) / kWarpsPerBlock;
    move_rows_kernel<<<static_cast<uint32_t>(blocks), kWarpsPerBlock * 32, 0, s>>>(pairs, n, dpad, rows, aux, sh);
    return cudaGetLastError();
}

cudaError_t launch_pack_queries_f64(const double* src, const uint32_t* q_index, uint64_t nq, uint32_t dims,
                                    uint32_t dpad, uint64_t qpad, __nv_bfloat16* dst, float* gthr, cudaStream_t s) {
    pack_queries_kernel<double><<<blocks_for(qpad * dpad, 256), 256, 0, s>>>(src, dims, q_index, nq, dims, dpad, qpad,
                                                                             dst, gthr);
    return cudaGetLastError();
}
cudaError_t launch_pack_queries_f32(const float* src, const uint32_t* q_index, uint64_t nq, uint32_t dims,
                                    uint32_t dpad, uint64_t qpad, __nv_bfloat16* dst, float* gthr, cudaStream_t s) {
    pack_queries_kernel<float><<<blocks_for(qpad * dpad, 256), 256, 0, s>>>(src, dims, q_index, nq, dims, dpad, qpad,
                                                                            dst, gthr);
    return cudaGetLastError();
}
cudaError_t launch_pack_queries_bf16(const __nv_bfloat16* src, uint64_t src_stride, const uint32_t* q_index,
                                     uint64_t nq, uint32_t dims, uint32_t dpad, uint64_t qpad, __nv_bfloat16* dst,
                                     float* gthr, cudaStream_t s) {
    pack_queries_kernel<__nv_bfloat16><<<blocks_for(qpad * dpad, 256), 256, 0, s>>>(src, src_stride, q_index, nq, dims,
                                                                                    dpad, qpad, dst, gthr);
    return cudaGetLastError();
}

cudaError_t launch_merge_refine(const MergeArgs& a, cudaStream_t s) {
    if (a.n_q == 0) return cudaSuccess;
    if (a.n_slabs > 160u || a.n_extract > 128u) return cudaErrorInvalidValue;  // host planner limits (kMergeMaxCand, kFastMaxKPrime)
    const size_t per_warp = align_up(static_cast<size_t>(a.n_extract) * 8 +
                                         static_cast<size_t>(a.n_slabs) * (a.list_len | 1u) * 8 +
                                         static_cast<size_t>(a.n_extract) * 8, 16);
    uint32_t warps = static_cast<uint32_t>(std::min<size_t>(4, std::max<size_t>(1, (48 * 1024) / per_warp)));
    // one warp per query is latency / conversion-unit bound (f64 re-rank: two F2F per element at 16 lanes/clk/SM):
    // spread a small batch over all 148 SMs instead of packing 4 queries per block
    warps = std::max<uint32_t>(1, std::min<uint32_t>(warps, a.n_q / 296));
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

size_t select_workspace_bytes(uint32_t nq) { return select_layout(nullptr, nq).total; }

template <typename RT, typename QT>
static cudaError_t select_rerank_launch(const SelectArgs& a, const SelectWs& w, cudaStream_t s) {
    query_norms_kernel<QT><<<(a.nq + kWarpsPerBlock - 1) / kWarpsPerBlock, kWarpsPerBlock * 32, 0, s>>>(a.ev, 0, a.nq, w.qnorm);
    return cudaGetLastError();
}
template <typename RT, typename QT>
static cudaError_t select_rerank_launch2(const SelectArgs& a, const SelectWs& w, cudaStream_t s) {
    const dim3 grid(std::max<uint32_t>(1, std::min<uint32_t>(kSelectCap / 32, 296 / std::max<uint32_t>(1, a.nq) + 1)), a.nq);
    select_rerank_kernel<RT, QT><<<grid, 256, 0, s>>>(a.ev, a.metric, w.count, w.cand, w.qnorm, w.dkey);
    return cudaGetLastError();
}

cudaError_t launch_select_search(const SelectArgs& a, cudaStream_t s) {
    if (a.nq == 0 || a.k_out == 0) return cudaSuccess;
    if (a.k_out > kSelectMaxK) return cudaErrorInvalidValue;
    const SelectWs w = select_layout(static_cast<uint8_t*>(a.workspace), a.nq);
    cudaError_t e;
#define PWKNN_SEL_DISPATCH(FN)                                                      \
    switch (a.ev.row_kind * 3 + a.ev.q_kind) {                                      \
        case 0: e = FN<__nv_bfloat16, __nv_bfloat16>(a, w, s); break;               \
        case 3: e = FN<float, __nv_bfloat16>(a, w, s); break;                       \
        case 4: e = FN<float, float>(a, w, s); break;                               \
        case 5: e = FN<float, double>(a, w, s); break;                              \
        case 6: e = FN<double, __nv_bfloat16>(a, w, s); break;                      \
        case 7: e = FN<double, float>(a, w, s); break;                              \
        case 8: e = FN<double, double>(a, w, s); break;                             \
        default: return cudaErrorInvalidValue;                                      \
    }
    PWKNN_SEL_DISPATCH(select_rerank_launch)  // |q|^2 first: the gather threshold needs eps(q)
    if (e != cudaSuccess) return e;
    // enough blocks per query to fill the GPU, few enough that the per-block histogram merge stays cheap
    const uint32_t bx = std::max<uint32_t>(1, std::min<uint32_t>((a.n_valid + 256 * 64 - 1) / (256 * 64), 592 / a.nq + 1));
    const dim3 grid(bx, a.nq);
    for (int pass = 0; pass < 2; ++pass) {
        e = cudaMemsetAsync(w.hist, 0, static_cast<size_t>(a.nq) * kSelBins * 4, s);
        if (e != cudaSuccess) return e;
        select_hist_kernel<<<grid, 256, 0, s>>>(a.scores, a.ld, a.n_valid, a.fv, pass, w.sel_bin, w.hist);
        select_pick_kernel<<<a.nq, 256, 0, s>>>(w.hist, pass, a.k_out, a.q_k, w.sel_bin, w.sel_rem, w.thr, w.count,
                                                a.noise, w.qnorm);
    }
    select_gather_kernel<<<grid, 256, 0, s>>>(a.scores, a.ld, a.n_valid, a.fv, w.thr, w.count, w.cand);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    PWKNN_SEL_DISPATCH(select_rerank_launch2)
    if (e != cudaSuccess) return e;
#undef PWKNN_SEL_DISPATCH
    const size_t smem = static_cast<size_t>(kSelectCap) * 12;
    e = cudaFuncSetAttribute(select_sort_emit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    select_sort_emit_kernel<<<a.nq, 1024, smem, s>>>(w.count, w.cand, w.dkey, a.k_out, a.out_slots, a.out_dist,
                                                     a.overflow_out);
    return cudaGetLastError();
}

cudaError_t launch_read_rows(const __nv_bfloat16* rows, uint64_t slot0, uint64_t n, uint32_t dims, uint32_t dpad,
                             double* out, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    read_rows_kernel<<<blocks_for(n * dims, 256), 256, 0, s>>>(rows, slot0, n, dims, dpad, out);
    return cudaGetLastError();
}

}  // namespace pwknn
