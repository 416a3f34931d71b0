// Fused  scores = Q . Corpus^T  (tcgen05, bf16 -> f32 in TMEM)  +  per-query top-k'  for sm_100a.
//
// Replaces the hot loop of BruteForceKNNIndex::search
//   (/root/reference/src/external_integration/brute_force_knn_integration.rs:205-232:
//    index_arr.dot(query) -> fill_{cos,l2sq}_distances -> k_smallest per query)
// with ONE persistent kernel that never materialises the [N x Q] distance matrix:
//
//   work item = (corpus slab s, query tile m);  128 queries x 256 corpus rows per MMA tile
//   warp 0    : TMA producer   -- query chunk [128 x 64] + corpus chunk [256 x 64] bf16 per stage,
//                                 128-B swizzle, mbarrier expect_tx
//   warp 1    : MMA issuer     -- 4 x tcgen05.mma (128x256x16) per stage into a double-buffered
//                                 TMEM accumulator (2 x 256 columns); tcgen05.commit frees smem / signals epilogue
//   warps 2-5 : epilogue       -- tcgen05.ld 32 columns at a time; thread <-> query (TMEM lane);
//                                 v = fma(dot, scale[row], bias[row]); threshold test against the k'-th best;
//                                 rare path: sorted insert into the thread's smem-resident candidate list
//   at the end of a slab every thread writes its k' candidates (score, slot); a second small kernel
//   (merge_refine) merges slabs and re-ranks in f64.
//
// Selection order is (score desc, slot asc): columns are visited in ascending slot order and a later
// equal score never displaces an earlier one, which is the (OrderedFloat(distance), idx) order of the
// reference (:220-224) restricted to fp32 scores; exactness is restored by the f64 re-rank.
#include <cfloat>
#include <cmath>

#include "ptx_sm100.cuh"
#include "pwknn_internal.h"

namespace pwknn {

namespace {

constexpr uint32_t kNumThreads = 192;
constexpr uint32_t kStageBytesA = kBlockM * kBlockK * 2;  // 16 KB query chunk per CTA
// corpus chunk per CTA: the whole [256 x 64] tile (1 CTA) or this CTA's half of it (CTA pair)
__host__ __device__ constexpr uint32_t stage_bytes_b(int ctas) { return kBlockN / ctas * kBlockK * 2; }
__host__ __device__ constexpr uint32_t stage_bytes(int ctas) { return kStageBytesA + stage_bytes_b(ctas); }
__host__ __device__ constexpr uint32_t max_stages(int ctas) { return ctas == 2 ? 6u : 4u; }
constexpr uint32_t kBarSlots = 6;    // full[6], empty[6], tfull[2], tempty[2]
constexpr uint32_t kTmemCols = 512;  // 2 accumulator buffers x 256 fp32 columns
constexpr uint32_t kAuxBytes = 2 * kBlockN * sizeof(RowAux);
constexpr uint32_t kStashBytes = 32 * kBlockM * 4;  // one 32-column chunk of scores per epilogue thread
constexpr uint32_t kBarBytes = 256;

// Lower bound shared by all slabs of a query: max over slabs of the local k'-th best score.  Any
// slab's k'-th best is <= the global k'-th best, so scores below it can never reach the final top-k'.
__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
    if (v >= 0.f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
    else atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

// Register-resident sorted candidate list (descending score; equal scores keep arrival = slot order).
template <int K>
struct RegList {
    float v[K];
    uint32_t i[K];
    __device__ __forceinline__ void init() {
#pragma unroll
        for (int j = 0; j < K; ++j) { v[j] = -INFINITY; i[j] = kInvalidSlot; }
    }
    __device__ __forceinline__ float kth() const { return v[K - 1]; }
    // precondition: x > v[K-1]
    __device__ __forceinline__ void insert(float x, uint32_t slot) {
        v[K - 1] = x;
        i[K - 1] = slot;
#pragma unroll
        for (int j = K - 1; j > 0; --j) {
            const bool sw = v[j] > v[j - 1];  // strict: an equal, earlier entry stays in front
            const float a = v[j - 1], b = v[j];
            const uint32_t ia = i[j - 1], ib = i[j];
            v[j - 1] = sw ? b : a; v[j] = sw ? a : b;
            i[j - 1] = sw ? ib : ia; i[j] = sw ? ia : ib;
        }
    }
};

// Shared-memory candidate lists for 32 < k' <= 108: one sorted (descending) list per query, stored
// row-major [128][stride] with an odd stride, and maintained COOPERATIVELY by the 32 lanes of the warp that
// owns the query (a per-thread insertion sort would serialise 32 divergent O(k') shifts).
// Inserts (v, slot) into the list of query row `qrow`; precondition v > current k'-th best.  Warp-uniform.
// Returns the new k'-th best score.
__device__ __forceinline__ float coop_list_insert(float* s_val, uint32_t* s_idx, uint32_t stride, uint32_t kprime,
                                                  uint32_t qrow, uint32_t lane, float v, uint32_t slot) {
    float* lv = s_val + qrow * stride;
    uint32_t* li = s_idx + qrow * stride;
    float prev_v[4];
    uint32_t prev_i[4];
    uint32_t pos = 0;  // entries that stay in front: old >= v (equal scores: the earlier, lower slot stays first)
#pragma unroll
    for (uint32_t c = 0; c < 4; ++c) {
        const uint32_t p = c * 32 + lane;
        bool ge = false;
        prev_v[c] = 0.f;
        prev_i[c] = 0;
        if (p < kprime) {
            ge = !(lv[p] < v);
            if (p > 0) {
                prev_v[c] = lv[p - 1];
                prev_i[c] = li[p - 1];
            }
        }
        pos += __popc(__ballot_sync(0xffffffffu, ge));
    }
    __syncwarp();
#pragma unroll
    for (uint32_t c = 0; c < 4; ++c) {
        const uint32_t p = c * 32 + lane;
        if (p < kprime) {
            if (p > pos) {
                lv[p] = prev_v[c];
                li[p] = prev_i[c];
            } else if (p == pos) {
                lv[p] = v;
                li[p] = slot;
            }
        }
    }
    __syncwarp();
    return lv[kprime - 1];
}

// KREG > 0: candidate lists live in registers (k' == KREG);  KREG == 0: in shared memory (k' <= 108).
// kCtas == 2: a CTA pair (cluster of 2, cta_group::2) works on one corpus tile and TWO query tiles: each CTA
//   TMA-loads its own query chunk and HALF of the corpus chunk, the leader issues M=256 MMAs that read both
//   CTAs' smem and write both CTAs' TMEM, commits are multicast to both CTAs -- half the L2->SMEM corpus
//   traffic per flop and a 6-deep ring instead of 4.
// kMode: 0 = plain, 1 = also dump every score (debug), 2 = filter push-down (per-query row bitmask)
template <int kMode, int KREG, int kCtas>
__global__ void __launch_bounds__(kNumThreads, 1)
fused_knn_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_c,
                 const FusedParams p) {
    constexpr bool kDump = kMode == 1;
    constexpr bool kFilter = kMode == 2;
    constexpr uint32_t kStageBytesB = stage_bytes_b(kCtas);
    constexpr uint32_t kStageBytes = stage_bytes(kCtas);
    const uint32_t cta_rank = kCtas == 2 ? cluster_ctarank() : 0u;
    const uint32_t group_id = blockIdx.x / kCtas;        // CTA (pair) index
    const uint32_t n_groups_grid = gridDim.x / kCtas;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t warp = threadIdx.x >> 5;
    const uint32_t lane = threadIdx.x & 31;

    const uint32_t raw_addr = smem_u32(smem_raw);
    const uint32_t base = (raw_addr + 1023u) & ~1023u;  // SWIZZLE_128B tiles need 1024-B alignment
    uint8_t* gbase = smem_raw + (base - raw_addr);
    const uint32_t stages = p.stages;
    const uint32_t kprime = p.kprime;
    const uint32_t off_lists = stages * kStageBytes;
    const uint32_t list_stride = kprime | 1u;  // odd: conflict-free for both row- and column-wise access
    const uint32_t list_bytes = KREG > 0 ? 0u : list_stride * kBlockM * 8;
    float* s_val = reinterpret_cast<float*>(gbase + off_lists);
    uint32_t* s_idx = reinterpret_cast<uint32_t*>(gbase + off_lists + list_bytes / 2);
    uint32_t* s_stash = reinterpret_cast<uint32_t*>(gbase + off_lists + list_bytes);
    RowAux* s_aux = reinterpret_cast<RowAux*>(gbase + off_lists + list_bytes + kStashBytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(gbase + off_lists + list_bytes + kStashBytes + kAuxBytes);
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 16);

    const uint32_t bar_full = smem_u32(bars);                   // [kBarSlots]  (pair: only the leader's are used)
    const uint32_t bar_empty = bar_full + 8 * kBarSlots;        // [kBarSlots]
    const uint32_t bar_tfull = bar_empty + 8 * kBarSlots;       // [2]
    const uint32_t bar_tempty = bar_tfull + 16;                 // [2]        (pair: only the leader's are used)

    if (p.cta_times && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        p.cta_times[2 * blockIdx.x] = t;
    }
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmap_q);
        tma_prefetch_desc(&tmap_c);
        for (uint32_t s = 0; s < kBarSlots; ++s) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_empty + 8 * s, 1);
        }
        for (uint32_t b = 0; b < 2; ++b) {
            mbar_init(bar_tfull + 8 * b, 1);
            mbar_init(bar_tempty + 8 * b, 4 * kCtas);  // one arrive per epilogue warp (of both CTAs)
        }
        fence_mbar_init();
    }
    if (warp == 1) {
        if constexpr (kCtas == 2) tmem_alloc_2cta<kTmemCols>(smem_u32(s_tmem));
        else tmem_alloc<kTmemCols>(smem_u32(s_tmem));
    }
    tc_fence_before();
    __syncthreads();
    if constexpr (kCtas == 2) cluster_sync_all();  // peer barriers are initialised before any remote arrive / TMA
    tc_fence_after();
    const uint32_t tmem_base = *s_tmem;

    // work item = (slab, query-tile group); a group is kCtas consecutive query tiles, one per CTA of the pair
    const uint32_t n_items = p.n_slabs * p.n_qgroups;

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (uint32_t item = group_id; item < n_items; item += n_groups_grid) {
                if (p.group_flags && __ldcg(p.group_flags + item % p.n_qgroups) == 0) continue;  // all roles agree
                const uint32_t mt = (item % p.n_qgroups) * kCtas + cta_rank;
                const uint32_t slab = item / p.n_qgroups;
                const uint32_t t0 = static_cast<uint32_t>(static_cast<uint64_t>(slab) * p.n_tiles / p.n_slabs);
                const uint32_t t1 = static_cast<uint32_t>(static_cast<uint64_t>(slab + 1) * p.n_tiles / p.n_slabs);
                const bool gate = p.slab_progress != nullptr && p.lead_tiles > 0 && cta_rank == 0 && p.n_qgroups > 1;
                volatile uint32_t* prog = gate ? p.slab_progress + static_cast<uint64_t>(slab) * p.n_qgroups : nullptr;
                const uint32_t my_g = item % p.n_qgroups;
                uint32_t sib_min = 0x7FFFFFFFu;  // cached: tiles issued by the slowest started sibling
                for (uint32_t t = t0; t < t1; ++t) {
                    if (gate) {
                        const uint32_t mine = t - t0 + 1;  // tiles issued once this one is
                        while (mine > sib_min + p.lead_tiles || sib_min == 0x7FFFFFFFu) {
                            uint32_t m = 0x7FFFFFFFu;
                            for (uint32_t g = 0; g < p.n_qgroups; ++g) {
                                if (g == my_g) continue;
                                const uint32_t v = prog[g];
                                if (v != 0 && v < m) m = v;  // started and not finished
                            }
                            sib_min = m;
                            if (m == 0x7FFFFFFFu || mine <= m + p.lead_tiles) break;  // nobody to wait for
                            __nanosleep(256);
                        }
                    }
                    for (uint32_t kc = 0; kc < p.k_chunks; ++kc) {
                        mbar_wait(bar_empty + 8 * stage, phase ^ 1);  // this CTA's smem slot is free
                        const uint32_t sa = base + stage * kStageBytes;
                        if constexpr (kCtas == 2) {
                            // both CTAs' bytes are counted on the leader's barrier; only the leader arms it
                            const uint32_t lbar = mapa_cluster(bar_full + 8 * stage, 0);
                            if (cta_rank == 0) mbar_arrive_expect_tx(bar_full + 8 * stage, 2 * kStageBytes);
                            tma_load_2d_2cta(sa, &tmap_q, lbar, static_cast<int32_t>(kc * kBlockK),
                                             static_cast<int32_t>(mt * kBlockM), kL2EvictLast);
                            tma_load_2d_2cta(sa + kStageBytesA, &tmap_c, lbar, static_cast<int32_t>(kc * kBlockK),
                                             static_cast<int32_t>(t * kBlockN + cta_rank * (kBlockN / 2)), kL2EvictNormal);
                        } else {
                            mbar_arrive_expect_tx(bar_full + 8 * stage, kStageBytes);
                            tma_load_2d_hint(sa, &tmap_q, bar_full + 8 * stage, static_cast<int32_t>(kc * kBlockK),
                                             static_cast<int32_t>(mt * kBlockM), kL2EvictLast);
                            tma_load_2d(sa + kStageBytesA, &tmap_c, bar_full + 8 * stage,
                                        static_cast<int32_t>(kc * kBlockK), static_cast<int32_t>(t * kBlockN));
                        }
                        if (++stage == stages) {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
                    if (gate) prog[my_g] = (t + 1 == t1) ? 0x7FFFFFFFu : (t - t0 + 1);
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer (one thread, leader CTA only)
        if (lane == 0 && cta_rank == 0) {
            constexpr uint32_t idesc = make_idesc_bf16_f32(kBlockM * kCtas, kBlockN);
            uint32_t stage = 0, phase = 0, it = 0;
            for (uint32_t item = group_id; item < n_items; item += n_groups_grid) {
                if (p.group_flags && __ldcg(p.group_flags + item % p.n_qgroups) == 0) continue;
                const uint32_t slab = item / p.n_qgroups;
                const uint32_t t0 = static_cast<uint32_t>(static_cast<uint64_t>(slab) * p.n_tiles / p.n_slabs);
                const uint32_t t1 = static_cast<uint32_t>(static_cast<uint64_t>(slab + 1) * p.n_tiles / p.n_slabs);
                for (uint32_t t = t0; t < t1; ++t, ++it) {
                    const uint32_t buf = it & 1, use = it >> 1;
                    mbar_wait(bar_tempty + 8 * buf, (use & 1) ^ 1);  // epilogue drained this accumulator
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + buf * kBlockN;
                    for (uint32_t kc = 0; kc < p.k_chunks; ++kc) {
                        mbar_wait(bar_full + 8 * stage, phase);  // TMA bytes landed
                        tc_fence_after();
                        const uint32_t sa = base + stage * kStageBytes;
                        const uint64_t a_desc = make_smem_desc_sw128(sa);
                        const uint64_t b_desc = make_smem_desc_sw128(sa + kStageBytesA);
#pragma unroll
                        for (uint32_t k = 0; k < kBlockK / kUmmaK; ++k) {
                            // +32 B per K=16 slice inside the 128-B swizzle atom: start-address field += 2
                            if constexpr (kCtas == 2)
                                umma_bf16_ss_2cta(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kc | k) != 0 ? 1u : 0u);
                            else
                                umma_bf16_ss(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kc | k) != 0 ? 1u : 0u);
                        }
                        // smem slot free (in both CTAs) once these MMAs retire
                        if constexpr (kCtas == 2) umma_commit_2cta_mc(bar_empty + 8 * stage, 3);
                        else umma_commit(bar_empty + 8 * stage);
                        if (++stage == stages) {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
                    // accumulator complete -> epilogue (of both CTAs)
                    if constexpr (kCtas == 2) umma_commit_2cta_mc(bar_tfull + 8 * buf, 3);
                    else umma_commit(bar_tfull + 8 * buf);
                }
            }
        }
    } else {
        // ------------------------------------------------------------ epilogue: thread <-> query
        const uint32_t et = threadIdx.x - 64;       // 0..127
        const uint32_t lane_base = (warp & 3) * 32;  // TMEM lanes this warp may read
        const uint32_t row = lane_base + lane;       // query row inside the tile
        const uint64_t qpad = static_cast<uint64_t>(p.n_qgroups) * kCtas * kBlockM;
        uint32_t it = 0;
        const bool share = p.gthr != nullptr;
        for (uint32_t item = group_id; item < n_items; item += n_groups_grid) {
            if (p.group_flags && __ldcg(p.group_flags + item % p.n_qgroups) == 0) continue;
            const uint32_t mt = (item % p.n_qgroups) * kCtas + cta_rank;
            const uint32_t slab = item / p.n_qgroups;
            const uint32_t t0 = static_cast<uint32_t>(static_cast<uint64_t>(slab) * p.n_tiles / p.n_slabs);
            const uint32_t t1 = static_cast<uint32_t>(static_cast<uint64_t>(slab + 1) * p.n_tiles / p.n_slabs);
            const uint32_t q = mt * kBlockM + row;
            RegList<(KREG > 0 ? KREG : 1)> rl;
            if constexpr (KREG > 0) {
                rl.init();
            } else {
                for (uint32_t r = 0; r < 32; ++r) {  // this warp's 32 lists, lanes across positions
                    for (uint32_t j = lane; j < kprime; j += 32) {
                        s_val[(lane_base + r) * list_stride + j] = -INFINITY;
                        s_idx[(lane_base + r) * list_stride + j] = kInvalidSlot;
                    }
                }
                __syncwarp();
            }
            // filter push-down: this query only sees rows whose bit is set in its filter's mask
            const uint32_t* fmask = nullptr;
            if constexpr (kFilter) {
                const uint32_t fid = p.q_filter[q];
                if (fid != kNoFilter) fmask = p.filter_masks + static_cast<uint64_t>(fid) * p.filter_words;
            }
            float thr = -INFINITY;  // local k'-th best (strict test)
            for (uint32_t t = t0; t < t1; ++t, ++it) {
                const uint32_t buf = it & 1, use = it >> 1;
                const uint32_t col0 = t * kBlockN;
                RowAux* aux_t = s_aux + buf * kBlockN;
                aux_t[et] = p.aux[col0 + et];
                aux_t[et + kBlockM] = p.aux[col0 + et + kBlockM];
                const float gthr = share ? __ldcg(p.gthr + q) : -INFINITY;  // cross-slab lower bound (non-strict test)
                named_bar_sync(1, 128);
                mbar_wait(bar_tfull + 8 * buf, use & 1);
                tc_fence_after();
                const uint32_t taddr = tmem_base + (lane_base << 16) + buf * kBlockN;
#pragma unroll 1
                for (uint32_t c = 0; c < ((p.debug_flags & 1u) ? 0u : kBlockN); c += 32) {
                    uint32_t r[32];
                    tmem_ld_32x32b_x32(taddr + c, r);
                    tmem_ld_wait();
                    // Per score: one FFMA + one FMNMX.  The exact tests (v > thr_local && v >= thr_global) run only
                    // for chunks whose maximum passes them -- rare after warm-up -- which keeps the epilogue's
                    // instruction (and energy) cost per tile low: the kernel is power-capped.
                    float vmax = -INFINITY;
                    uint32_t fw = 0xFFFFFFFFu;  // 32 rows of this chunk <-> one mask word (col0 + c is a multiple of 32)
                    if constexpr (kFilter) {
                        if (fmask) fw = __ldg(fmask + ((col0 + c) >> 5));
                    }
#pragma unroll
                    for (uint32_t j = 0; j < 32; ++j) {
                        const RowAux a = aux_t[c + j];
                        float v = fmaf(__uint_as_float(r[j]), a.scale, a.bias);
                        if constexpr (kFilter) v = ((fw >> j) & 1u) ? v : -INFINITY;
                        r[j] = __float_as_uint(v);
                        vmax = fmaxf(vmax, v);  // NaN scores drop out (they are never candidates)
                    }
                    if constexpr (kDump) {
                        // Score dump (select path for large k, debug): transpose the warp's [32 queries x 32 rows]
                        // block through the stash so that every store instruction writes 32 consecutive floats of ONE
                        // query (a thread holds 32 rows of its own query: direct stores would be 32 scattered sectors).
                        // XOR-swizzled columns keep both the parking stores and the transposed loads conflict-free.
#pragma unroll
                        for (uint32_t j = 0; j < 32; ++j) s_stash[j * kBlockM + lane_base + (lane ^ j)] = r[j];
                        __syncwarp();
                        const uint32_t col = col0 + c + lane;
                        const uint32_t q0w = mt * kBlockM + lane_base;
#pragma unroll 4
                        for (uint32_t rr = 0; rr < 32; ++rr) {
                            if (q0w + rr < p.n_q && col < p.n_valid)
                                p.dump[static_cast<uint64_t>(q0w + rr) * p.dump_ld + col] =
                                    __uint_as_float(s_stash[lane * kBlockM + lane_base + (rr ^ lane)]);
                        }
                        __syncwarp();
                        continue;  // dump launches keep no candidate lists
                    }
                    const bool hit = (vmax > thr) && (vmax >= gthr);
                    if constexpr (KREG > 0) {
                        if (hit) {
                            uint32_t mask = 0;  // exact tests, straight from registers
#pragma unroll
                            for (uint32_t j = 0; j < 32; ++j) {
                                const float v = __uint_as_float(r[j]);
                                mask |= ((v > thr) && (v >= gthr)) ? (1u << j) : 0u;
                            }
                            // park the chunk in smem so the set bits can be walked with a dynamic index
#pragma unroll
                            for (uint32_t j = 0; j < 32; ++j) s_stash[j * kBlockM + row] = r[j];
                            do {
                                const uint32_t j = __ffs(mask) - 1;
                                mask &= mask - 1;
                                const float v = __uint_as_float(s_stash[j * kBlockM + row]);
                                if (v > thr) {  // thr may have risen inside this chunk
                                    rl.insert(v, col0 + c + j);
                                    thr = rl.kth();
                                }
                            } while (mask);
                        }
                    } else {
                        uint32_t lanes = __ballot_sync(0xffffffffu, hit);
                        if (lanes) {  // warp-uniform: every lane parks its chunk, then the warp serves lane after lane
                            uint32_t mask = 0;
#pragma unroll
                            for (uint32_t j = 0; j < 32; ++j) {
                                const float v = __uint_as_float(r[j]);
                                mask |= ((v > thr) && (v >= gthr)) ? (1u << j) : 0u;
                                s_stash[j * kBlockM + row] = r[j];
                            }
                            __syncwarp();
                            do {
                                const uint32_t src = __ffs(lanes) - 1;
                                lanes &= lanes - 1;
                                uint32_t m = __shfl_sync(0xffffffffu, mask, src);
                                float thr_src = __shfl_sync(0xffffffffu, thr, src);
                                const uint32_t qrow = lane_base + src;
                                while (m) {
                                    const uint32_t j = __ffs(m) - 1;
                                    m &= m - 1;
                                    const float v = __uint_as_float(s_stash[j * kBlockM + qrow]);  // broadcast read
                                    if (v > thr_src)
                                        thr_src = coop_list_insert(s_val, s_idx, list_stride, kprime, qrow, lane, v,
                                                                   col0 + c + j);
                                }
                                if (lane == src) thr = thr_src;
                            } while (lanes);
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if constexpr (kCtas == 2) mbar_arrive_cluster(mapa_cluster(bar_tempty + 8 * buf, 0));
                    else mbar_arrive(bar_tempty + 8 * buf);
                }
                if (share && thr > gthr) atomic_max_float(p.gthr + q, thr);  // publish (only once the list is full)
            }
            // flush this (slab, query) candidate list
            const uint64_t o = (static_cast<uint64_t>(slab) * qpad + q) * kprime;
            if constexpr (kDump) {
                (void)o;  // dump launches keep no lists (part_val / part_idx may be null)
            } else if constexpr (KREG > 0) {
#pragma unroll
                for (int j = 0; j < KREG; ++j) {
                    p.part_val[o + j] = rl.v[j];
                    p.part_idx[o + j] = rl.i[j];
                }
            } else {
                __syncwarp();
                for (uint32_t r = 0; r < 32; ++r) {  // coalesced: lanes across list positions
                    const uint64_t orow = (static_cast<uint64_t>(slab) * qpad + mt * kBlockM + lane_base + r) * kprime;
                    for (uint32_t j = lane; j < kprime; j += 32) {
                        p.part_val[orow + j] = s_val[(lane_base + r) * list_stride + j];
                        p.part_idx[orow + j] = s_idx[(lane_base + r) * list_stride + j];
                    }
                }
                __syncwarp();
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (p.cta_times && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        p.cta_times[2 * blockIdx.x + 1] = t;
    }
    if constexpr (kCtas == 2) cluster_sync_all();  // the peer may still be reading our smem / signalling our barriers
    tc_fence_after();
    if (warp == 1) {
        if constexpr (kCtas == 2) tmem_dealloc_2cta<kTmemCols>(tmem_base);
        else tmem_dealloc<kTmemCols>(tmem_base);
    }
}

}  // namespace

static inline uint32_t list_smem_bytes(uint32_t kprime) {
    return (kprime == 16 || kprime == 32) ? 0u : (kprime | 1u) * kBlockM * 8;
}

size_t fused_smem_bytes(uint32_t kprime, uint32_t stages, int ctas) {
    return 1024 + static_cast<size_t>(stages) * stage_bytes(ctas) + list_smem_bytes(kprime) + kStashBytes + kAuxBytes +
           kBarBytes;
}

uint32_t fused_pick_stages(uint32_t kprime, size_t smem_optin, int ctas) {
    for (uint32_t s = max_stages(ctas); s >= 2; --s)
        if (fused_smem_bytes(kprime, s, ctas) <= smem_optin) return s;
    return 0;
}

uint32_t fused_round_kprime(uint32_t kprime) {
    if (kprime <= 16) return 16;   // register lists
    if (kprime <= 32) return 32;
    return kprime;                 // shared-memory lists (<= kFastMaxKPrime)
}

template <int kMode, int KREG, int kCtas>
static cudaError_t launch_variant(const CUtensorMap& tmap_q, const CUtensorMap& tmap_c, const FusedParams& p,
                                  uint32_t grid, size_t smem, cudaStream_t stream) {
    auto kern = fused_knn_kernel<kMode, KREG, kCtas>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kNumThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kCtas;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, tmap_q, tmap_c, p);
}

template <int kMode, int kCtas>
static cudaError_t launch_k(const CUtensorMap& tmap_q, const CUtensorMap& tmap_c, const FusedParams& p, uint32_t grid,
                            size_t smem, cudaStream_t stream) {
    if (p.kprime == 16) return launch_variant<kMode, 16, kCtas>(tmap_q, tmap_c, p, grid, smem, stream);
    if (p.kprime == 32) return launch_variant<kMode, 32, kCtas>(tmap_q, tmap_c, p, grid, smem, stream);
    return launch_variant<kMode, 0, kCtas>(tmap_q, tmap_c, p, grid, smem, stream);
}

// grid counts CTAs (a multiple of p.ctas); tmap_c must have a [256/ctas x 64] box.
cudaError_t launch_fused_knn(const CUtensorMap& tmap_q, const CUtensorMap& tmap_c, const FusedParams& p, uint32_t grid,
                             cudaStream_t stream) {
    const size_t smem = fused_smem_bytes(p.kprime, p.stages, p.ctas);
    const int mode = p.dump ? 1 : (p.q_filter ? 2 : 0);
    if (p.ctas == 2) {
        if (mode == 1) return launch_k<1, 2>(tmap_q, tmap_c, p, grid, smem, stream);
        if (mode == 2) return launch_k<2, 2>(tmap_q, tmap_c, p, grid, smem, stream);
        return launch_k<0, 2>(tmap_q, tmap_c, p, grid, smem, stream);
    }
    if (mode == 1) return launch_k<1, 1>(tmap_q, tmap_c, p, grid, smem, stream);
    if (mode == 2) return launch_k<2, 1>(tmap_q, tmap_c, p, grid, smem, stream);
    return launch_k<0, 1>(tmap_q, tmap_c, p, grid, smem, stream);
}

}  // namespace pwknn
