/*
 * C restatement (f64, CPU) of the reference's brute-force KNN index.
 *
 * TEST INFRASTRUCTURE ONLY: the checker for tests/ and the timed CPU baseline
 * of bench.py (`cpu_baseline`, `--impl reference`).  Never linked into or
 * called from the product library (pathway_b200/).
 *
 * Follows, line by line, /root/reference/src/external_integration/
 *   brute_force_knn_integration.rs:33-239  (storage, growth/shrink, search)
 *   mod.rs:209-256                         (KeyToU64IdMapper)
 * The reference is Rust (ndarray 0.15.6 -> matrixmultiply 0.3.9 dgemm, single
 * threaded: Cargo.toml:67 enables only `serde`) and cannot be built in this
 * image (no cargo/rustc), so the GEMM here is a plain C loop nest compiled with
 * -O3 -march=native (substitution stated in DESIGN.md).  Parity of this file is
 * pinned by tests/test_oracle_golden.py (reference known-answer tests) and by
 * tests/test_oracle_c_vs_numpy.py (agreement with the numpy restatement).
 *
 * Build: see oracle/Makefile  ->  oracle/_build/libpworacle.so
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

typedef struct { uint64_t lo, hi; } pwo_key;

/* ---- KeyToU64IdMapper (mod.rs:209-256): key->id open-addressing table + dense id->key ---- */
typedef struct {
    pwo_key *keys; uint64_t *ids; uint8_t *state; /* 0 empty, 1 used, 2 tomb */
    size_t cap, used, tombs;
    pwo_key *id_to_key; uint8_t *id_live; size_t id_cap;
    uint64_t next_id;
} pwo_mapper;

static uint64_t mix64(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return x;
}
static size_t key_hash(pwo_key k) { return (size_t)mix64(k.lo ^ mix64(k.hi + 0x9e3779b97f4a7c15ULL)); }
static int key_eq(pwo_key a, pwo_key b) { return a.lo == b.lo && a.hi == b.hi; }

static void mapper_init(pwo_mapper *m) {
    memset(m, 0, sizeof(*m));
    m->cap = 64;
    m->keys = calloc(m->cap, sizeof(pwo_key)); m->ids = calloc(m->cap, sizeof(uint64_t));
    m->state = calloc(m->cap, 1);
}
static void mapper_free(pwo_mapper *m) {
    free(m->keys); free(m->ids); free(m->state); free(m->id_to_key); free(m->id_live);
}
static void mapper_raw_insert(pwo_mapper *m, pwo_key k, uint64_t id) {
    size_t h = key_hash(k) & (m->cap - 1);
    size_t first_tomb = (size_t)-1;
    for (;;) {
        if (m->state[h] == 0) break;
        if (m->state[h] == 2) { if (first_tomb == (size_t)-1) first_tomb = h; }
        else if (key_eq(m->keys[h], k)) { m->ids[h] = id; return; }
        h = (h + 1) & (m->cap - 1);
    }
    if (first_tomb != (size_t)-1) { h = first_tomb; m->tombs--; }
    m->keys[h] = k; m->ids[h] = id; m->state[h] = 1; m->used++;
}
static void mapper_grow(pwo_mapper *m) {
    pwo_key *ok = m->keys; uint64_t *oi = m->ids; uint8_t *os = m->state; size_t oc = m->cap;
    m->cap = (m->used * 4 > oc) ? oc * 2 : oc;
    m->keys = calloc(m->cap, sizeof(pwo_key)); m->ids = calloc(m->cap, sizeof(uint64_t));
    m->state = calloc(m->cap, 1); m->used = 0; m->tombs = 0;
    for (size_t i = 0; i < oc; i++) if (os[i] == 1) mapper_raw_insert(m, ok[i], oi[i]);
    free(ok); free(oi); free(os);
}
static int mapper_find(const pwo_mapper *m, pwo_key k, uint64_t *id) {
    size_t h = key_hash(k) & (m->cap - 1);
    while (m->state[h] != 0) {
        if (m->state[h] == 1 && key_eq(m->keys[h], k)) { *id = m->ids[h]; return 1; }
        h = (h + 1) & (m->cap - 1);
    }
    return 0;
}
static void mapper_set_id_key(pwo_mapper *m, uint64_t id, pwo_key k) {
    if (id >= m->id_cap) {
        size_t nc = m->id_cap ? m->id_cap : 64;
        while (nc <= id) nc *= 2;
        m->id_to_key = realloc(m->id_to_key, nc * sizeof(pwo_key));
        m->id_live = realloc(m->id_live, nc);
        memset(m->id_live + m->id_cap, 0, nc - m->id_cap);
        m->id_cap = nc;
    }
    m->id_to_key[id] = k; m->id_live[id] = 1;
}
/* mod.rs:224-233 */
static uint64_t mapper_get_next_free_u64_id(pwo_mapper *m, pwo_key k) {
    uint64_t id;
    if (mapper_find(m, k, &id)) return id;
    id = m->next_id++;
    if ((m->used + m->tombs + 1) * 2 > m->cap) mapper_grow(m);
    mapper_raw_insert(m, k, id);
    mapper_set_id_key(m, id, k);
    return id;
}
/* mod.rs:235-238 */
static void mapper_assign_key(pwo_mapper *m, pwo_key k, uint64_t id) {
    if ((m->used + m->tombs + 1) * 2 > m->cap) mapper_grow(m);
    mapper_raw_insert(m, k, id);
    mapper_set_id_key(m, id, k);
}
/* mod.rs:244-246 */
static int mapper_get_key_for_id(const pwo_mapper *m, uint64_t id, pwo_key *k) {
    if (id >= m->id_cap || !m->id_live[id]) return 0;
    *k = m->id_to_key[id]; return 1;
}
/* mod.rs:248-255 */
static int mapper_remove_key(pwo_mapper *m, pwo_key k, uint64_t *id) {
    size_t h = key_hash(k) & (m->cap - 1);
    while (m->state[h] != 0) {
        if (m->state[h] == 1 && key_eq(m->keys[h], k)) {
            *id = m->ids[h]; m->state[h] = 2; m->used--; m->tombs++;
            if (*id < m->id_cap) m->id_live[*id] = 0;
            return 1;
        }
        h = (h + 1) & (m->cap - 1);
    }
    return 0;
}

/* ---- BruteForceKNNIndex (brute_force_knn_integration.rs:22-51) ---- */
typedef struct {
    double *index_array;
    size_t current_size, current_allocated, minimum_allocated, auxiliary_space, dimensions;
    int metric; /* 0 = L2sq, 1 = Cos */
    pwo_mapper mapper;
} pwo_index;

pwo_index *pwo_create(size_t dimensions, size_t reserved_space, size_t auxiliary_space, int metric) {
    pwo_index *ix = calloc(1, sizeof(pwo_index));
    ix->index_array = calloc(reserved_space * dimensions + 1, sizeof(double)); /* Array2::default */
    ix->current_allocated = reserved_space; ix->minimum_allocated = reserved_space;
    ix->auxiliary_space = auxiliary_space; ix->dimensions = dimensions; ix->metric = metric;
    mapper_init(&ix->mapper);
    return ix;
}
void pwo_destroy(pwo_index *ix) { if (!ix) return; free(ix->index_array); mapper_free(&ix->mapper); free(ix); }
size_t pwo_size(const pwo_index *ix) { return ix->current_size; }
size_t pwo_allocated(const pwo_index *ix) { return ix->current_allocated; }

static void realloc_rows(pwo_index *ix, size_t new_allocated) {
    /* :121-126 / :178-184: new zeroed array, flat prefix copy */
    size_t D = ix->dimensions;
    double *na = calloc(new_allocated * D + 1, sizeof(double));
    size_t ncopy = (new_allocated < ix->current_allocated ? new_allocated : ix->current_allocated) * D;
    memcpy(na, ix->index_array, ncopy * sizeof(double));
    free(ix->index_array); ix->index_array = na; ix->current_allocated = new_allocated;
}

/* :114-144.  rows: n x row_len row-major (row_len may differ from dimensions: zip semantics). */
void pwo_add(pwo_index *ix, size_t n, const uint64_t *keys, const double *rows, size_t row_len, int32_t *status) {
    if (n + ix->current_size > ix->current_allocated) {
        size_t na = 2 * ix->current_allocated;
        if (na < n + ix->current_size) na = n + ix->current_size;
        realloc_rows(ix, na);
    }
    ix->current_size += n;
    size_t m = row_len < ix->dimensions ? row_len : ix->dimensions;
    for (size_t i = 0; i < n; i++) {
        pwo_key k = { keys[2 * i], keys[2 * i + 1] };
        uint64_t idx = mapper_get_next_free_u64_id(&ix->mapper, k);
        memcpy(ix->index_array + idx * ix->dimensions, rows + i * row_len, m * sizeof(double));
        if (status) status[i] = 0;
    }
}

/* :146-187.  status: 0 ok, 1 MissingKey.  returns -1 if the reference would panic. */
int pwo_remove(pwo_index *ix, size_t n, const uint64_t *keys, int32_t *status) {
    size_t D = ix->dimensions;
    for (size_t i = 0; i < n; i++) {
        pwo_key k = { keys[2 * i], keys[2 * i + 1] }, last;
        if (ix->current_size == 0 || !mapper_get_key_for_id(&ix->mapper, ix->current_size - 1, &last)) return -1;
        uint64_t rid;
        if (!mapper_remove_key(&ix->mapper, k, &rid)) { if (status) status[i] = 1; continue; }
        ix->current_size -= 1;
        ix->mapper.next_id -= 1; /* decrement_next_free_id */
        if (!key_eq(last, k)) {
            memcpy(ix->index_array + rid * D, ix->index_array + ix->current_size * D, D * sizeof(double));
            mapper_assign_key(&ix->mapper, last, rid);
        }
        if (status) status[i] = 0;
    }
    if (4 * ix->current_size < ix->current_allocated && ix->current_allocated / 2 >= ix->minimum_allocated)
        realloc_rows(ix, ix->current_allocated / 2);
    return 0;
}

/* (OrderedFloat(d), idx) total order: NaN greatest (ordered-float), then idx -- :220-224 */
typedef struct { double d; size_t i; } pwo_cand;
static inline int cand_less(pwo_cand a, pwo_cand b) {
    int an = isnan(a.d), bn = isnan(b.d);
    if (an != bn) return bn;            /* non-NaN < NaN */
    if (!an && a.d != b.d) return a.d < b.d;
    return a.i < b.i;
}
static void heap_sift_down(pwo_cand *h, size_t n, size_t p) { /* max-heap on cand_less */
    for (;;) {
        size_t l = 2 * p + 1, r = l + 1, g = p;
        if (l < n && cand_less(h[g], h[l])) g = l;
        if (r < n && cand_less(h[g], h[r])) g = r;
        if (g == p) return;
        pwo_cand t = h[g]; h[g] = h[p]; h[p] = t; p = g;
    }
}
static int cand_cmp_qsort(const void *a, const void *b) {
    const pwo_cand *x = a, *y = b;
    return cand_less(*x, *y) ? -1 : (cand_less(*y, *x) ? 1 : 0);
}
/* k_smallest(limit): heap of size k over the iterator, ascending output */
static size_t k_smallest(const double *col, size_t stride, size_t n, size_t limit, pwo_cand *heap) {
    if (limit == 0) return 0;
    size_t hn = 0;
    for (size_t i = 0; i < n; i++) {
        pwo_cand c = { col[i * stride], i };
        if (hn < limit) {
            heap[hn++] = c;
            if (hn == limit) for (size_t p = hn / 2; p-- > 0;) heap_sift_down(heap, hn, p);
        } else if (cand_less(c, heap[0])) {
            heap[0] = c; heap_sift_down(heap, hn, 0);
        }
    }
    qsort(heap, hn, sizeof(pwo_cand), cand_cmp_qsort);
    return hn;
}

/* one worker's `search` over queries [q0,q1) -- :189-238.  queries: nq x dims row-major f64.
 * out_keys: 2*sum(limits) u64, out_scores: sum(limits) f64, written at offsets[q]. */
static void search_range(const pwo_index *ix, const double *queries, const uint32_t *limits,
                         const uint64_t *offsets, size_t q0, size_t q1,
                         uint64_t *out_keys, double *out_scores, uint32_t *out_counts) {
    const size_t N = ix->current_size, D = ix->dimensions;
    if (N == 0) { for (size_t q = q0; q < q1; q++) out_counts[q] = 0; return; }   /* :193-198 */
    size_t max_queries = ix->auxiliary_space / N; if (max_queries < 1) max_queries = 1;  /* :205 */
    const double *A = ix->index_array;
    double *dot_p = malloc(N * max_queries * sizeof(double));
    double *index_sq = malloc(N * sizeof(double));
    double *query_sq = malloc(max_queries * sizeof(double));
    uint32_t maxlim = 0; for (size_t q = q0; q < q1; q++) if (limits[q] > maxlim) maxlim = limits[q];
    pwo_cand *heap = malloc(((size_t)maxlim + 1) * sizeof(pwo_cand));
    for (size_t b = q0; b < q1; b += max_queries) {
        size_t nb = (q1 - b < max_queries) ? q1 - b : max_queries;
        /* :215  dot_p[N x nb] = index[N x D] . Q^T */
        for (size_t i = 0; i < N; i++) {
            const double *row = A + i * D;
            for (size_t j = 0; j < nb; j++) {
                const double *qv = queries + (b + j) * D;
                double acc = 0.0;
#pragma omp simd reduction(+ : acc)
                for (size_t d = 0; d < D; d++) acc += row[d] * qv[d];
                dot_p[i * nb + j] = acc;
            }
        }
        /* :75-83 / :97-104  norms recomputed for every chunk */
        for (size_t i = 0; i < N; i++) {
            const double *row = A + i * D; double acc = 0.0;
#pragma omp simd reduction(+ : acc)
            for (size_t d = 0; d < D; d++) acc += row[d] * row[d];
            index_sq[i] = acc;
        }
        for (size_t j = 0; j < nb; j++) {
            const double *qv = queries + (b + j) * D; double acc = 0.0;
            for (size_t d = 0; d < D; d++) acc += qv[d] * qv[d];
            query_sq[j] = acc;
        }
        if (ix->metric == 1) {  /* :85-89 */
            for (size_t i = 0; i < N; i++) for (size_t j = 0; j < nb; j++)
                dot_p[i * nb + j] = 1.0 - dot_p[i * nb + j] / sqrt(index_sq[i] * query_sq[j]);
        } else {                /* :106-110 */
            for (size_t i = 0; i < N; i++) for (size_t j = 0; j < nb; j++)
                dot_p[i * nb + j] = index_sq[i] + query_sq[j] - 2.0 * dot_p[i * nb + j];
        }
        for (size_t j = 0; j < nb; j++) {  /* :218-232 */
            size_t q = b + j;
            size_t got = k_smallest(dot_p + j, nb, N, limits[q], heap);
            for (size_t r = 0; r < got; r++) {
                pwo_key k = {0, 0};
                mapper_get_key_for_id(&ix->mapper, heap[r].i, &k);
                out_keys[2 * (offsets[q] + r)] = k.lo; out_keys[2 * (offsets[q] + r) + 1] = k.hi;
                out_scores[offsets[q] + r] = -heap[r].d;  /* score = -distance, :230 */
            }
            out_counts[q] = (uint32_t)got;
        }
    }
    free(dot_p); free(index_sq); free(query_sq); free(heap);
}

/* `workers` emulates PATHWAY_THREADS: each worker (a pthread) answers a contiguous 1/W of the
 * queries against a SHARED read-only index (the reference holds W replicas; external_index.rs:92). */
typedef struct {
    const pwo_index *ix; const double *queries; const uint32_t *limits; const uint64_t *offsets;
    size_t q0, q1; uint64_t *out_keys; double *out_scores; uint32_t *out_counts;
} pwo_job;
static void *worker_main(void *p) {
    pwo_job *j = p;
    search_range(j->ix, j->queries, j->limits, j->offsets, j->q0, j->q1, j->out_keys, j->out_scores, j->out_counts);
    return NULL;
}
void pwo_search(const pwo_index *ix, size_t nq, const double *queries, const uint32_t *limits,
                uint64_t *out_keys, double *out_scores, uint32_t *out_counts, int workers) {
    uint64_t *offsets = malloc((nq + 1) * sizeof(uint64_t));
    offsets[0] = 0; for (size_t q = 0; q < nq; q++) offsets[q + 1] = offsets[q] + limits[q];
    if (workers < 1) workers = 1;
    pthread_t *th = malloc(sizeof(pthread_t) * (size_t)workers);
    pwo_job *jobs = malloc(sizeof(pwo_job) * (size_t)workers);
    for (int w = 0; w < workers; w++) {
        pwo_job j = { ix, queries, limits, offsets, nq * (size_t)w / (size_t)workers,
                      nq * (size_t)(w + 1) / (size_t)workers, out_keys, out_scores, out_counts };
        jobs[w] = j;
        if (workers == 1) worker_main(&jobs[w]);
        else pthread_create(&th[w], NULL, worker_main, &jobs[w]);
    }
    if (workers > 1) for (int w = 0; w < workers; w++) pthread_join(th[w], NULL);
    free(th); free(jobs); free(offsets);
}

/* Raw row access for tests (storage-layout parity with the numpy restatement). */
const double *pwo_rows(const pwo_index *ix) { return ix->index_array; }
int pwo_key_for_slot(const pwo_index *ix, uint64_t slot, uint64_t *lo, uint64_t *hi) {
    pwo_key k; if (!mapper_get_key_for_id(&ix->mapper, slot, &k)) return 0; *lo = k.lo; *hi = k.hi; return 1;
}
int pwo_max_threads(void) { long n = sysconf(_SC_NPROCESSORS_ONLN); return n < 1 ? 1 : (int)n; }
