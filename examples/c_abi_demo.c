/* Plain-C caller of libpwknn.so: the same call sequence the Rust shim performs per timestamp --
 * remove(-diff keys) -> add(+diff rows) -> search(queries) -- with nothing but pointers and sizes.
 *
 *   gcc -O2 -Iinclude examples/c_abi_demo.c -Lpathway_b200 -lpwknn -Wl,-rpath,'$ORIGIN/../pathway_b200' -lm -o examples/c_abi_demo
 *
 * Prints "c_abi_demo ok" and exits 0 when the nearest neighbours of a few hand-made vectors come back as
 * expected (KAT of /root/reference/python/pathway/tests/external_index/test_brute_force_knn.py:252-307, L2SQ). */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "pwknn.h"

#define CHECK(x) do { int rc_ = (x); if (rc_ != PWKNN_OK) { fprintf(stderr, "%s -> %d: %s\n", #x, rc_, h ? pwknn_last_error(h) : pwknn_last_global_error()); return 1; } } while (0)

static int run(pwknn_index* h) {
    /* index rows (i, 0.1, 0.1), i = 1..9, keys {i, 0} */
    uint64_t keys[2 * 9]; double rows[3 * 9]; int32_t st[9];
    for (int i = 0; i < 9; i++) { keys[2 * i] = (uint64_t)(i + 1); keys[2 * i + 1] = 0; rows[3 * i] = i + 1; rows[3 * i + 1] = 0.1; rows[3 * i + 2] = 0.1; }
    CHECK(pwknn_add(h, 9, keys, rows, st));
    /* retract key 9, upsert key 5 with the same vector (remove-then-add inside one timestamp) */
    uint64_t rm[4] = {9, 0, 5, 0};
    CHECK(pwknn_remove(h, 2, rm, st));
    uint64_t k5[2] = {5, 0}; double r5[3] = {5, 0.1, 0.1};
    CHECK(pwknn_add(h, 1, k5, r5, st));
    if (pwknn_size(h) != 8) { fprintf(stderr, "size %llu != 8\n", (unsigned long long)pwknn_size(h)); return 1; }
    /* 4 queries (0.5, 0.1, 0.1) with limits 1..4 */
    double q[3 * 4]; uint32_t limits[4] = {1, 2, 3, 4};
    for (int i = 0; i < 4; i++) { q[3 * i] = 0.5; q[3 * i + 1] = 0.1; q[3 * i + 2] = 0.1; }
    uint64_t out_keys[2 * 10]; double out_scores[10]; uint32_t counts[4]; int32_t qst[4];
    CHECK(pwknn_search(h, 4, q, limits, out_keys, out_scores, counts, qst));
    const double want[4] = {0.25, 2.25, 6.25, 12.25};
    int off = 0;
    for (int i = 0; i < 4; i++) {
        if (counts[i] != limits[i]) { fprintf(stderr, "query %d: %u matches\n", i, counts[i]); return 1; }
        for (uint32_t r = 0; r < counts[i]; r++) {
            if (out_keys[2 * (off + r)] != r + 1 || fabs(-out_scores[off + r] - want[r]) > 0.02) {
                fprintf(stderr, "query %d rank %u: key %llu dist %f\n", i, r, (unsigned long long)out_keys[2 * (off + r)], -out_scores[off + r]);
                return 1;
            }
        }
        off += limits[i];
    }
    /* unknown key -> per-row MissingKey, call still succeeds */
    uint64_t nokey[2] = {77, 0};
    CHECK(pwknn_remove(h, 1, nokey, st));
    if (st[0] != PWKNN_ROW_MISSING_KEY) { fprintf(stderr, "expected MissingKey\n"); return 1; }
    CHECK(pwknn_destroy(h));
    return 0;
}

int main(void) {
    pwknn_index* h = NULL;
    CHECK(pwknn_create(3, 10, 1000, PWKNN_METRIC_L2SQ, 0, &h));
    if (run(h)) return 1;
    /* the same calls on one handle spread over several shards (here: two on GPU 0; on a multi-GPU box {0, 1, ...}) */
    const int devices[2] = {0, 0};
    h = NULL;
    CHECK(pwknn_create_multi(3, 10, 1000, PWKNN_METRIC_L2SQ, devices, 2, 0, &h));
    if (run(h)) return 1;
    printf("c_abi_demo ok\n");
    return 0;
}
