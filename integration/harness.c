/* harness.c -- the reference's benchmark loops (benchmarks/shared_benchmarks.zig: runInsertionBenchmark :61-88,
 * runSearchBenchmark :90-125) restated in C over the C ABI of include/zvdb_b200.h: what a COMPILED caller -- the Zig
 * binding of integration/zvdb_b200.zig -- pays per `insert(point)` / `search(query, k)` call, without the Python
 * interpreter that zvdb_b200/benchmarks.py adds. Same timed regions as the reference: a uniform [0,1) random point /
 * query is generated INSIDE the loop, one call per iteration, the result slice is built (k nodes, each with its
 * point pointer, like []const Node) and freed inside the loop; the index is filled by untimed inserts before the
 * search loop. Prints the reference's result block.
 *
 *   cc -O2 -std=c99 -Iinclude integration/harness.c -Lzvdb_b200/lib -lzvdb_b200 -Wl,-rpath,$PWD/zvdb_b200/lib -o harness
 *   ./harness [points=100000] [dim=128] [queries=10000] [k=10]
 */
#define _POSIX_C_SOURCE 199309L
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <time.h>
#include "zvdb_b200.h"

typedef struct { uint64_t id; const float *point; float distance; } node_t;   /* the fields of `Node` a caller reads */

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static float next_uniform(void) {          /* xorshift64*: stands in for std.crypto.random.float(f32) */
    rng_state ^= rng_state >> 12; rng_state ^= rng_state << 25; rng_state ^= rng_state >> 27;
    return (float)((rng_state * 0x2545F4914F6CDD1Dull) >> 40) * (1.0f / 16777216.0f);
}
static float *random_point(uint32_t dim) {  /* allocated per call, like randomPoint (:53-59) */
    float *p = (float *)malloc(sizeof(float) * dim);
    if (p) for (uint32_t i = 0; i < dim; ++i) p[i] = next_uniform();
    return p;
}
static double now_s(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec; }
static void die(const char *what) { fprintf(stderr, "%s: %s\n", what, zvdb_last_error()); exit(1); }

int main(int argc, char **argv) {
    const uint64_t points = argc > 1 ? strtoull(argv[1], 0, 10) : 100000;
    const uint32_t dim = argc > 2 ? (uint32_t)atoi(argv[2]) : 128;
    const uint64_t queries = argc > 3 ? strtoull(argv[3], 0, 10) : 10000;
    const uint32_t k = argc > 4 ? (uint32_t)atoi(argv[4]) : 10;
    zvdb_index *ix = 0;
    if (zvdb_create(&ix, 0, 16, 200, ZVDB_METRIC_L2, 0) != ZVDB_OK) die("create");

    double t0 = now_s();
    for (uint64_t i = 0; i < points; ++i) {
        float *p = random_point(dim);
        if (!p || zvdb_insert(ix, p, dim) != ZVDB_OK) die("insert");
        free(p);
    }
    double dt = now_s() - t0;
    printf("Insertion Benchmark:\n  Points: %llu\n  Dimensions: %u\n  Total time: %.2f seconds\n  Insertion per second: %.2f\n\n",
           (unsigned long long)points, dim, dt, (double)points / dt);

    uint64_t *ids = (uint64_t *)malloc(sizeof(uint64_t) * k);
    float *dist = (float *)malloc(sizeof(float) * k);
    if (!ids || !dist) return 2;
    { float *q = random_point(dim); uint32_t c = 0; if (zvdb_search(ix, q, dim, k, ids, dist, &c) != ZVDB_OK) die("search"); free(q); }  /* device copy built, untimed like the inserts */
    uint64_t found = 0;
    t0 = now_s();
    for (uint64_t i = 0; i < queries; ++i) {
        float *q = random_point(dim);
        uint32_t count = 0;
        if (!q || zvdb_search(ix, q, dim, k, ids, dist, &count) != ZVDB_OK) die("search");
        node_t *results = (node_t *)malloc(sizeof(node_t) * (count ? count : 1));   /* []const Node, freed by the caller */
        for (uint32_t r = 0; r < count; ++r) { results[r].id = ids[r]; results[r].point = zvdb_get_point(ix, ids[r]); results[r].distance = dist[r]; }
        found += count;
        free(results);
        free(q);
    }
    dt = now_s() - t0;
    printf("Search Benchmark:\n  Points: %llu\n  Dimensions: %u\n  Queries: %llu\n  k: %u\n  Total time: %.2f seconds\n  Search per second: %.2f\n",
           (unsigned long long)points, dim, (unsigned long long)queries, k, dt, (double)queries / dt);
    printf("  (results returned: %llu, %.1f us per call)\n", (unsigned long long)found, 1e6 * dt / (double)queries);
    free(ids); free(dist);
    zvdb_destroy(ix);
    return 0;
}
