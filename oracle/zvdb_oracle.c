/*
 * zvdb_oracle.c -- CPU oracle for the zvdb HNSW search hot path.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT. A plain-C restatement of /root/reference/src/hnsw.zig
 * (insert :73-117, connect :119-141, shrinkConnections :143-170, randomLevel :172-180,
 * distance :182-192, search :194-236, CandidateNode :238-245), written to be checked against,
 * never shipped. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load the library this file builds (oracle/liboracle.so).
 *
 * Pinning status. Zig is not installed here, so the reference itself cannot be run
 * (oracle/_ref does not exist). The oracle is pinned against the reference's own tests
 * (src/test_hnsw.zig) restated as golden vectors G1-G9 in tests/golden/ (SURVEY 8c).
 * What stays unpinned: the ORDER IN WHICH EXACTLY-TIED distances leave the candidate heap
 * ("parity unpinned" for tie order) -- it depends on Zig std.PriorityQueue, which is not
 * vendored in the reference tree; its published 0.13 algorithm is restated in oracle_impl.h.
 * north_star exempts ties within 1e-5 relative from id/order identity.
 *
 * Build: gcc -O3 -march=x86-64-v3 -ffp-contract=off -fno-fast-math -fopenmp -shared -fPIC
 * (see oracle/Makefile). -ffp-contract=off keeps diff*diff and the add as two roundings,
 * as Zig's strict float mode does.
 */
#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_DIST_SEQ   0      /* reference order, hnsw.zig:186-190 */
#define ORC_DIST_TREE  1      /* GPU lane/butterfly order (f32 only) */
#define ORC_METRIC_L2  0x00   /* reference */
#define ORC_METRIC_COS 0x10   /* extension: 1 - dot on normalised rows */
#define ORC_METRIC_DOT 0x20   /* extension: -dot */
#define ORC_HEAP_ZIG   0      /* key = distance only, Zig 0.13 PriorityQueue tie behaviour */
#define ORC_HEAP_DET   1      /* key = (distance, id) strict total order (what the GPU implements) */

#define ORC_CAT_(a, b) a##b
#define ORC_CAT(a, b) ORC_CAT_(a, b)

#define T float
#define SFX(x) ORC_CAT(x, _f32)
#define ORC_T_IS_F32 1
#include "oracle_impl.h"
#undef ORC_T_IS_F32
#undef SFX
#undef T

#define T double
#define SFX(x) ORC_CAT(x, _f64)
#include "oracle_impl.h"
#undef SFX
#undef T

#define T int32_t
#define SFX(x) ORC_CAT(x, _i32)
#include "oracle_impl.h"
#undef SFX
#undef T

#define EXPORT __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------
 * Exported C API (ctypes-friendly). One block per element type.
 * ------------------------------------------------------------------------------------------ */
#define ORC_DEFINE_API(T, S)                                                                          \
    EXPORT void *orc_create_##S(int m, int efc, uint64_t seed) { return orc_create_impl_##S(m, efc, seed); } \
    EXPORT void orc_destroy_##S(void *ix) { orc_destroy_impl_##S((orc_index_##S *)ix); }              \
    EXPORT void orc_set_dist_mode_##S(void *ix, int mode) { ((orc_index_##S *)ix)->dist_mode = mode; } \
    EXPORT int orc_insert_##S(void *ix, const T *p, int dim, int forced_level) {                      \
        return orc_insert_impl_##S((orc_index_##S *)ix, p, dim, forced_level);                        \
    }                                                                                                 \
    EXPORT int orc_insert_batch_##S(void *ix, const T *p, size_t n, int dim, const int *levels) {     \
        for (size_t i = 0; i < n; ++i) {                                                              \
            int rc = orc_insert_impl_##S((orc_index_##S *)ix, p + i * (size_t)dim, dim,               \
                                         levels ? levels[i] : -1);                                    \
            if (rc) return rc;                                                                        \
        }                                                                                             \
        return 0;                                                                                     \
    }                                                                                                 \
    EXPORT size_t orc_count_##S(const void *ix) { return ((const orc_index_##S *)ix)->n; }            \
    EXPORT int orc_dim_##S(const void *ix) { return ((const orc_index_##S *)ix)->dim; }               \
    EXPORT int orc_max_level_##S(const void *ix) { return ((const orc_index_##S *)ix)->max_level; }   \
    EXPORT long orc_entry_##S(const void *ix) {                                                       \
        const orc_index_##S *x = (const orc_index_##S *)ix;                                           \
        return x->has_entry ? (long)x->entry : -1;                                                    \
    }                                                                                                 \
    EXPORT int orc_level_##S(const void *ix, size_t i) { return ((const orc_index_##S *)ix)->level[i]; } \
    EXPORT const T *orc_points_##S(const void *ix) { return ((const orc_index_##S *)ix)->pts; }       \
    EXPORT void orc_export_layer_##S(const void *ix, int layer, size_t pitch, uint32_t *adj, uint32_t *deg) { \
        orc_export_layer_impl_##S((const orc_index_##S *)ix, layer, pitch, adj, deg);                   \
    }                                                                                                 \
    /* search(query, k) on the index's own layer 0: the reference call, hnsw.zig:194. */             \
    EXPORT long orc_search_##S(const void *ixv, const T *q, size_t k, int dist_mode, int heap_mode,   \
                               uint32_t *ids, T *d, uint32_t *pops, uint32_t *evals) {                \
        const orc_index_##S *ix = (const orc_index_##S *)ixv;                                         \
        const size_t pitch = (size_t)ix->m + 1;                                                       \
        uint32_t *adj = (uint32_t *)malloc((ix->n ? ix->n : 1) * pitch * sizeof(uint32_t));           \
        if (!adj) return -1;                                                                          \
        orc_export_layer_impl_##S(ix, 0, pitch, adj, NULL);                                             \
        orc_scratch_##S sc; memset(&sc, 0, sizeof(sc));                                               \
        long r = orc_search_view_##S(ix->pts, ix->dim, ix->n, adj, NULL, pitch, 0, ix->has_entry,     \
                                     ix->entry, q, k, dist_mode, heap_mode, &sc, ids, d, pops, evals); \
        orc_scratch_free_##S(&sc); free(adj);                                                         \
        return r;                                                                                     \
    }                                                                                                 \
    /* Batched search(q, ef)[0..k] on a supplied padded graph (adj[n*pitch], 0xFFFFFFFF padding),     \
     * one query per OpenMP thread, read-only, no lock. nthreads <= 0: all. global_lock != 0          \
     * serialises the calls like the reference's mutex (hnsw.zig:195-196). */                         \
    EXPORT int orc_search_graph_desc_##S(const T *pts, int dim, size_t n, const uint32_t *adj, size_t pitch, \
                                    long entry, const T *queries, size_t nq, size_t ef, size_t k,     \
                                    int dist_mode, int heap_mode, int nthreads, int global_lock,      \
                                    const uint8_t *levels, const uint32_t *upper_base,                \
                                    const uint32_t *upper_adj, int max_level, long start,             \
                                    uint32_t *ids, T *d, uint32_t *counts, uint32_t *pops,            \
                                    uint32_t *evals) {                                                \
        int err = 0;                                                                                  \
        if (k > ef) k = ef;                                                                           \
        const int descend = levels && max_level > 0 && start >= 0 && n > 0;                           \
        _Pragma("omp parallel num_threads(nthreads > 0 ? nthreads : omp_get_max_threads())")          \
        {                                                                                             \
            /* per-thread scratch (visited stamps, heap) lives across calls: a 4n-byte stamp array    \
             * per call would dominate short timed samples */                                         \
            static __thread orc_scratch_##S sc;                                                       \
            uint32_t *tid = (uint32_t *)malloc((ef + 1) * sizeof(uint32_t));                          \
            T *td = (T *)malloc((ef + 1) * sizeof(T));                                                \
            _Pragma("omp for schedule(dynamic, 8)")                                                   \
            for (long qi = 0; qi < (long)nq; ++qi) {                                                  \
                long r;                                                                               \
                uint32_t p = 0, e = 0, de = 0;                                                        \
                long en = entry;                                                                      \
                if (descend) {                                                                        \
                    en = (long)orc_descend_##S(pts, dim, levels, upper_base, upper_adj, pitch, max_level, \
                                               (size_t)start, queries + (size_t)qi * dim, dist_mode, &de, NULL); \
                    de -= 1;   /* the beam below evaluates its entry again; count that row once */    \
                }                                                                                     \
                if (global_lock) {                                                                    \
                    _Pragma("omp critical(orc_global_lock)")                                          \
                    r = orc_search_view_##S(pts, dim, n, adj, NULL, pitch, 0, en >= 0, (size_t)en,    \
                                            queries + (size_t)qi * dim, ef, dist_mode, heap_mode, &sc, \
                                            tid, td, &p, &e);                                         \
                } else {                                                                              \
                    r = orc_search_view_##S(pts, dim, n, adj, NULL, pitch, 0, en >= 0, (size_t)en,    \
                                            queries + (size_t)qi * dim, ef, dist_mode, heap_mode, &sc, \
                                            tid, td, &p, &e);                                         \
                }                                                                                     \
                if (r < 0) { err = 1; r = 0; }                                                        \
                const size_t nr = (size_t)r < k ? (size_t)r : k;                                      \
                for (size_t j = 0; j < nr; ++j) { ids[qi * k + j] = tid[j]; d[qi * k + j] = td[j]; }  \
                for (size_t j = nr; j < k; ++j) { ids[qi * k + j] = 0xFFFFFFFFu; d[qi * k + j] = (T)0; } \
                if (counts) counts[qi] = (uint32_t)nr;                                                \
                if (pops) pops[qi] = p;                                                               \
                if (evals) evals[qi] = e + de;                                                        \
            }                                                                                         \
            free(tid); free(td);                                                                      \
        }                                                                                             \
        return err ? -1 : 0;                                                                          \
    }                                                                                                 \
    EXPORT int orc_search_graph_##S(const T *pts, int dim, size_t n, const uint32_t *adj, size_t pitch, \
                                    long entry, const T *queries, size_t nq, size_t ef, size_t k,     \
                                    int dist_mode, int heap_mode, int nthreads, int global_lock,      \
                                    uint32_t *ids, T *d, uint32_t *counts, uint32_t *pops,            \
                                    uint32_t *evals) {                                                \
        return orc_search_graph_desc_##S(pts, dim, n, adj, pitch, entry, queries, nq, ef, k, dist_mode, \
                                         heap_mode, nthreads, global_lock, NULL, NULL, NULL, 0, -1,   \
                                         ids, d, counts, pops, evals);                                \
    }                                                                                                 \
    /* Where the descent lands for one query (test hook): node id, its distance, evaluations. */      \
    EXPORT long orc_descend_one_##S(const T *pts, int dim, const uint8_t *levels, const uint32_t *upper_base, \
                                    const uint32_t *upper_adj, size_t m, int max_level, long start,   \
                                    const T *query, int dist_mode, uint32_t *evals, T *dist) {        \
        return (long)orc_descend_##S(pts, dim, levels, upper_base, upper_adj, m, max_level, (size_t)start, \
                                     query, dist_mode, evals, dist);                                  \
    }

#ifndef _OPENMP
static int omp_get_max_threads(void) { return 1; }
#endif

ORC_DEFINE_API(float, f32)
ORC_DEFINE_API(double, f64)
ORC_DEFINE_API(int32_t, i32)

EXPORT int orc_max_threads(void) { return omp_get_max_threads(); }

/* Single-pair distance, exposed so tests can pin the two summation orders. */
EXPORT float orc_distance_f32(const float *a, const float *b, int dim, int mode) {
    return orc_dist_f32(mode, a, b, dim);
}

/* Distances from one vector to a list of rows (used by tests/builder_ref.py). */
EXPORT void orc_dist_many_f32(const float *a, const float *pts, int dim, const uint32_t *ids, size_t count,
                              int mode, float *out) {
    for (size_t i = 0; i < count; ++i) out[i] = orc_dist_f32(mode, a, pts + (size_t)ids[i] * dim, dim);
}

/* ------------------------------------------------------------------------------------------
 * Exact k-NN oracle for K4 (no reference counterpart). Distances accumulated in
 * double from the f32 inputs, sequentially, then rounded to f32; ordering (distance, id).
 * metric: 0 squared L2, 1 cosine (1 - dot), 2 dot (-dot).
 * ------------------------------------------------------------------------------------------ */
EXPORT int orc_bruteforce_f32(const float *pts, size_t n, int dim, const float *queries, size_t nq,
                              size_t k, int metric, int nthreads, uint32_t *ids, float *d) {
    if (k > n) k = n;
    int err = 0;
#pragma omp parallel num_threads(nthreads > 0 ? nthreads : omp_get_max_threads())
    {
        double *bd = (double *)malloc((k + 1) * sizeof(double));
        uint32_t *bi = (uint32_t *)malloc((k + 1) * sizeof(uint32_t));
        if (!bd || !bi) err = 1;
#pragma omp for schedule(dynamic, 4)
        for (long qi = 0; qi < (long)nq; ++qi) {
            if (err) continue;
            const float *q = queries + (size_t)qi * dim;
            size_t len = 0;
            for (size_t i = 0; i < n; ++i) {
                const float *p = pts + i * (size_t)dim;
                double s = 0.0;
                if (metric == 0) {
                    for (int t = 0; t < dim; ++t) { const double df = (double)q[t] - (double)p[t]; s += df * df; }
                } else {
                    for (int t = 0; t < dim; ++t) s += (double)q[t] * (double)p[t];
                    s = metric == 1 ? 1.0 - s : -s;
                }
                if (len == k && !(s < bd[len - 1])) continue;   /* ids ascend, so ties keep the earlier id */
                size_t j = len < k ? len++ : len - 1;
                while (j > 0 && s < bd[j - 1]) { bd[j] = bd[j - 1]; bi[j] = bi[j - 1]; --j; }
                bd[j] = s; bi[j] = (uint32_t)i;
            }
            for (size_t j = 0; j < k; ++j) {
                ids[qi * k + j] = j < len ? bi[j] : 0xFFFFFFFFu;
                d[qi * k + j] = j < len ? (float)bd[j] : 0.0f;
            }
        }
        free(bd); free(bi);
    }
    return err ? -1 : 0;
}

/* ------------------------------------------------------------------------------------------
 * Oracle for the shard merge (K5): G per-shard result lists of k (distance, global id)
 * pairs with counts[g*nq+q] valid entries each; output the k smallest of their union under the
 * strict order (distance, global id). A shard's own list is ordered by (distance, pop order), so
 * exact ties inside it need not be id-ordered: the union is sorted, not head-merged.
 * Layout of the gathered input: [G][nq][k].
 * ------------------------------------------------------------------------------------------ */
typedef struct { float d; uint64_t id; } orc_pair;
static int orc_pair_cmp(const void *a, const void *b) {
    const orc_pair *x = (const orc_pair *)a, *y = (const orc_pair *)b;
    if (x->d < y->d) return -1;
    if (x->d > y->d) return 1;
    return x->id < y->id ? -1 : (x->id > y->id ? 1 : 0);
}
EXPORT void orc_merge_topk(const float *d_in, const uint64_t *id_in, const uint32_t *cnt_in, int G,
                           size_t nq, size_t k, float *d_out, uint64_t *id_out, uint32_t *cnt_out) {
    orc_pair *buf = (orc_pair *)malloc((size_t)G * k * sizeof(orc_pair) + 1);
    for (size_t q = 0; q < nq; ++q) {
        size_t len = 0;
        for (int g = 0; g < G; ++g) {
            const size_t base = (size_t)g * nq + q;
            const size_t c = cnt_in[base] < k ? cnt_in[base] : k;
            for (size_t j = 0; j < c; ++j) { buf[len].d = d_in[base * k + j]; buf[len].id = id_in[base * k + j]; ++len; }
        }
        qsort(buf, len, sizeof(orc_pair), orc_pair_cmp);
        if (len > k) len = k;
        for (size_t j = 0; j < len; ++j) { d_out[q * k + j] = buf[j].d; id_out[q * k + j] = buf[j].id; }
        for (size_t j = len; j < k; ++j) { d_out[q * k + j] = 0.0f; id_out[q * k + j] = ~0ull; }
        cnt_out[q] = (uint32_t)len;
    }
    free(buf);
}
