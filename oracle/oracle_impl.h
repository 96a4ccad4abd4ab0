/*
 * oracle_impl.h -- type-generic body of the CPU oracle (TEST INFRASTRUCTURE ONLY).
 *
 * Included once per element type by zvdb_oracle.c with
 *     #define T        element type          (float, double, int32_t)
 *     #define SFX(x)   x##_f32 / x##_f64 / x##_i32
 * mirroring the reference's comptime-generic `HNSW(T)` (src/hnsw.zig:8).
 *
 * Every function cites the reference lines it restates. Nothing under zvdb_b200/
 * may link, import or call this file; it exists so tests/, smoke() and bench.py's
 * cpu_baseline / --impl reference legs have something to check and time against.
 */

typedef struct SFX(orc_index) {
    int dim;                /* implied by the first inserted point (hnsw.zig has no dim field) */
    int m;                  /* hnsw.zig:48  */
    int ef_construction;    /* hnsw.zig:49  stored, never read */
    size_t n, cap;          /* nodes.count() hnsw.zig:77 */
    T *pts;                 /* n * dim, owned copies (hnsw.zig:24-26) */
    int *level;             /* per node: connections.len - 1 (hnsw.zig:19) */
    uint32_t **conn;        /* per node: (level+1) lists, pitch m+1 (a list is m+1 long only transiently, hnsw.zig:129-139) */
    uint32_t **cnt;         /* per node: (level+1) list lengths */
    int has_entry;          /* entry_point: ?usize  hnsw.zig:46 */
    size_t entry;
    int max_level;          /* hnsw.zig:47 */
    uint64_t rng;           /* seeded stand-in for std.crypto.random (hnsw.zig:176); irrelevant to search */
    int dist_mode;          /* ORC_DIST_SEQ | ORC_DIST_TREE, used by insert and by search on this index */
} SFX(orc_index);

/* ---- distance: src/hnsw.zig:182-192 ---------------------------------------------------- */
/* Sequential, i ascending, diff*diff rounded then added (Zig strict float mode: no FMA
 * contraction). The translation unit is compiled with -ffp-contract=off. */
static inline T SFX(orc_dist_seq)(const T *a, const T *b, int dim) {
    T sum = 0;
    for (int i = 0; i < dim; ++i) {
        const T diff = a[i] - b[i];
        sum += diff * diff;
    }
    return sum;
}

#ifdef ORC_T_IS_F32
/* GPU summation order, restated on the CPU so the kernel can be checked bit-for-bit
 * (zvdb_b200/csrc/search_kernel.cuh: accumulate_chunk / rows_distance). The row is cut into
 * 16-byte chunks of 4 floats (x,y,z,w); lane l of a 32-lane warp owns chunks l, l+32, l+64, ...
 * and keeps two running sums: A0 takes x then z, A1 takes y then w, each by ONE FUSED multiply-add
 * per element (the kernel uses packed fma.rn.f32x2); the difference q - v is rounded once, as in
 * hnsw.zig:188. The lane's partial is A0 + A1; the 32 partials are combined by an xor butterfly
 * with offsets 16,8,4,2,1. Elements past `dim` are zero padding.
 * metric: 0 = squared L2, 1 = cosine (1 - dot, rows pre-normalised), 2 = dot (-dot). */
static inline float orc_dist_tree_metric(const float *a, const float *b, int dim, int metric) {
    float lane[32];
    const int chunks = (dim + 3) / 4;
    for (int l = 0; l < 32; ++l) {
        float acc[2] = {0.0f, 0.0f};
        for (int c = l; c < chunks; c += 32) {
            for (int e = 0; e < 4; ++e) {
                const int i = c * 4 + e;
                const float x = i < dim ? a[i] : 0.0f;
                const float y = i < dim ? b[i] : 0.0f;
                if (metric == 0) {
                    const float d = x - y;
                    acc[e & 1] = fmaf(d, d, acc[e & 1]);
                } else {
                    acc[e & 1] = fmaf(x, y, acc[e & 1]);
                }
            }
        }
        lane[l] = acc[0] + acc[1];
    }
    for (int off = 16; off >= 1; off >>= 1) {
        float nxt[32];
        for (int l = 0; l < 32; ++l) nxt[l] = lane[l] + lane[l ^ off];
        memcpy(lane, nxt, sizeof(lane));
    }
    if (metric == 1) return 1.0f - lane[0];
    if (metric == 2) return 0.0f - lane[0];
    return lane[0];
}
static inline float orc_dist_tree_f32(const float *a, const float *b, int dim) {
    return orc_dist_tree_metric(a, b, dim, 0);
}
#endif

/* Extension metrics (the reference has squared L2 only): cosine = 1 - dot on rows
 * L2-normalised at insert, dot = -dot. Sequential order, unfused, like hnsw.zig:186-190. */
static inline T SFX(orc_dot_seq)(const T *a, const T *b, int dim) {
    T sum = 0;
    for (int i = 0; i < dim; ++i) {
        const T p = a[i] * b[i];
        sum += p;
    }
    return sum;
}

/* mode = order | metric: order ORC_DIST_SEQ (reference) or ORC_DIST_TREE (GPU lane order, f32
 * only); metric ORC_METRIC_L2 (reference), ORC_METRIC_COS, ORC_METRIC_DOT. */
static inline T SFX(orc_dist)(int mode, const T *a, const T *b, int dim) {
    const int metric = (mode >> 4) & 3;
#ifdef ORC_T_IS_F32
    if (mode & ORC_DIST_TREE) return orc_dist_tree_metric(a, b, dim, metric);
#endif
    if (metric == 1) return (T)1 - SFX(orc_dot_seq)(a, b, dim);
    if (metric == 2) return (T)0 - SFX(orc_dot_seq)(a, b, dim);
    return SFX(orc_dist_seq)(a, b, dim);
}

/* ---- init / deinit: src/hnsw.zig:52-62, 64-71 -------------------------------------------- */
static SFX(orc_index) *SFX(orc_create_impl)(int m, int ef_construction, uint64_t seed) {
    SFX(orc_index) *ix = (SFX(orc_index) *)calloc(1, sizeof(*ix));
    if (!ix) return NULL;
    ix->m = m;
    ix->ef_construction = ef_construction;
    ix->has_entry = 0;         /* .entry_point = null  hnsw.zig:56 */
    ix->max_level = 0;         /* .max_level = 0       hnsw.zig:57 */
    ix->rng = seed * 0x9E3779B97F4A7C15ull + 0xD1B54A32D192ED03ull;
    ix->dist_mode = ORC_DIST_SEQ;
    return ix;
}

static void SFX(orc_destroy_impl)(SFX(orc_index) *ix) {
    if (!ix) return;
    for (size_t i = 0; i < ix->n; ++i) { free(ix->conn[i]); free(ix->cnt[i]); }
    free(ix->conn); free(ix->cnt); free(ix->level); free(ix->pts); free(ix);
}

/* ---- randomLevel: src/hnsw.zig:172-180 --------------------------------------------------- */
/* geometric p = 0.5, cap 31. splitmix64 stands in for std.crypto.random.float(f32) < 0.5. */
static int SFX(orc_random_level)(SFX(orc_index) *ix) {
    int level = 0;
    while (level < 31) {
        uint64_t z = (ix->rng += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        z ^= z >> 31;
        const float f = (float)(z >> 40) * (1.0f / 16777216.0f);
        if (!(f < 0.5f)) break;
        level += 1;
    }
    return level;
}

/* ---- shrinkConnections: src/hnsw.zig:143-170 --------------------------------------------- */
/* If the list is longer than m: stable insertion sort of the ids by distance to the node
 * (std.sort.insertion, hnsw.zig:166; x moves left while dist(x) < dist(left neighbour)),
 * keep the first m. The list is left distance-sorted. */
static void SFX(orc_shrink)(SFX(orc_index) *ix, size_t node, int level) {
    uint32_t *list = ix->conn[node] + (size_t)level * (ix->m + 1);
    uint32_t len = ix->cnt[node][level];
    if (len <= (uint32_t)ix->m) return;                       /* hnsw.zig:146 */
    const T *p = ix->pts + node * ix->dim;
    for (uint32_t i = 1; i < len; ++i) {
        const uint32_t x = list[i];
        const T dx = SFX(orc_dist)(ix->dist_mode, p, ix->pts + (size_t)x * ix->dim, ix->dim);
        uint32_t j = i;
        while (j > 0) {
            const T dl = SFX(orc_dist)(ix->dist_mode, p, ix->pts + (size_t)list[j - 1] * ix->dim, ix->dim);
            if (!(dx < dl)) break;                              /* hnsw.zig:162 strict < */
            list[j] = list[j - 1];
            --j;
        }
        list[j] = x;
    }
    ix->cnt[node][level] = (uint32_t)ix->m;                     /* hnsw.zig:168-169 */
}

/* ---- connect: src/hnsw.zig:119-141 ------------------------------------------------------- */
static void SFX(orc_connect)(SFX(orc_index) *ix, size_t source, size_t target, int level) {
    const int pitch = ix->m + 1;
    if (level <= ix->level[source])                              /* hnsw.zig:128 */
        ix->conn[source][(size_t)level * pitch + ix->cnt[source][level]++] = (uint32_t)target;
    if (level <= ix->level[target])                              /* hnsw.zig:131 */
        ix->conn[target][(size_t)level * pitch + ix->cnt[target][level]++] = (uint32_t)source;
    if (level <= ix->level[source]) SFX(orc_shrink)(ix, source, level);   /* hnsw.zig:135-137 */
    if (level <= ix->level[target]) SFX(orc_shrink)(ix, target, level);   /* hnsw.zig:138-140 */
}

/* ---- insert: src/hnsw.zig:73-117 --------------------------------------------------------- */
/* forced_level < 0 draws from randomLevel(). Returns 0, or -1 on OOM / dim mismatch. */
static int SFX(orc_insert_impl)(SFX(orc_index) *ix, const T *point, int dim, int forced_level) {
    if (ix->n == 0 && ix->dim == 0) ix->dim = dim;
    if (dim != ix->dim) return -2;                               /* @panic hnsw.zig:183-185 */
    if (ix->n == ix->cap) {
        size_t nc = ix->cap ? ix->cap * 2 : 1024;
        T *np = (T *)realloc(ix->pts, nc * (size_t)dim * sizeof(T));
        if (!np) return -1;
        ix->pts = np;
        int *nl = (int *)realloc(ix->level, nc * sizeof(int));
        if (!nl) return -1;
        ix->level = nl;
        uint32_t **ncn = (uint32_t **)realloc(ix->conn, nc * sizeof(*ncn));
        if (!ncn) return -1;
        ix->conn = ncn;
        uint32_t **nct = (uint32_t **)realloc(ix->cnt, nc * sizeof(*nct));
        if (!nct) return -1;
        ix->cnt = nct;
        ix->cap = nc;
    }
    const size_t id = ix->n;                                     /* hnsw.zig:77 */
    const int level = forced_level >= 0 ? forced_level : SFX(orc_random_level)(ix);   /* :78 */
    ix->conn[id] = (uint32_t *)malloc((size_t)(level + 1) * (ix->m + 1) * sizeof(uint32_t));
    ix->cnt[id] = (uint32_t *)calloc((size_t)(level + 1), sizeof(uint32_t));
    if (!ix->conn[id] || !ix->cnt[id]) return -1;
    ix->level[id] = level;
    memcpy(ix->pts + id * dim, point, (size_t)dim * sizeof(T)); /* owned copy, hnsw.zig:24-26 */
    ix->n = id + 1;                                              /* nodes.put hnsw.zig:82 */
    const T *np_ = ix->pts + id * dim;

    if (ix->has_entry) {                                         /* hnsw.zig:84 */
        size_t ep = ix->entry;
        T curr = SFX(orc_dist)(ix->dist_mode, np_, ix->pts + ep * dim, dim);   /* :86 */
        for (int layer = 0; layer <= ix->max_level; ++layer) {   /* ASCENDING, :88 */
            int changed = 1;
            while (changed) {                                    /* :90 */
                changed = 0;
                const size_t cur = ep;                           /* node captured before the scan, :92 */
                if (layer <= ix->level[cur]) {                   /* :93 */
                    const uint32_t *list = ix->conn[cur] + (size_t)layer * (ix->m + 1);
                    const uint32_t len = ix->cnt[cur][layer];
                    for (uint32_t t = 0; t < len; ++t) {         /* scans the WHOLE captured list, :94 */
                        const uint32_t nb = list[t];
                        const T d = SFX(orc_dist)(ix->dist_mode, np_, ix->pts + (size_t)nb * dim, dim);
                        if (d < curr) { ep = nb; curr = d; changed = 1; }   /* strict <, :97-101 */
                    }
                }
            }
            if (layer <= level) SFX(orc_connect)(ix, id, ep, layer);        /* :106-108 */
        }
    } else {
        ix->has_entry = 1;                                       /* only ever for id 0, :110-112 */
        ix->entry = id;
    }
    if (level > ix->max_level) ix->max_level = level;            /* AFTER the loop, :114-116 */
    return 0;
}

/* ---- candidate queue ---------------------------------------------------------------------
 * CandidateNode{id, distance}, ordered by distance only (hnsw.zig:238-245), held in Zig's
 * std.PriorityQueue (hnsw.zig:202,212,220). Zig's std is NOT vendored in /root/reference
 * (build.zig.zon:9 pins only minimum_zig_version 0.13.0), so the published 0.13 algorithm is
 * restated: array binary min-heap;
 *   add    = append, then siftUp: move up while child is strictly < parent;
 *   remove = take items[0], move the last item to the root, siftDown: the right child is chosen
 *            only if strictly < the left one, and sinking stops once the moving element is
 *            strictly < the chosen child (on a tie it keeps sinking).
 * These rules matter only on exact distance ties. PARITY UNPINNED on tie order: no reference
 * test or fixture depends on it.
 * ORC_HEAP_DET replaces "distance only" by the strict total order (distance, id), which is
 * what the CUDA kernel implements; it differs from ORC_HEAP_ZIG only on exact ties. */
typedef struct { uint32_t id; T d; } SFX(orc_cand);

static inline int SFX(orc_lt)(int heap_mode, SFX(orc_cand) a, SFX(orc_cand) b) {
    if (a.d < b.d) return 1;
    if (heap_mode == ORC_HEAP_DET && a.d == b.d) return a.id < b.id;
    return 0;
}

typedef struct {
    SFX(orc_cand) *items; size_t len, cap;
} SFX(orc_heap);

static int SFX(orc_heap_add)(SFX(orc_heap) *h, int mode, SFX(orc_cand) e) {
    if (h->len == h->cap) {
        size_t nc = h->cap ? h->cap * 2 : 256;
        SFX(orc_cand) *ni = (SFX(orc_cand) *)realloc(h->items, nc * sizeof(*ni));
        if (!ni) return -1;
        h->items = ni; h->cap = nc;
    }
    size_t child = h->len++;
    while (child > 0) {
        const size_t parent = (child - 1) >> 1;
        if (!SFX(orc_lt)(mode, e, h->items[parent])) break;
        h->items[child] = h->items[parent];
        child = parent;
    }
    h->items[child] = e;
    return 0;
}

static SFX(orc_cand) SFX(orc_heap_remove)(SFX(orc_heap) *h, int mode) {
    const SFX(orc_cand) top = h->items[0];
    const SFX(orc_cand) last = h->items[h->len - 1];
    h->len -= 1;
    if (h->len == 0) return top;
    size_t index = 0;
    for (;;) {
        size_t lesser = index * 2 + 1;
        if (!(lesser < h->len)) break;
        const size_t right = lesser + 1;
        if (right < h->len && SFX(orc_lt)(mode, h->items[right], h->items[lesser])) lesser = right;
        if (SFX(orc_lt)(mode, last, h->items[lesser])) break;
        h->items[index] = h->items[lesser];
        index = lesser;
    }
    h->items[index] = last;
    return top;
}

/* ---- search: src/hnsw.zig:194-236 -------------------------------------------------------- */
/* The graph is passed as a view so the same loop runs on the index's own layer 0 and on an
 * externally supplied graph (resolution 1): node i has cnt[i] (or, if cnt is
 * NULL, the entries before the first 0xFFFFFFFF) neighbours at adj[i*pitch ...].
 * Pops exactly min(k, reachable) candidates best-first from `entry`; the popped set is the
 * result (hnsw.zig:211-214); neighbours are marked visited when PUSHED (hnsw.zig:220-221);
 * the candidate heap is unbounded; the result is then stable-insertion-sorted by distance
 * (hnsw.zig:227-233).  Scratch (visited stamps + heap) is per thread. */
typedef struct {
    uint32_t *stamp; size_t stamp_n; uint32_t epoch;
    SFX(orc_heap) heap;
    SFX(orc_cand) *res; size_t res_cap;
} SFX(orc_scratch);

static void SFX(orc_scratch_free)(SFX(orc_scratch) *s) {
    free(s->stamp); free(s->heap.items); free(s->res);
    memset(s, 0, sizeof(*s));
}

static long SFX(orc_search_view)(const T *pts, int dim, size_t n, const uint32_t *adj, const uint32_t *cnt,
                                 size_t pitch, size_t cnt_stride, int has_entry, size_t entry,
                                 const T *query, size_t k, int dist_mode, int heap_mode,
                                 SFX(orc_scratch) *s, uint32_t *out_ids, T *out_d,
                                 uint32_t *out_pops, uint32_t *out_evals) {
    uint32_t pops = 0, evals = 0;
    size_t nres = 0;
    if (s->res_cap < k + 1) {
        free(s->res);
        s->res = (SFX(orc_cand) *)malloc((k + 1) * sizeof(*s->res));
        if (!s->res) return -1;
        s->res_cap = k + 1;
    }
    if (has_entry && n > 0) {                                    /* hnsw.zig:201 */
        if (s->stamp_n < n) {
            free(s->stamp);
            s->stamp = (uint32_t *)calloc(n, sizeof(uint32_t));
            if (!s->stamp) return -1;
            s->stamp_n = n; s->epoch = 0;
        }
        if (++s->epoch == 0) { memset(s->stamp, 0, s->stamp_n * sizeof(uint32_t)); s->epoch = 1; }
        s->heap.len = 0;
        SFX(orc_cand) e0 = { (uint32_t)entry, SFX(orc_dist)(dist_mode, query, pts + entry * dim, dim) };
        evals++;
        if (SFX(orc_heap_add)(&s->heap, heap_mode, e0)) return -1;      /* :208 */
        s->stamp[entry] = s->epoch;                               /* :209 */
        while (s->heap.len > 0 && nres < k) {                     /* :211 */
            const SFX(orc_cand) cur = SFX(orc_heap_remove)(&s->heap, heap_mode);   /* :212 */
            s->res[nres++] = cur;                                 /* :214 */
            pops++;
            const uint32_t *list = adj + (size_t)cur.id * pitch;
            for (size_t t = 0; t < pitch; ++t) {                  /* :216 */
                if (cnt) { if (t >= cnt[(size_t)cur.id * cnt_stride]) break; }
                else if (list[t] == 0xFFFFFFFFu) break;
                const uint32_t nb = list[t];
                if (s->stamp[nb] != s->epoch) {                   /* :217 */
                    SFX(orc_cand) c = { nb, SFX(orc_dist)(dist_mode, query, pts + (size_t)nb * dim, dim) };  /* :219 */
                    evals++;
                    if (SFX(orc_heap_add)(&s->heap, heap_mode, c)) return -1;      /* :220 */
                    s->stamp[nb] = s->epoch;                      /* :221 */
                }
            }
        }
    }
    /* std.sort.insertion by distance, stable (hnsw.zig:227-233). The reference recomputes
     * distance(query, point) in the comparator; that is the same deterministic function of the
     * same inputs as the value computed at push time, so the stored value is used. */
    for (size_t i = 1; i < nres; ++i) {
        const SFX(orc_cand) x = s->res[i];
        size_t j = i;
        while (j > 0 && x.d < s->res[j - 1].d) { s->res[j] = s->res[j - 1]; --j; }
        s->res[j] = x;
    }
    for (size_t i = 0; i < nres; ++i) { out_ids[i] = s->res[i].id; out_d[i] = s->res[i].d; }
    if (out_pops) *out_pops = pops;
    if (out_evals) *out_evals = evals;
    return (long)nres;
}

/* ---- descent over layers >= 1 (EXTENSION, no reference counterpart in search) ---------------
 * zvdb's search never leaves layer 0 (hnsw.zig:216). The kernel's optional descent (K2) is stated
 * here with the reference's own greedy walk, the loop insert runs on each layer (hnsw.zig:89-104):
 * scan the WHOLE list of the node captured before the scan (:92,:94), move to a neighbour only if
 * strictly closer (:97), repeat while a scan moved (:90); a node that lacks the layer is not
 * scanned (:93). Layers are taken from max_level down to 1, starting at `start`; the node reached
 * seeds the layer-0 search. Upper layers in flat form: node i has levels[i] lists of `m` ids
 * (0xFFFFFFFF padded) starting at list upper_base[i]. *evals counts every distance evaluated here,
 * the start node's included. */
static size_t SFX(orc_descend)(const T *pts, int dim, const uint8_t *levels, const uint32_t *upper_base,
                               const uint32_t *upper_adj, size_t m, int max_level, size_t start, const T *query,
                               int dist_mode, uint32_t *evals, T *out_dist) {
    size_t ep = start;
    T curr = SFX(orc_dist)(dist_mode, query, pts + ep * dim, dim);
    uint32_t ev = 1;
    for (int layer = max_level; layer >= 1; --layer) {
        int changed = 1;
        while (changed) {
            changed = 0;
            const size_t cur = ep;
            if (layer <= (int)levels[cur]) {
                const uint32_t *list = upper_adj + ((size_t)upper_base[cur] + (size_t)(layer - 1)) * m;
                for (size_t t = 0; t < m && list[t] != 0xFFFFFFFFu; ++t) {
                    const uint32_t nb = list[t];
                    const T d = SFX(orc_dist)(dist_mode, query, pts + (size_t)nb * dim, dim);
                    ev++;
                    if (d < curr) { ep = nb; curr = d; changed = 1; }
                }
            }
        }
    }
    if (evals) *evals = ev;
    if (out_dist) *out_dist = curr;
    return ep;
}

/* Flatten layer `layer` of the index into a padded table: adj[n*pitch] (0xFFFFFFFF padding) and
 * deg[n]; nodes without that layer get degree 0. */
static void SFX(orc_export_layer_impl)(const SFX(orc_index) *ix, int layer, size_t pitch, uint32_t *adj, uint32_t *deg) {
    for (size_t i = 0; i < ix->n; ++i) {
        uint32_t c = 0;
        if (layer <= ix->level[i]) {
            c = ix->cnt[i][layer];
            const uint32_t *list = ix->conn[i] + (size_t)layer * (ix->m + 1);
            for (uint32_t t = 0; t < c && t < pitch; ++t) adj[i * pitch + t] = list[t];
        }
        for (size_t t = c; t < pitch; ++t) adj[i * pitch + t] = 0xFFFFFFFFu;
        if (deg) deg[i] = c;
    }
}
