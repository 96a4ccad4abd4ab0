// search_kernel.cuh -- K1: batched layer-0 best-first search, one CTA ("team" of W warps) per query.
//
// Replaces the reference's search loop (src/hnsw.zig:201-224), its distance (:182-192) and the
// result sort (:227-233), for nq queries at once:
//
//   result(q) = search(q, ef)[0..k]      (hnsw.zig:194; the reference has no ef, its pop count
//                                         is the knob -- SURVEY S2; ef = k is the reference call)
//
// Semantics kept from the reference:
//   * start at `entry` (always node 0 in the reference, hnsw.zig:110-112), layer 0 only (:216);
//   * pop the best candidate, append it to the result, push every not-yet-visited neighbour with
//     its distance, mark it visited when PUSHED (:211-221); stop after ef pops or when empty;
//   * the popped set, stable-sorted by distance over pop order, is the result (:227-233).
// What is ours:
//   * candidate order is the strict total order (distance, id) -- the reference orders by distance
//     only (:238-245) and lets its heap layout decide exact ties;
//   * the unbounded heap is replaced by a sorted list that keeps only the best (ef - pops)
//     candidates: at most that many more pops can happen, so nothing that could be popped is lost;
//     dropped nodes stay in the visited set, as they would in the reference;
//   * the visited set is an exact open-addressing hash table in shared memory (never lossy);
//   * distance is summed in lane order + xor butterfly (see row_distance), not sequentially: it
//     differs from the reference by a few ulp (oracle mode ORC_DIST_TREE mirrors it bit for bit).
//
// Memory traffic per query (the HBM-gather roofline numerator):
//   evals * row_bytes (row gathers) + pops * m * 4 (adjacency rows) + dim*4 (query) + k*12 (output).
#pragma once
#include "common.cuh"

namespace zvdb {

struct SearchParams {
    const float4 *arena;     // [n][row_chunks] float4, rows 128-byte aligned, zero padded
    const uint32_t *adj;     // [n][m] layer-0 neighbour ids, kInvalidId padded at the tail
    const float *queries;    // [nq][dim]
    uint64_t *ids;           // [nq][k]  global ids (id * id_stride + id_base), ~0 when unused
    float *dist;             // [nq][k]
    uint32_t *counts;        // [nq]
    uint32_t *pops;          // [nq] or null
    uint32_t *evals;         // [nq] or null
    uint64_t id_stride, id_base;
    uint32_t row_chunks;     // float4 per arena row
    uint32_t m, n, entry, dim, nq, k, ef;
    uint32_t slots;          // visited-table slots (> max entries)
    uint32_t hash_words;     // words reserved for the table (>= slots, >= 2*next_pow2(ef): reused by the final sort)
};

enum : int { kMetricL2 = 0, kMetricCos = 1, kMetricDot = 2 };

// Rows a warp fetches before it starts reducing (loads in flight per lane = kUnroll * CPL float4).
template <int CPL> struct Unroll { static constexpr int value = CPL <= 1 ? 8 : (CPL <= 2 ? 4 : 2); };

// One lane's share of a row-vs-query distance. Lane l owns 16-byte chunks l, l+32, ... of the
// row; inside a chunk x,y,z,w are accumulated in order with an UNFUSED multiply and add
// (__fmul_rn/__fadd_rn stop ptxas from contracting to FFMA), mirroring the reference's
// `diff*diff` then `sum +=` (hnsw.zig:188-189) element by element.
template <int METRIC>
__device__ __forceinline__ float accumulate_chunk(float acc, const float4 q, const float4 v) {
    if (METRIC == kMetricL2) {
        float d;
        d = __fsub_rn(q.x, v.x); acc = __fadd_rn(acc, __fmul_rn(d, d));
        d = __fsub_rn(q.y, v.y); acc = __fadd_rn(acc, __fmul_rn(d, d));
        d = __fsub_rn(q.z, v.z); acc = __fadd_rn(acc, __fmul_rn(d, d));
        d = __fsub_rn(q.w, v.w); acc = __fadd_rn(acc, __fmul_rn(d, d));
    } else {
        acc = __fadd_rn(acc, __fmul_rn(q.x, v.x));
        acc = __fadd_rn(acc, __fmul_rn(q.y, v.y));
        acc = __fadd_rn(acc, __fmul_rn(q.z, v.z));
        acc = __fadd_rn(acc, __fmul_rn(q.w, v.w));
    }
    return acc;
}

__device__ __forceinline__ float warp_butterfly_sum(float acc) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) acc = __fadd_rn(acc, __shfl_xor_sync(kFullMask, acc, off));
    return acc;
}

template <int METRIC>
__device__ __forceinline__ float finish_distance(float s) {
    if (METRIC == kMetricCos) return __fsub_rn(1.0f, s);
    if (METRIC == kMetricDot) return __fsub_rn(0.0f, s);
    return s;
}

// Distances of up to U rows (ids[0..nrows)) to the query held in qv, by one warp. All the row
// loads are issued before the first reduction so a warp keeps U*CPL 16-byte loads per lane in
// flight. Every lane returns the same bits.
template <int CPL, int METRIC, int U>
__device__ __forceinline__ void rows_distance(const float4 *__restrict__ arena, uint32_t row_chunks,
                                              const uint32_t (&ids)[U], int nrows, const float4 (&qv)[CPL],
                                              int lane, float (&out)[U]) {
    float4 v[U][CPL];
#pragma unroll
    for (int u = 0; u < U; ++u) {
#pragma unroll
        for (int c = 0; c < CPL; ++c) {
            const uint32_t chunk = lane + 32u * c;
            v[u][c] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (u < nrows && chunk < row_chunks)
                v[u][c] = __ldg(arena + static_cast<size_t>(ids[u]) * row_chunks + chunk);
        }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
        float acc = 0.0f;
#pragma unroll
        for (int c = 0; c < CPL; ++c) acc = accumulate_chunk<METRIC>(acc, qv[c], v[u][c]);
        out[u] = acc;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) out[u] = finish_distance<METRIC>(warp_butterfly_sum(out[u]));
}

// Exact visited set: open addressing, linear probing, multiplicative hash reduced with a
// multiply-high. Returns true if `id` was not present (and is now). Safe for concurrent lanes.
__device__ __forceinline__ bool visited_insert(uint32_t *table, uint32_t slots, uint32_t id) {
    uint32_t s = __umulhi(id * 0x9E3779B1u, slots);
    for (;;) {
        const uint32_t prev = atomicCAS(table + s, kInvalidId, id);
        if (prev == kInvalidId) return true;
        if (prev == id) return false;
        s = (s + 1 == slots) ? 0 : s + 1;
    }
}

// Number of keys in sorted a[0..n) that are < x.
__device__ __forceinline__ uint32_t lower_bound_u64(const uint64_t *a, uint32_t n, uint64_t x) {
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (a[mid] < x) lo = mid + 1; else hi = mid;
    }
    return lo;
}

template <int CPL, int METRIC>
__global__ void __launch_bounds__(256, (CPL <= 2 ? 4 : 2))
search_layer0_kernel(const SearchParams p) {
    constexpr int U = Unroll<CPL>::value;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t *keys = reinterpret_cast<uint64_t *>(smem_raw);   // [ef]: [0,np) popped in pop order, [np,np+nc) candidates ascending
    uint64_t *newk = keys + p.ef;                              // [32] keys pushed by the current pop
    uint32_t *table = reinterpret_cast<uint32_t *>(newk + 32); // [hash_words] visited set; scratch for the final sort
    uint32_t *todo = table + p.hash_words;                     // [32] unvisited neighbour ids of the current pop
    uint32_t *rank_ex = todo + 32;                             // [32] #existing candidates below each new key
    uint32_t *misc = rank_ex + 32;                             // [0] = #new, [1] = first candidate slot that moves

    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5, W = blockDim.x >> 5, T = blockDim.x;
    const uint32_t q = blockIdx.x;
    const float4 *__restrict__ arena = p.arena;

    // Query -> registers (every warp of the team holds the whole query, chunked like a row).
    float4 qv[CPL];
    {
        const float *qp = p.queries + static_cast<size_t>(q) * p.dim;
#pragma unroll
        for (int c = 0; c < CPL; ++c) {
            const uint32_t i = (lane + 32u * c) * 4u;
            qv[c].x = i + 0 < p.dim ? qp[i + 0] : 0.f;
            qv[c].y = i + 1 < p.dim ? qp[i + 1] : 0.f;
            qv[c].z = i + 2 < p.dim ? qp[i + 2] : 0.f;
            qv[c].w = i + 3 < p.dim ? qp[i + 3] : 0.f;
        }
    }
    for (uint32_t i = tid; i < p.slots; i += T) table[i] = kInvalidId;
    team_sync();

    uint32_t np = 0, nc = 1, nev = 1;   // pops done, candidates held, distance evaluations
    if (warp == 0) {                    // hnsw.zig:208-209: push the entry point, mark it visited
        uint32_t ids[U]; float d[U];
        ids[0] = p.entry;
        rows_distance<CPL, METRIC, U>(arena, p.row_chunks, ids, 1, qv, lane, d);
        if (lane == 0) { keys[0] = pack_key(d[0], p.entry); visited_insert(table, p.slots, p.entry); }
    }
    team_sync();

    while (nc > 0 && np < p.ef) {                              // hnsw.zig:211
        const uint32_t cur = key_id(keys[np]);                 // the minimum: candidates are sorted (:212)
        ++np; --nc;                                            // it is now result[np-1] in place (:214)
        const uint32_t cap = p.ef - np;                        // only this many more pops can ever happen
        uint64_t *cand = keys + np;
        for (uint32_t base = 0; base < p.m; base += 32) {      // hnsw.zig:216, 32 neighbours per pass
            if (warp == 0) {
                const uint32_t nb = (base + lane < p.m) ? __ldg(p.adj + static_cast<size_t>(cur) * p.m + base + lane)
                                                        : kInvalidId;
                const bool fresh = (nb != kInvalidId) && visited_insert(table, p.slots, nb);   // :217, :221
                const unsigned mask = __ballot_sync(kFullMask, fresh);
                if (fresh) todo[__popc(mask & ((1u << lane) - 1u))] = nb;   // adjacency order kept
                if (lane == 0) misc[0] = __popc(mask);
            }
            team_sync();
            const uint32_t t = misc[0];
            if (t == 0) { team_sync(); continue; }
            nev += t;

            // ---- distances (:219): warp w takes rows w, w+W, ... U at a time ----
            for (uint32_t j0 = warp * U; j0 < t; j0 += W * U) {
                uint32_t ids[U]; float d[U];
                const int nrows = min(static_cast<int>(t - j0), U);
#pragma unroll
                for (int u = 0; u < U; ++u) ids[u] = (u < nrows) ? todo[j0 + u] : 0u;
                rows_distance<CPL, METRIC, U>(arena, p.row_chunks, ids, nrows, qv, lane, d);
#pragma unroll
                for (int u = 0; u < U; ++u)
                    if (static_cast<int>(lane) == u && u < nrows) newk[j0 + u] = pack_key(d[u], ids[u]);
            }
            team_sync();

            // ---- push (:220) = merge the t new keys into the sorted candidates, keep the best `cap` ----
            uint32_t my_pos = kInvalidId; uint64_t my_key = 0;
            if (warp == 0) {
                uint32_t lo_mine = kInvalidId;
                if (lane < t) {
                    my_key = newk[lane];
                    const uint32_t re = lower_bound_u64(cand, nc, my_key);
                    uint32_t rn = 0;
                    for (uint32_t l = 0; l < t; ++l) rn += (newk[l] < my_key) ? 1u : 0u;
                    rank_ex[lane] = re;
                    if (re + rn < cap) { my_pos = re + rn; lo_mine = re; }
                }
#pragma unroll
                for (int off = 16; off >= 1; off >>= 1) lo_mine = min(lo_mine, __shfl_xor_sync(kFullMask, lo_mine, off));
                if (lane == 0) misc[1] = lo_mine;
            }
            team_sync();
            const uint32_t lo = misc[1];
            if (lo != kInvalidId) {
                // candidates [lo, nc) move right by the number of new keys below them, top chunk first
                for (uint32_t hi = nc; hi > lo;) {
                    const uint32_t span = min(T, hi - lo);
                    const bool active = tid < span;
                    const uint32_t i = hi - 1u - tid;
                    uint64_t e = 0; uint32_t s = 0;
                    if (active) {
                        e = cand[i];
                        for (uint32_t l = 0; l < t; ++l) s += (rank_ex[l] <= i) ? 1u : 0u;
                    }
                    team_sync();
                    if (active && i + s < cap) cand[i + s] = e;
                    hi -= span;
                }
                team_sync();
                if (my_pos != kInvalidId) cand[my_pos] = my_key;
            }
            nc = min(nc + t, cap);
            team_sync();
        }
    }

    // ---- result: stable sort of the popped entries by distance over pop order (hnsw.zig:227-233) ----
    uint64_t *sorted = reinterpret_cast<uint64_t *>(table);
    const uint32_t p2 = next_pow2(np);
    for (uint32_t i = tid; i < p2; i += T)
        sorted[i] = i < np ? ((keys[i] & 0xFFFFFFFF00000000ull) | i) : ~0ull;
    bitonic_sort_u64(sorted, p2);
    const uint32_t nres = min(np, p.k);
    for (uint32_t r = tid; r < p.k; r += T) {
        const size_t o = static_cast<size_t>(q) * p.k + r;
        if (r < nres) {
            const uint64_t key = keys[static_cast<uint32_t>(sorted[r])];
            p.ids[o] = static_cast<uint64_t>(key_id(key)) * p.id_stride + p.id_base;
            p.dist[o] = key_dist(key);
        } else {
            p.ids[o] = ~0ull;
            p.dist[o] = 0.0f;
        }
    }
    if (tid == 0) {
        p.counts[q] = nres;
        if (p.pops) p.pops[q] = np;
        if (p.evals) p.evals[q] = nev;
    }
}

}  // namespace zvdb
