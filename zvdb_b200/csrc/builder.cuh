// builder.cuh -- K6: a better graph producer in the reference's layout.
//
// The reference's insert links one neighbour per layer (hnsw.zig:106-108), which leaves most nodes
// unreachable from node 0. This builder keeps the layout the search kernel and the
// reference share -- at most m layer-0 neighbours per node, entry point = node 0 -- but chooses
// the neighbours from a candidate list (e.g. approximate k-NN ids), so the reference's own search
// loop reaches a useful recall. It is an extension, not a restatement: parity for it is checked
// against tests/builder_ref.py, a CPU statement of exactly this algorithm.
//
// Algorithm (every distance in the kernel's own summation order, ties by id):
//   1. forward:  C_i = unique valid candidates of i, sorted by (d(i,c), c).
//                Walk C_i; keep c unless an already kept s has d(c,s) < d(i,c)  (relative-
//                neighbourhood rule); stop at m.                                      -> fw[i]
//   2. reverse:  rev[j] = { i : j in fw[i] }                                          (CSR)
//   3. final:    U_i = unique(fw[i] + rev[i]) sorted by (d(i,x), x), cut to the nearest kUnionCap;
//                apply the same rule (stop at m); if fewer than m survive, top up with the
//                nearest rejected ones.                                                -> adj[i]
#pragma once
#include "search_kernel.cuh"

namespace zvdb {

constexpr uint32_t kBuildBuf = 128;    // sorted work buffer (keys) per node
constexpr uint32_t kUnionCap = 96;     // union members kept (nearest first) before the final rule

struct BuildParams {
    const float4 *arena;
    uint32_t row_chunks;
    uint32_t n, m;
    const uint32_t *cand;     // [n][K]
    uint32_t K;
    uint32_t *fw;             // [n][m]  kInvalidId padded
    const uint64_t *rev_off;  // [n+1]
    const uint32_t *rev;      // [rev_off[n]]
    uint32_t *adj;            // [n][m]  final table
};

template <int CPL>
__device__ __forceinline__ void load_row(const float4 *__restrict__ arena, uint32_t row_chunks, uint32_t id, uint32_t lane,
                                         Chunk2 (&v)[CPL]) {
#pragma unroll
    for (int c = 0; c < CPL; ++c) {
        const uint32_t chunk = lane + 32u * c;
        const float4 f = chunk < row_chunks ? __ldg(arena + static_cast<size_t>(id) * row_chunks + chunk) : make_float4(0.f, 0.f, 0.f, 0.f);
        v[c].xy = pack2(f.x, f.y); v[c].zw = pack2(f.z, f.w);
    }
}

// keys[0..cnt) hold pack_key(d(i,x), x) for ids[0..cnt) (warp-cooperative, U rows at a time).
template <int CPL, int METRIC>
__device__ __forceinline__ void fill_keys(const float4 *arena, uint32_t row_chunks, const Chunk2 (&qv)[CPL], uint32_t self,
                                          const uint32_t *ids_smem, uint32_t cnt, uint64_t *keys, uint32_t lane, uint32_t n) {
    constexpr int U = Unroll<CPL, false>::value;
    constexpr uint32_t LPR = 32 / U;
    for (uint32_t j0 = 0; j0 < cnt; j0 += U) {
        uint32_t ids[U];
        const int nrows = min(static_cast<int>(cnt - j0), U);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            ids[u] = (u < nrows) ? ids_smem[j0 + u] : 0u;
            if (ids[u] >= n) ids[u] = self;            // invalid / padding: harmless row, key voided below
        }
        const float d = rows_distance<CPL, METRIC, U>(arena, row_chunks, ids, qv, lane);
        const uint32_t r = lane / LPR;
        if ((lane % LPR) == 0 && r < static_cast<uint32_t>(nrows)) {
            const uint32_t raw = ids_smem[j0 + r];
            keys[j0 + r] = (raw >= n || raw == self) ? ~0ull : pack_key(d, raw);
        }
    }
}

// Relative-neighbourhood selection over sorted unique keys[0..cnt): writes up to m ids to out[],
// returns the count. `taken` (bit per key, in shared memory) marks the selected positions.
template <int CPL, int METRIC>
__device__ __forceinline__ uint32_t rng_select(const float4 *arena, uint32_t row_chunks, const uint64_t *keys, uint32_t cnt,
                                               uint32_t m, uint32_t *out, uint32_t *taken, uint32_t lane) {
    constexpr int U = Unroll<CPL, false>::value;
    constexpr uint32_t LPR = 32 / U;
    uint32_t nsel = 0;
    for (uint32_t idx = 0; idx < cnt && nsel < m; ++idx) {
        const uint64_t key = keys[idx];
        if (key == ~0ull) break;                                    // sorted: the rest is void
        if (idx > 0 && key_id(keys[idx - 1]) == key_id(key)) continue;   // duplicate id (same distance, adjacent)
        const uint32_t c = key_id(key);
        const uint32_t d_ic = static_cast<uint32_t>(key >> 32);     // ordered bits: integer compare == float compare
        Chunk2 cv[CPL];
        load_row<CPL>(arena, row_chunks, c, lane, cv);
        bool ok = true;
        for (uint32_t s0 = 0; s0 < nsel && ok; s0 += U) {
            uint32_t ids[U];
            const int nrows = min(static_cast<int>(nsel - s0), U);
#pragma unroll
            for (int u = 0; u < U; ++u) ids[u] = (u < nrows) ? out[s0 + u] : 0u;
            const float d = rows_distance<CPL, METRIC, U>(arena, row_chunks, ids, cv, lane);
            const bool closer = (lane / LPR) < static_cast<uint32_t>(nrows) && float_to_ordered(d) < d_ic;
            if (__any_sync(kFullMask, closer)) ok = false;
        }
        if (ok) {
            if (lane == 0) { out[nsel] = c; taken[idx >> 5] |= 1u << (idx & 31); }
            ++nsel;
            __syncwarp();
        }
    }
    return nsel;
}

// Stage 1. One warp (= one CTA) per node.
template <int CPL, int METRIC>
__global__ void __launch_bounds__(32) build_forward_kernel(const BuildParams p) {
    __shared__ uint64_t keys[kBuildBuf];
    __shared__ uint32_t ids[kBuildBuf];
    __shared__ uint32_t sel[64];
    __shared__ uint32_t taken[kBuildBuf / 32];
    const uint32_t i = blockIdx.x;
    const uint32_t lane = threadIdx.x;
    Chunk2 qv[CPL];
    load_row<CPL>(p.arena, p.row_chunks, i, lane, qv);
    for (uint32_t t = lane; t < kBuildBuf; t += 32) { keys[t] = ~0ull; ids[t] = t < p.K ? p.cand[static_cast<size_t>(i) * p.K + t] : kInvalidId; }
    if (lane < kBuildBuf / 32) taken[lane] = 0;
    __syncwarp();
    fill_keys<CPL, METRIC>(p.arena, p.row_chunks, qv, i, ids, p.K, keys, lane, p.n);
    bitonic_sort_u64(keys, kBuildBuf);
    const uint32_t nsel = rng_select<CPL, METRIC>(p.arena, p.row_chunks, keys, p.K, p.m, sel, taken, lane);
    __syncwarp();
    for (uint32_t t = lane; t < p.m; t += 32) p.fw[static_cast<size_t>(i) * p.m + t] = t < nsel ? sel[t] : kInvalidId;
}

// Stage 2a/2b: in-degree histogram, then fill of the reverse CSR rows.
__global__ void count_reverse_kernel(const uint32_t *__restrict__ fw, uint64_t total, uint32_t *__restrict__ indeg) {
    for (uint64_t e = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; e < total; e += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
        const uint32_t j = fw[e];
        if (j != kInvalidId) atomicAdd(indeg + j, 1u);
    }
}
__global__ void fill_reverse_kernel(const uint32_t *__restrict__ fw, uint64_t total, uint32_t m, const uint64_t *__restrict__ off,
                                    uint32_t *__restrict__ cursor, uint32_t *__restrict__ rev) {
    for (uint64_t e = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; e < total; e += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
        const uint32_t j = fw[e];
        if (j != kInvalidId) rev[off[j] + atomicAdd(cursor + j, 1u)] = static_cast<uint32_t>(e / m);
    }
}

// Stage 3. One warp per node. The union is streamed through the sorted buffer 32 ids at a time:
// the nearest kUnionCap survive each round, so any in-degree is handled with bounded memory and
// the result does not depend on the (atomic) order of rev[].
template <int CPL, int METRIC>
__global__ void __launch_bounds__(32) build_final_kernel(const BuildParams p) {
    __shared__ uint64_t keys[kBuildBuf];
    __shared__ uint32_t ids[32];
    __shared__ uint32_t sel[64];
    __shared__ uint32_t taken[kBuildBuf / 32];
    const uint32_t i = blockIdx.x;
    const uint32_t lane = threadIdx.x;
    Chunk2 qv[CPL];
    load_row<CPL>(p.arena, p.row_chunks, i, lane, qv);
    for (uint32_t t = lane; t < kBuildBuf; t += 32) keys[t] = ~0ull;
    if (lane < kBuildBuf / 32) taken[lane] = 0;
    const uint64_t rb = p.rev_off[i], re = p.rev_off[i + 1];
    const uint64_t total = p.m + (re - rb);                 // forward slots first, then the reverse row
    __syncwarp();
    for (uint64_t base = 0; base < total; base += 32) {
        const uint64_t e = base + lane;
        uint32_t x = kInvalidId;
        if (e < p.m) x = p.fw[static_cast<size_t>(i) * p.m + e];
        else if (e < total) x = p.rev[rb + (e - p.m)];
        ids[lane] = x;
        __syncwarp();
        // new keys go to the tail [kUnionCap, kUnionCap+32) of the buffer, then the whole buffer is re-sorted
        fill_keys<CPL, METRIC>(p.arena, p.row_chunks, qv, i, ids, 32, keys + kUnionCap, lane, p.n);
        __syncwarp();
        bitonic_sort_u64(keys, kBuildBuf);
        // drop duplicates that crossed rounds (a node can be both a forward and a reverse neighbour)
        uint64_t mine[kBuildBuf / 32];
#pragma unroll
        for (uint32_t r = 0; r < kBuildBuf / 32; ++r) {
            const uint32_t t = lane + 32 * r;
            uint64_t k = keys[t];
            if (t > 0 && k != ~0ull && keys[t - 1] == k) k = ~0ull;
            mine[r] = k;
        }
        __syncwarp();
#pragma unroll
        for (uint32_t r = 0; r < kBuildBuf / 32; ++r) keys[lane + 32 * r] = mine[r];
        __syncwarp();
        bitonic_sort_u64(keys, kBuildBuf);
        for (uint32_t t = kUnionCap + lane; t < kBuildBuf; t += 32) keys[t] = ~0ull;   // keep the nearest kUnionCap
        __syncwarp();
    }
    uint32_t nsel = rng_select<CPL, METRIC>(p.arena, p.row_chunks, keys, kUnionCap, p.m, sel, taken, lane);
    __syncwarp();
    if (lane == 0) {   // top up with the nearest rejected members
        for (uint32_t idx = 0; idx < kUnionCap && nsel < p.m; ++idx) {
            if (keys[idx] == ~0ull) break;
            if (!(taken[idx >> 5] >> (idx & 31) & 1u)) sel[nsel++] = key_id(keys[idx]);
        }
        ids[0] = nsel;
    }
    __syncwarp();
    nsel = ids[0];
    for (uint32_t t = lane; t < p.m; t += 32) p.adj[static_cast<size_t>(i) * p.m + t] = t < nsel ? sel[t] : kInvalidId;
}

}  // namespace zvdb
