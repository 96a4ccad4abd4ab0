// bruteforce.cuh -- K4: exact brute-force k-NN on the 5th-generation tensor cores (north_star
// subsystem 3). No reference counterpart: zvdb has no exact search.
// It provides ground truth for recall and a re-rank path, behind zvdb_bruteforce_knn*().
//
//   scores  S[q][r] = dot(Q[q], X[r])                    one GEMM, nq x n x dim
//   L2:     d = |x_r|^2 - 2 S   (|q|^2 is constant per query and dropped for ranking)
//   cosine: d = -S  (rows are normalised at insert);  dot: d = -S
//
// fp32-faithful products on TF32 tensor cores (3xTF32): every operand is split on the device into
// hi = x with the low 13 mantissa bits cleared and lo = (x - hi) with its low 13 bits cleared, both
// exactly representable in TF32, and  S = lo*hi + hi*lo + hi*hi  is accumulated in fp32 in TMEM
// (the dropped lo*lo term is below 2^-22 relative). Because the operands carry no bits the tensor
// core would discard, the result does not depend on whether the hardware truncates or rounds.
//
// Kernel structure (one persistent CTA per SM, 192 threads, sm_100a only):
//   warp 0   TMA producer : cp.async.bulk.tensor (128B swizzle) of Q_hi, Q_lo, X_hi, X_lo K-chunks
//                           into a multi-stage shared-memory ring, mbarrier complete_tx
//   warp 1   MMA issuer   : one thread issues tcgen05.mma.cta_group::1.kind::tf32, M=128 (queries)
//                           x N=128 (rows) x K=8, accumulators in TMEM, double buffered (2 x 128
//                           columns); tcgen05.commit releases smem stages / publishes accumulators
//   warps 2-5 epilogue    : tcgen05.ld 32x32b of the accumulator (thread t owns query t of the
//                           tile = TMEM lane t), distance formed in registers, threshold filter
//                           against the thread's running k-th best, rare insertion into a sorted
//                           per-thread list (shared memory; global memory when k is large).
//                           The score matrix never exists in memory.
// Work is cut into SEGMENTS (query tile, row-tile range, result slot) laid out by the host
// (plan_segments in capi.cu): whole waves of equal segments, one per CTA, sweeping the same rows at
// the same time (every X tile is then shared through L2 by all CTAs), and a last wave whose leftover
// segments are cut finer so that all CTAs finish together. Few, long segments matter: every segment
// starts with an empty candidate list and pays ~kp*ln(rows/kp) list insertions per query to warm
// its threshold up. Each segment leaves a sorted candidate list per query in its slot. A
// second small kernel merges the slots, RE-COMPUTES the distance of the best k+slack candidates
// exactly (difference form, the same summation order as the search kernel, so both kernels return
// bit-identical distances for the same (query, id)), orders by (distance, id) and writes k.
#pragma once
#include <cuda.h>
#include "search_kernel.cuh"

namespace zvdb {
namespace bf {

constexpr uint32_t kBM = 128;                        // queries per tile (UMMA M)
constexpr uint32_t kBN = 128;                        // rows per tile    (UMMA N)
constexpr uint32_t kBK = 32;                         // floats per K chunk = 128 bytes = one swizzle row
constexpr uint32_t kUmmaK = 8;                       // K of one tcgen05.mma.kind::tf32
constexpr uint32_t kTileBytes = kBM * kBK * 4;       // 16 KiB
constexpr uint32_t kStageBytes = 4 * kTileBytes;     // Q_hi, Q_lo, X_hi, X_lo
constexpr uint32_t kThreads = 192;
constexpr uint32_t kSlack = 8;                       // candidates kept beyond k for the exact re-rank (3xTF32)
constexpr uint32_t kSlackFilter = 24;                // ... when the GEMM is the single-product TF32 filter
constexpr uint32_t kMaxStages = 3;

struct BfParams {
    const float *xnorm;      // [n] squared row norms (L2 only)
    uint64_t *part_keys;     // [n_slots][nq][kp] sorted candidate keys per (slot, query); ~0 where a slot is unused
    uint64_t *glists;        // [gridDim.x][128][cap] per-thread lists when they do not fit in smem, else null
    const uint4 *segs;       // segments (query tile, first row tile, end row tile, slot), grouped by CTA
    const uint32_t *seg_off; // [gridDim.x + 1] CTA b owns segs[seg_off[b] .. seg_off[b+1])
    uint32_t n, nq, kchunks, kp;
    uint32_t n_slots, stages, metric;
    uint32_t cap;            // entries reserved per candidate list (= kp for EPI 0)
    uint32_t terms;          // 3 = 3xTF32 (fp32-faithful scores); 1 = hi*hi only: a FILTER, ~2^-11 relative score error
};

// ---- PTX wrappers ------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, uint64_t *bar, int32_t c0, int32_t c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// 32 consecutive accumulator columns of this thread's TMEM lane.
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Shared-memory matrix descriptor of a [128 rows][32 floats] K-major tile written by TMA with the
// 128-byte swizzle: rows are 128 bytes, 8-row groups are 1024 bytes apart (SBO), LBO is unused.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);     // start address, 16-byte units
    d |= static_cast<uint64_t>(1) << 16;                     // leading byte offset (ignored for swizzled K-major)
    d |= static_cast<uint64_t>(1024 >> 4) << 32;             // stride byte offset
    d |= static_cast<uint64_t>(1) << 46;                     // descriptor version of sm_100
    d |= static_cast<uint64_t>(2) << 61;                     // SWIZZLE_128B
    return d;
}
// Instruction descriptor (built in the kernel): D = f32 (1<<4), A = B = tf32 (2<<7, 2<<10), both
// K-major, N >> 3 at bit 17, M >> 4 at bit 24.

// Candidate lists of the 128 epilogue threads of a CTA: thread t (= one query) owns kp ascending keys,
// CONTIGUOUS ([128][kp]), so that a whole warp can work on one thread's list with one entry per lane
// and no bank conflicts. GL = false: in shared memory, addressed as such (ld/st.shared, not generic);
// GL = true: in global memory (k too large for shared memory; the lists stay L2-resident).
template <bool GL>
struct CandLists {
    uint64_t *g; uint32_t s; uint32_t kp; uint32_t cap;      // cap = entries reserved per thread (>= kp)
    __device__ __forceinline__ uint64_t get(uint32_t t, uint32_t i) const {
        if (GL) return g[static_cast<size_t>(t) * cap + i];
        uint64_t v; asm volatile("ld.shared.b64 %0, [%1];" : "=l"(v) : "r"(s + (t * cap + i) * 8u) : "memory"); return v;
    }
    __device__ __forceinline__ void put(uint32_t t, uint32_t i, uint64_t v) const {
        if (GL) { g[static_cast<size_t>(t) * cap + i] = v; return; }
        asm volatile("st.shared.b64 [%0], %1;" ::"r"(s + (t * cap + i) * 8u), "l"(v) : "memory");
    }
};

// Ascending bitonic sort of 32*E keys held E per lane: key i lives in lane i % 32, register i / 32.
template <int E>
__device__ __forceinline__ void warp_sort_keys(uint64_t (&v)[E], uint32_t lane) {
#pragma unroll
    for (uint32_t k = 2; k <= 32u * E; k <<= 1) {
#pragma unroll
        for (uint32_t j = k >> 1; j >= 1; j >>= 1) {
            if (j >= 32) {                                   // partner is another register of the same lane
                const uint32_t rj = j >> 5;
#pragma unroll
                for (uint32_t r = 0; r < static_cast<uint32_t>(E); ++r) {
                    if ((r & rj) == 0) {
                        const bool asc = ((r << 5) & k) == 0;
                        const uint64_t a = v[r], b = v[r | rj];
                        const uint64_t lo = min(a, b), hi = max(a, b);
                        v[r] = asc ? lo : hi; v[r | rj] = asc ? hi : lo;
                    }
                }
            } else {                                         // partner is lane ^ j, same register
#pragma unroll
                for (uint32_t r = 0; r < static_cast<uint32_t>(E); ++r) {
                    const uint64_t pv = shfl_xor_u64(v[r], static_cast<int>(j));
                    const bool asc = (((r << 5) | lane) & k) == 0;
                    const bool lower = (lane & j) == 0;
                    v[r] = (lower == asc) ? min(v[r], pv) : max(v[r], pv);
                }
            }
        }
    }
}

// Append-and-compact top-k (the epilogue's default): a thread APPENDS a passing candidate to its own
// unsorted list (one store, no cooperation); only when a list reaches its capacity does the warp sort
// it in registers (32*E keys, E per lane) and keep the kp best, which also tightens that thread's
// threshold. A compaction absorbs cap - kp appends, so the warm-up of a segment costs a few dozen
// warp-wide sorts per query instead of ~kp*ln(rows/kp) dependent list insertions.
// Sorts the first n_valid entries of thread t's list, leaves the kp smallest (ascending, ~0 padded) in
// entries [0, kp), returns entry kp-1. All 32 lanes call it with the same arguments.
template <bool GL, int E>
__device__ __forceinline__ uint64_t compact_list(const CandLists<GL> L, uint32_t t, uint32_t n_valid, uint32_t lane) {
    uint64_t v[E];
#pragma unroll
    for (int r = 0; r < E; ++r) {
        const uint32_t i = r * 32u + lane;
        v[r] = i < n_valid ? L.get(t, i) : ~0ull;
    }
    warp_sort_keys<E>(v, lane);
    __syncwarp();
    uint64_t last = ~0ull;
#pragma unroll
    for (int r = 0; r < E; ++r) {
        const uint32_t i = r * 32u + lane;
        if (i < L.kp) L.put(t, i, v[r]);
        const uint64_t cand = __shfl_sync(kFullMask, static_cast<uint32_t>(v[r] >> 32), (L.kp - 1) & 31);
        const uint64_t cand_lo = __shfl_sync(kFullMask, static_cast<uint32_t>(v[r]), (L.kp - 1) & 31);
        if (static_cast<uint32_t>(r) == ((L.kp - 1) >> 5)) last = (cand << 32) | cand_lo;
    }
    __syncwarp();
    return last;
}

// Sorted insertion of `key` into the list of thread `t`, done by the WHOLE WARP (one entry per lane):
// a thread inserting on its own walks its list serially while the other 31 lanes wait, and a segment
// start pays ~kp*ln(rows/kp) insertions per query. The caller guarantees key < the list's last entry
// (or the list is not full yet), so the position is at most kp-1 and the last entry falls out.
// Returns the list's new last key. All 32 lanes must call it with the same arguments.
template <bool GL>
__device__ __forceinline__ uint64_t coop_insert(const CandLists<GL> L, uint32_t t, uint64_t key, uint32_t lane) {
    const uint32_t kp = L.kp;
    uint32_t pos = 0;                                         // entries <= key (keys are unique: the id is in them)
    for (uint32_t c0 = 0; c0 < kp; c0 += 32) {
        const uint32_t i = c0 + lane;
        const uint64_t e = i < kp ? L.get(t, i) : ~0ull;
        const uint32_t cnt = __popc(__ballot_sync(kFullMask, e <= key));
        pos += cnt;
        if (cnt < 32) break;
    }
    const uint64_t below_last = kp >= 2 ? L.get(t, kp - 2) : 0ull;   // becomes the last entry unless the key does
    // entries [pos, kp-2] move up by one, highest 32 first, so nothing is overwritten before it is read
    for (uint32_t hi = kp - 1; hi > pos;) {
        const uint32_t lo = max(pos + 1, hi >= 31 ? hi - 31 : 0u);
        const uint32_t i = lo + lane;
        uint64_t v = 0;
        if (i <= hi) v = L.get(t, i - 1);
        __syncwarp();
        if (i <= hi) L.put(t, i, v);
        __syncwarp();
        hi = lo - 1;
    }
    if (lane == 0) L.put(t, pos, key);
    __syncwarp();
    return pos == kp - 1 ? key : below_last;
}

// ---- 2-CTA (cta_group::2) helpers ---------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA load whose completion bytes are counted on the LEADER CTA's mbarrier (bit 24 of a shared::cluster
// address is the CTA's rank inside the pair; clearing it names the same offset in CTA 0).
__device__ __forceinline__ void tma_load_2d_pair(void *dst, const CUtensorMap *map, uint64_t *bar, int32_t c0, int32_t c1) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1) : "memory");
}
// commit of all prior MMAs of the pair: one arrival on the barrier at this offset in BOTH CTAs
__device__ __forceinline__ void tc_commit_pair(uint64_t *bar) {
    const uint16_t mask = 3;
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrive on the barrier at this offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t *bar, uint32_t rank) {
    asm volatile("{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, %1;\n\tmbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
                 ::"r"(smem_u32(bar)), "r"(rank) : "memory");
}

// PAIR = false: one CTA per SM, UMMA 128 x 128 x 8 (cta_group::1).
// PAIR = true : the two CTAs of a cluster (the two SMs of a TPC) run ONE UMMA 256 x 256 x 8
//   (cta_group::2): CTA r holds queries [r*128, r*128+128) of a 256-query tile and rows [r*128, ..+128)
//   of a 256-row tile in its own shared memory, and receives the 128 x 256 accumulator of its queries
//   in its own TMEM. Per MMA cycle each SM then reads half the shared-memory bytes of the 1-CTA shape
//   (the 1-CTA kernel is bound by the 128 B/clk shared-memory port: 128 B/clk of operand reads plus
//   85 B/clk of TMA writes at full tensor rate). Only the leader (rank 0) issues MMAs; both CTAs run a
//   TMA producer (completion counted on the leader's barrier) and an epilogue; commits are multicast
//   to both CTAs' barriers; the peer's epilogue frees accumulators by arriving on the leader's barrier.
// EPI: 0 = sorted lists with warp-cooperative insertion (any k); E > 0 = append-and-compact with lists of
// up to 32*E entries per thread (kp < cap <= 32*E).
template <bool GL, bool PAIR, int EPI>
__global__ void __launch_bounds__(kThreads, 1)
bf_gemm_topk_kernel(const __grid_constant__ CUtensorMap tm_qhi, const __grid_constant__ CUtensorMap tm_qlo,
                    const __grid_constant__ CUtensorMap tm_xhi, const __grid_constant__ CUtensorMap tm_xlo,
                    const BfParams p) {
    constexpr uint32_t kTN = PAIR ? 2 * kBN : kBN;                  // rows per tile (UMMA N)
    constexpr uint32_t kCols = 2 * kTN;                             // TMEM columns: two accumulator buffers
    constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((kTN >> 3) << 17) | (((PAIR ? 2 * kBM : kBM) >> 4) << 24);
    extern __shared__ uint8_t smem_raw[];
    // 128B-swizzled TMA tiles want a 1024-byte aligned base
    uint8_t *tiles = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint64_t *bars = reinterpret_cast<uint64_t *>(tiles + static_cast<size_t>(p.stages) * kStageBytes);
    uint64_t *full = bars, *empty = bars + kMaxStages, *tfull = bars + 2 * kMaxStages, *tempty = tfull + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 12);
    float *xn_s = reinterpret_cast<float *>(bars + 16);                           // [2][kTN], 16-byte aligned (read as float4)
    uint64_t *lists_s = reinterpret_cast<uint64_t *>(xn_s + 2 * 2 * kBN);         // [128][cap] when in smem

    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = PAIR ? cluster_ctarank() : 0u;            // CTA within the pair
    const uint32_t unit = PAIR ? blockIdx.x >> 1 : blockIdx.x;      // scheduling unit: CTA or CTA pair
    if (threadIdx.x == 0) {
        for (uint32_t s = 0; s < p.stages; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        for (uint32_t b = 0; b < 2; ++b) { mbar_init(tfull + b, 1); mbar_init(tempty + b, PAIR ? 8 : 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        if (PAIR) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kCols) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kCols) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tc_fence_before();
    if (PAIR) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t seg_begin = __ldg(p.seg_off + unit), seg_end = __ldg(p.seg_off + unit + 1);

    if (warp == 0) {
        // ===== TMA producer (both CTAs of a pair: own queries, own half of the rows) =====
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (uint32_t si = seg_begin; si < seg_end; ++si) {
                const uint4 sg = __ldg(p.segs + si);
                const uint32_t qt = sg.x, t0 = sg.y, t1 = sg.z;
                const int32_t qrow = static_cast<int32_t>(PAIR ? (qt * 2 + rank) * kBM : qt * kBM);
                for (uint32_t t = t0; t < t1; ++t) {
                    const int32_t xrow = static_cast<int32_t>(t * kTN + rank * kBN);
                    for (uint32_t kc = 0; kc < p.kchunks; ++kc) {
                        mbar_wait(empty + stage, phase ^ 1);
                        uint8_t *st = tiles + static_cast<size_t>(stage) * kStageBytes;
                        const int32_t kx = static_cast<int32_t>(kc * kBK);
                        const bool lo = p.terms == 3;                                    // the lo parts feed only the two cross terms
                        const uint32_t bytes = lo ? kStageBytes : kStageBytes / 2;
                        if (PAIR) {
                            if (rank == 0) mbar_expect_tx(full + stage, 2 * bytes);      // both CTAs' bytes land on the leader's barrier
                            tma_load_2d_pair(st, &tm_qhi, full + stage, kx, qrow);
                            if (lo) tma_load_2d_pair(st + kTileBytes, &tm_qlo, full + stage, kx, qrow);
                            tma_load_2d_pair(st + 2 * kTileBytes, &tm_xhi, full + stage, kx, xrow);
                            if (lo) tma_load_2d_pair(st + 3 * kTileBytes, &tm_xlo, full + stage, kx, xrow);
                        } else {
                            mbar_expect_tx(full + stage, bytes);
                            tma_load_2d(st, &tm_qhi, full + stage, kx, qrow);
                            if (lo) tma_load_2d(st + kTileBytes, &tm_qlo, full + stage, kx, qrow);
                            tma_load_2d(st + 2 * kTileBytes, &tm_xhi, full + stage, kx, xrow);
                            if (lo) tma_load_2d(st + 3 * kTileBytes, &tm_xlo, full + stage, kx, xrow);
                        }
                        if (++stage == p.stages) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: one thread (of the leader CTA) =====
        if (lane == 0 && rank == 0) {
            uint32_t stage = 0, phase = 0, tile_count = 0;
            for (uint32_t si = seg_begin; si < seg_end; ++si) {
                const uint4 sg = __ldg(p.segs + si);
                const uint32_t t0 = sg.y, t1 = sg.z;
                for (uint32_t t = t0; t < t1; ++t, ++tile_count) {
                    const uint32_t buf = tile_count & 1, use = tile_count >> 1;
                    mbar_wait(tempty + buf, (use & 1) ^ 1);          // epilogue(s) drained this accumulator
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + buf * kTN;
                    for (uint32_t kc = 0; kc < p.kchunks; ++kc) {
                        mbar_wait(full + stage, phase);
                        tc_fence_after();
                        const uint32_t sa = smem_u32(tiles + static_cast<size_t>(stage) * kStageBytes);
                        const uint64_t q_hi = make_smem_desc(sa), q_lo = make_smem_desc(sa + kTileBytes);
                        const uint64_t x_hi = make_smem_desc(sa + 2 * kTileBytes), x_lo = make_smem_desc(sa + 3 * kTileBytes);
#pragma unroll
                        for (uint32_t ks = 0; ks < kBK / kUmmaK; ++ks) {
                            const uint64_t off = (ks * kUmmaK * 4) >> 4;     // 32 bytes per K step, in 16-byte units
                            const uint32_t first = (kc | ks) != 0;
                            if (PAIR) {
                                if (p.terms == 3) {
                                    tc_mma_tf32_pair(d_tmem, q_lo + off, x_hi + off, kIdesc, first);   // small terms first
                                    tc_mma_tf32_pair(d_tmem, q_hi + off, x_lo + off, kIdesc, 1);
                                }
                                tc_mma_tf32_pair(d_tmem, q_hi + off, x_hi + off, kIdesc, p.terms == 3 ? 1u : first);
                            } else {
                                if (p.terms == 3) {
                                    tc_mma_tf32(d_tmem, q_lo + off, x_hi + off, kIdesc, first);
                                    tc_mma_tf32(d_tmem, q_hi + off, x_lo + off, kIdesc, 1);
                                }
                                tc_mma_tf32(d_tmem, q_hi + off, x_hi + off, kIdesc, p.terms == 3 ? 1u : first);
                            }
                        }
                        if (PAIR) tc_commit_pair(empty + stage); else tc_commit(empty + stage);   // stage reusable once these MMAs retire
                        if (++stage == p.stages) { stage = 0; phase ^= 1; }
                    }
                    if (PAIR) tc_commit_pair(tfull + buf); else tc_commit(tfull + buf);           // accumulator complete
                }
            }
        }
    } else {
        // ===== epilogue: 4 warps, thread = one query of the CTA's 128 = one TMEM lane =====
        const uint32_t wq = warp & 3;                                // TMEM lane quarter this warp may read
        const uint32_t et = wq * 32 + lane;                          // query row in the tile
        CandLists<GL> L;                                             // this warp's 32 lists
        L.g = GL ? p.glists + (static_cast<size_t>(blockIdx.x) * 128 + wq * 32) * p.cap : nullptr;
        L.s = smem_u32(lists_s + static_cast<size_t>(wq) * 32 * p.cap);
        L.kp = p.kp; L.cap = p.cap;
        const float scale = p.metric == kMetricL2 ? -2.0f : -1.0f;
        const float inf = __int_as_float(0x7f800000);
        uint32_t tile_count = 0;
        for (uint32_t si = seg_begin; si < seg_end; ++si) {
            const uint4 sg = __ldg(p.segs + si);
            const uint32_t qt = sg.x, t0 = sg.y, t1 = sg.z, slot = sg.w;
            if (EPI == 0) { for (uint32_t i = 0; i < p.kp; ++i) L.put(lane, i, ~0ull); }
            __syncwarp();
            float tau = inf;
            uint32_t cnt = 0;                                        // EPI > 0: entries in this thread's list
            for (uint32_t t = t0; t < t1; ++t, ++tile_count) {
                const uint32_t buf = tile_count & 1, use = tile_count >> 1;
                const uint32_t row0 = t * kTN;
#pragma unroll
                for (uint32_t h = 0; h < kTN / kBN; ++h) {
                    const uint32_t r = row0 + h * kBN + et;
                    xn_s[buf * kTN + h * kBN + et] = r < p.n ? (p.metric == kMetricL2 ? __ldg(p.xnorm + r) : 0.0f) : inf;
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");       // the 128 epilogue threads only
                mbar_wait(tfull + buf, use & 1);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((wq * 32u) << 16) + buf * kTN;
#pragma unroll 1
                for (uint32_t c = 0; c < kTN / 32; ++c) {
                    uint32_t acc[32];
                    tc_ld32(taddr + c * 32, acc);
                    const float4 *xn4 = reinterpret_cast<const float4 *>(xn_s + buf * kTN + c * 32);
                    float d[32];
                    float mn = inf;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 x = xn4[j];
                        d[4 * j + 0] = fmaf(scale, __uint_as_float(acc[4 * j + 0]), x.x);
                        d[4 * j + 1] = fmaf(scale, __uint_as_float(acc[4 * j + 1]), x.y);
                        d[4 * j + 2] = fmaf(scale, __uint_as_float(acc[4 * j + 2]), x.z);
                        d[4 * j + 3] = fmaf(scale, __uint_as_float(acc[4 * j + 3]), x.w);
                        mn = fminf(mn, fminf(fminf(d[4 * j + 0], d[4 * j + 1]), fminf(d[4 * j + 2], d[4 * j + 3])));
                    }
                    if (__any_sync(kFullMask, mn < tau)) {                      // rare once tau is warm
                        uint32_t pm = 0;                                         // this thread's columns that pass
#pragma unroll
                        for (int j = 0; j < 32; ++j) pm |= d[j] < tau ? 1u << j : 0u;
                        // Columns are taken in ascending order per thread (rows arrive in ascending id order, so on
                        // an exact tie the earlier id stays: strict <); every round each thread offers its lowest
                        // remaining column and the warp inserts the offers one list at a time.
                        while (__any_sync(kFullMask, pm != 0)) {
                            const int jj = pm ? __ffs(pm) - 1 : -1;
                            pm &= pm - 1;
                            float dsel = inf;
#pragma unroll
                            for (int j = 0; j < 32; ++j) dsel = j == jj ? d[j] : dsel;
                            if constexpr (EPI == 0) {
                                unsigned m = __ballot_sync(kFullMask, dsel < tau);   // tau may have tightened since pm was built
                                while (m) {
                                    const uint32_t owner = __ffs(m) - 1;             // the lane (= query) whose candidate this is
                                    m &= m - 1;
                                    const float dk = __shfl_sync(kFullMask, dsel, owner);
                                    const uint32_t col = __shfl_sync(kFullMask, jj, owner);
                                    const uint64_t last = coop_insert<GL>(L, owner, pack_key(dk, row0 + c * 32 + col), lane);
                                    if (lane == owner) tau = last == ~0ull ? inf : key_dist(last);
                                }
                            } else {
                                if (dsel < tau) { L.put(lane, cnt, pack_key(dsel, row0 + c * 32 + jj)); ++cnt; }
                                unsigned full = __ballot_sync(kFullMask, cnt == p.cap);
                                while (full) {                                       // rare: once per cap - kp appends of a thread
                                    const uint32_t owner = __ffs(full) - 1;
                                    full &= full - 1;
                                    __syncwarp();
                                    const uint64_t last = compact_list<GL, EPI>(L, owner, p.cap, lane);
                                    if (lane == owner) { cnt = p.kp; tau = key_dist(last); }
                                }
                            }
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if (PAIR) mbar_arrive_remote(tempty + buf, 0); else mbar_arrive(tempty + buf);
                }
            }
            if constexpr (EPI > 0) {                                 // leave every list sorted, kp long, ~0 padded
                __syncwarp();
                for (uint32_t owner = 0; owner < 32; ++owner)
                    compact_list<GL, EPI>(L, owner, __shfl_sync(kFullMask, cnt, owner), lane);
            }
            const uint32_t q = (PAIR ? qt * 2 + rank : qt) * kBM + et;
            if (q < p.nq) {
                uint64_t *out = p.part_keys + (static_cast<size_t>(slot) * p.nq + q) * p.kp;
                for (uint32_t i = 0; i < p.kp; ++i) out[i] = L.get(lane, i);
            }
        }
    }

    tc_fence_before();
    if (PAIR) cluster_sync_all(); else __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kCols) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kCols) : "memory");
    }
}

// ---- operand split ----------------------------------------------------------------------------
// One warp per row: hi/lo TF32 parts into [rows][dst_pitch] (zero padded beyond `cols`), and the
// squared norm of the row (fp32, lane order + butterfly) when `norm` is given.
__global__ void split_tf32_kernel(const float *__restrict__ src, uint32_t src_pitch, uint32_t cols, uint64_t rows,
                                  float *__restrict__ hi, float *__restrict__ lo, uint32_t dst_pitch, float *__restrict__ norm) {
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t warps = (static_cast<uint64_t>(gridDim.x) * blockDim.x) >> 5;
    for (uint64_t r = (static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; r < rows; r += warps) {
        float acc = 0.0f;
        for (uint32_t c = lane; c < dst_pitch; c += 32) {
            const float x = c < cols ? src[r * src_pitch + c] : 0.0f;
            const float h = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
            const float l = __uint_as_float(__float_as_uint(__fsub_rn(x, h)) & 0xFFFFE000u);
            hi[r * dst_pitch + c] = h;
            lo[r * dst_pitch + c] = l;
            acc = fmaf(x, x, acc);
        }
        if (norm) {
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) acc += __shfl_xor_sync(kFullMask, acc, off);
            if (lane == 0) norm[r] = acc;
        }
    }
}

// ---- merge of the splits + exact re-rank ---------------------------------------------------------
struct BfFinalParams {
    const float4 *arena;
    const float *queries;      // [nq][dim]
    const uint64_t *part_keys; // [n_splits][nq][kp]  (n_splits = result slots per query; unused ones hold ~0)
    uint64_t *ids;             // [nq][k]
    float *dist;               // [nq][k]
    uint32_t *counts;          // [nq]
    uint64_t id_stride, id_base;
    uint32_t row_chunks, dim, nq, k, kp, n_splits, p2, kk2;   // p2 = pow2 >= n_splits*kp ; kk2 = pow2 >= kp (= k + slack)
};

template <int CPL, int METRIC>
__global__ void __launch_bounds__(32) bf_finalize_kernel(const BfFinalParams p) {
    constexpr int U = Unroll<CPL, false>::value;
    constexpr uint32_t LPR = 32 / U;
    extern __shared__ __align__(16) unsigned char smem_fin[];
    uint64_t *keys = reinterpret_cast<uint64_t *>(smem_fin);      // [p2]
    uint64_t *exact = keys + p.p2;                                // [kk2]
    const uint32_t q = blockIdx.x, lane = threadIdx.x;
    const uint32_t total = p.n_splits * p.kp;
    for (uint32_t i = lane; i < p.p2; i += 32) {
        uint64_t key = ~0ull;
        if (i < total) key = p.part_keys[(static_cast<size_t>(i / p.kp) * p.nq + q) * p.kp + (i % p.kp)];
        keys[i] = key;
    }
    bitonic_sort_u64(keys, p.p2);
    const uint32_t want = min(p.kp, total);
    Chunk2 qv[CPL];
    {
        const float *qp = p.queries + static_cast<size_t>(q) * p.dim;
#pragma unroll
        for (int c = 0; c < CPL; ++c) {
            const uint32_t i = (lane + 32u * c) * 4u;
            const float x = i + 0 < p.dim ? qp[i + 0] : 0.f, y = i + 1 < p.dim ? qp[i + 1] : 0.f;
            const float z = i + 2 < p.dim ? qp[i + 2] : 0.f, w = i + 3 < p.dim ? qp[i + 3] : 0.f;
            qv[c].xy = pack2(x, y); qv[c].zw = pack2(z, w);
        }
    }
    for (uint32_t i = lane; i < p.kk2; i += 32) exact[i] = ~0ull;
    __syncwarp();
    uint32_t nvalid = 0;
    for (uint32_t j0 = 0; j0 < want; j0 += U) {
        uint32_t ids[U];
        bool ok[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint64_t key = j0 + u < want ? keys[j0 + u] : ~0ull;
            ok[u] = key != ~0ull;
            ids[u] = ok[u] ? key_id(key) : 0u;                   // row 0 always exists when any key is valid
        }
        const float d = rows_distance<CPL, METRIC, U>(p.arena, p.row_chunks, ids, qv, lane);
        const uint32_t r = lane / LPR;
        bool mine = false;
#pragma unroll
        for (int u = 0; u < U; ++u) if (r == static_cast<uint32_t>(u)) mine = ok[u];
        if ((lane % LPR) == 0 && mine) {
            uint32_t id = 0;
#pragma unroll
            for (int u = 0; u < U; ++u) if (r == static_cast<uint32_t>(u)) id = ids[u];
            exact[j0 + r] = pack_key(d, id);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) nvalid += ok[u] ? 1u : 0u;
    }
    __syncwarp();
    bitonic_sort_u64(exact, p.kk2);
    const uint32_t nres = min(nvalid, p.k);
    for (uint32_t r = lane; r < p.k; r += 32) {
        const size_t o = static_cast<size_t>(q) * p.k + r;
        if (r < nres) {
            p.ids[o] = static_cast<uint64_t>(key_id(exact[r])) * p.id_stride + p.id_base;
            p.dist[o] = key_dist(exact[r]);
        } else {
            p.ids[o] = ~0ull;
            p.dist[o] = 0.0f;
        }
    }
    if (lane == 0) p.counts[q] = nres;
}

}  // namespace bf
}  // namespace zvdb
