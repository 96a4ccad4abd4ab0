"""GPU parity tests for K1L, the latency form of the layer-0 search (one CTA of 8 warps per query,
zvdb_b200/csrc/search_team_kernel.cuh), called through the C ABI against the CPU oracle.

K1L is what a small batch and the reference's own call pattern -- one search(query, k) at a time
(benchmarks/shared_benchmarks.zig:104-109) -- run on; it must return what the one-warp kernel and the oracle
return, bit for bit: ids, order, distance bits, result counts and the pop / evaluation counters
(src/hnsw.zig:194-236). zvdb_set_kernel_variant bits 14-15 force it off (1), on with teams of 8 warps (2) or on with
teams of 4 warps (3: what batches too large for full teams to be resident run on), so every form is held to the oracle
on the same shapes, whatever the automatic batch-size rule picks.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RTOL = 1e-5
TEAM_NEVER, TEAM_ALWAYS, TEAM_HALF = 1 << 14, 2 << 14, 3 << 14   # one warp per query / teams of 8 warps / teams of 4 warps


def _gauss(n, dim, seed):
    return np.random.default_rng(seed).standard_normal((n, dim), dtype=np.float32)


def _bit_exact(zv, got, ref, k):
    ids, dist, counts, pops, evals = got
    assert np.array_equal(counts, ref["counts"])
    assert np.array_equal(pops, ref["pops"])
    assert np.array_equal(evals, ref["evals"])
    mask = np.arange(k)[None, :] < counts[:, None]
    assert np.array_equal(ids[mask], ref["ids"].astype(np.uint64)[mask])
    assert np.array_equal(dist.view(np.uint32)[mask], ref["dist"].view(np.uint32)[mask])
    assert np.all(ids[~mask] == zv.INVALID_ID)


@pytest.mark.parametrize("team", [TEAM_NEVER, TEAM_ALWAYS, TEAM_HALF])
@pytest.mark.parametrize("n,dim,m,k,ef", [
    (10000, 128, 16, 10, 10),     # the reference call, ef = k
    (10000, 128, 16, 10, 64),     # C5's operating point
    (10000, 128, 16, 10, 512),    # 8 193 candidate slots, a 512-entry final sort
    (10000, 128, 16, 300, 600),   # long result list
    (4000, 3, 16, 5, 20),         # tiny dim: most lanes hold padding
    (4000, 200, 16, 10, 40),      # dim not a multiple of 128 floats (predicated chunk loads)
    (3000, 768, 32, 100, 128),    # C3 shape: 6 chunks per virtual lane, two 16-neighbour passes per pop
    (3000, 1024, 8, 10, 16),      # 8 chunks per virtual lane, half the half-warps idle
    (2000, 64, 40, 10, 30),       # m = 40: three passes, the last one 8 wide
    (300, 16, 4, 5, 300),         # ef = n: the search runs dry
    (2000, 64, 40, 5, 12),        # m = 40 with the per-slot adjacency cache on chip (three passes, cp.async of 40-word rows)
    (3000, 96, 32, 10, 20),       # m = 32 with the cache (two full passes)
])
def test_both_kernels_bit_exact_vs_oracle(zv, oracle, team, n, dim, m, k, ef):
    X = _gauss(n, dim, 241)
    h = zv.HNSW(m, 200)
    h.insert_batch(X)
    adj, _ = h.export_layer(0)
    h.set_kernel_variant(team)
    Q = _gauss(97, dim, 242)
    got = h.search_batch(Q, k, ef, counters=True)
    ref = oracle.search_graph(X, adj, Q, ef, k, dist_mode=oracle.DIST_TREE, heap_mode=oracle.HEAP_DET)
    _bit_exact(zv, got, ref, k)
    # ... and against the reference's own arithmetic and heap (sequential sum, Zig PriorityQueue): distances within 1e-5
    # relative position by position. Where two candidates tie within that tolerance the two arithmetics may pop them in
    # the other order, and what the later pops then expand can differ: such queries must be rare.
    seq = oracle.search_graph(X, adj, Q, ef, k, dist_mode=oracle.DIST_SEQ, heap_mode=oracle.HEAP_ZIG)
    assert np.array_equal(got[2], seq["counts"])
    mask = np.arange(k)[None, :] < got[2][:, None]
    close = np.isclose(got[1], seq["dist"], rtol=RTOL, atol=1e-30) | ~mask
    assert (~close.all(axis=1)).sum() <= max(1, 0.05 * len(Q)), f"{(~close.all(axis=1)).sum()} of {len(Q)} queries differ beyond near-ties"
    h.deinit()


@pytest.mark.parametrize("name,n,dim,m,k,ef", [
    ("cos", 3000, 768, 32, 100, 128),    # C3 shape
    ("cos", 5000, 32, 16, 10, 64),
    ("dot", 5000, 128, 16, 10, 64),
    ("dot", 4000, 200, 8, 5, 100),
])
def test_team_kernel_under_metric_vs_oracle(zv, oracle, name, n, dim, m, k, ef):
    zm, om = {"cos": (zv.METRIC_COSINE, oracle.METRIC_COS), "dot": (zv.METRIC_DOT, oracle.METRIC_DOT)}[name]
    h = zv.HNSW(m, 200, metric=zm)
    h.insert_batch(_gauss(n, dim, 243))
    Xs = np.stack([h.point(i) for i in range(h.count())]).astype(np.float32)   # rows as stored (normalised for cosine)
    adj, _ = h.export_layer(0)
    h.set_kernel_variant(TEAM_ALWAYS)
    Q = _gauss(65, dim, 244)
    got = h.search_batch(Q, k, ef, counters=True)
    ref = oracle.search_graph(Xs, adj, Q, ef, k, dist_mode=oracle.DIST_TREE | om, heap_mode=oracle.HEAP_DET)
    _bit_exact(zv, got, ref, k)
    h.deinit()


def test_team_kernel_on_a_quality_graph(zv, oracle):
    """Full-degree nodes (16 fresh neighbours per pop, every half-warp busy) on a builder graph."""
    from zvdb_b200 import builder
    n, dim, m = 20000, 128, 16
    X = _gauss(n, dim, 245)
    h = zv.HNSW(m, 200)
    builder.build_quality_graph(h, X, m, K=48)
    adj, _ = h.export_layer(0)
    Q = _gauss(120, dim, 246)
    for team in (TEAM_NEVER, TEAM_ALWAYS, TEAM_HALF):
        h.set_kernel_variant(team)
        for k, ef in ((10, 10), (10, 64), (10, 200)):
            got = h.search_batch(Q, k, ef, counters=True)
            ref = oracle.search_graph(X, adj, Q, ef, k, dist_mode=oracle.DIST_TREE, heap_mode=oracle.HEAP_DET)
            _bit_exact(zv, got, ref, k)
    h.deinit()


def test_automatic_choice_single_call_and_batch_sizes_agree(zv, oracle):
    """search(query, k) -- the reference's call -- and batches on either side of the automatic threshold return the
    same rows whichever kernel serves them; with the prefetch variants too."""
    n, dim, m = 8000, 128, 16
    X = _gauss(n, dim, 247)
    h = zv.HNSW(m, 200)
    h.insert_batch(X)
    adj, _ = h.export_layer(0)
    Q = _gauss(700, dim, 248)
    ref = oracle.search_graph(X, adj, Q, 48, 10, dist_mode=oracle.DIST_TREE, heap_mode=oracle.HEAP_DET)
    full = h.search_batch(Q, 10, 48, counters=True)                  # 700 queries: the one-warp kernel
    _bit_exact(zv, full, ref, 10)
    for variant in (0, TEAM_NEVER, TEAM_ALWAYS, TEAM_HALF, TEAM_ALWAYS | 0x100, TEAM_ALWAYS | 0x400):   # automatic, off, teams of 8 / 4 warps, forced next to K1-only bits
        h.set_kernel_variant(variant)
        for s, bs in ((0, 1), (1, 7), (8, 64), (100, 296), (50, 600), (0, 700)):
            got = h.search_batch(Q[s:s + bs], 10, 48, counters=True)
            for a, b in zip(got, full):
                assert np.array_equal(a.view(np.uint8), b[s:s + bs].view(np.uint8)), (variant, s, bs)
    h.set_kernel_variant(0)
    one = h.search(Q[5], 10)                                         # search(q, k) == search_batch(ef = k)
    r1 = oracle.search_graph(X, adj, Q[5:6], 10, 10, dist_mode=oracle.DIST_TREE, heap_mode=oracle.HEAP_DET)
    assert [r.id for r in one] == [int(x) for x in r1["ids"][0, :r1["counts"][0]]]
    assert [np.float32(r.distance).view(np.uint32) for r in one] == [x for x in r1["dist"][0, :r1["counts"][0]].view(np.uint32)]
    h.deinit()


def test_team_kernel_after_the_descent(zv, oracle):
    """K2 seeds the team kernel like it seeds the one-warp kernel (reference-built hierarchy, hnsw.zig:88-108)."""
    rng = np.random.default_rng(249)
    n, dim, m = 6000, 32, 8
    X = rng.standard_normal((n, dim), dtype=np.float32)
    Q = rng.standard_normal((90, dim), dtype=np.float32)
    lv = np.minimum(rng.geometric(0.5, n) - 1, 31).astype(np.int32)
    h = zv.HNSW(m, 200)
    h.insert_batch(X, levels=lv)
    adj, _ = h.export_layer(0)
    lvl, ub, ua = h.export_upper_layers()
    up = (lvl, ub, ua, h.max_level, h.descent_start)
    h.set_descent(True)
    for team in (TEAM_NEVER, TEAM_ALWAYS, TEAM_HALF):
        h.set_kernel_variant(team)
        for k, ef in ((10, 10), (10, 64)):
            got = h.search_batch(Q, k, ef, counters=True)
            ref = oracle.search_graph(X, adj, Q, ef, k, dist_mode=oracle.DIST_TREE, heap_mode=oracle.HEAP_DET, upper=up)
            _bit_exact(zv, got, ref, k)
    h.deinit()


def test_team_kernel_too_large_for_shared_memory_falls_back(zv, oracle):
    """ef * m beyond what one CTA's shared memory holds: 'whenever the shape fits' means the one-warp kernel here."""
    n, dim, m = 30000, 32, 32
    X = _gauss(n, dim, 250)
    h = zv.HNSW(m, 200)
    h.insert_batch(X)
    adj, _ = h.export_layer(0)
    h.set_kernel_variant(TEAM_ALWAYS)
    Q = _gauss(40, dim, 251)
    k, ef = 10, 700                      # 1 + 700 * 32 = 22 401 candidate slots + table: 292 KB > 227 KB
    got = h.search_batch(Q, k, ef, counters=True)
    ref = oracle.search_graph(X, adj, Q, ef, k, dist_mode=oracle.DIST_TREE, heap_mode=oracle.HEAP_DET)
    _bit_exact(zv, got, ref, k)
    h.deinit()
