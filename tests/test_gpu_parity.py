"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C ABI, against
the CPU oracle on the same seeded inputs.

Two comparison levels:
  * bit-exact: ids, order, distance bits and the pop/eval counters equal the oracle run in the
    kernel's own arithmetic (ORC_DIST_TREE summation order, ORC_HEAP_DET tie-break);
  * reference-faithful: against the oracle in the reference's arithmetic (sequential sum, Zig
    heap): distances within 1e-5 relative and ids/order identical except where distances tie
    within 1e-5 relative (the tolerance BASELINE.json's north_star states).
The first half of the file restates src/test_hnsw.zig test by test.
"""
import json
import os
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RTOL = 1e-5   # north_star: "distances within 1e-5 relative"
GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_tests.json")))


def _gauss(n, dim, seed):
    return np.random.default_rng(seed).standard_normal((n, dim), dtype=np.float32)


def _uniform(n, dim, seed):
    return np.random.default_rng(seed).random((n, dim), dtype=np.float32)


def euclidean_distance(a, b):   # test_hnsw.zig:15-22
    return float(np.sqrt(((np.asarray(a, np.float32) - np.asarray(b, np.float32)) ** 2).sum(dtype=np.float32)))


def assert_reference_faithful(ids, dist, counts, ref):
    """Against the oracle in the REFERENCE's arithmetic (sequential sum, Zig heap): same result
    counts, distances within RTOL position by position, and ids/order identical except at near-ties.
    A position whose id differs is, by the distance check, a pair of nodes whose distances tie within
    RTOL (an order swap of two near-equal neighbours, or the last pop choosing the other of two
    near-equal candidates); such positions must be rare."""
    assert np.array_equal(counts, ref["counts"])
    mask = np.arange(ids.shape[1])[None, :] < counts[:, None]
    np.testing.assert_allclose(dist[mask], ref["dist"][mask], rtol=RTOL, atol=1e-30)
    differ = (ids != ref["ids"].astype(np.uint64)) & mask
    assert differ.sum() <= max(2, 0.002 * mask.sum()), f"{differ.sum()} of {mask.sum()} ids differ"


# ------------------------------------------------------------------------------------------------
# src/test_hnsw.zig, test by test
# ------------------------------------------------------------------------------------------------

def test_basic_functionality(zv):                       # test_hnsw.zig:24-41
    hnsw = zv.HNSW(16, 200)
    hnsw.insert([1, 2, 3])
    hnsw.insert([4, 5, 6])
    hnsw.insert([7, 8, 9])
    query = [3, 4, 5]
    results = hnsw.search(query, 2)
    assert len(results) == 2
    assert euclidean_distance(query, results[0].point) <= euclidean_distance(query, results[1].point)
    g = GOLDEN["G1_basic"]
    assert [r.id for r in results] == g["ids"] and [r.distance for r in results] == g["dist"]
    hnsw.deinit()


def test_empty_index(zv):                               # test_hnsw.zig:43-53
    hnsw = zv.HNSW(16, 200)
    assert hnsw.search([1, 2, 3], 5) == []
    ids, dist, counts = hnsw.search_batch(np.zeros((4, 3), np.float32), 5)
    assert np.all(counts == 0) and np.all(ids == zv.INVALID_ID)
    hnsw.deinit()


def test_single_point(zv):                              # test_hnsw.zig:55-68
    hnsw = zv.HNSW(16, 200)
    point = np.array([1, 2, 3], np.float32)
    hnsw.insert(point)
    results = hnsw.search(point, 1)
    assert len(results) == 1
    assert np.array_equal(results[0].point, point)
    hnsw.deinit()


def test_large_dataset(zv):                             # test_hnsw.zig:70-102
    hnsw = zv.HNSW(16, 200)
    X = _uniform(10000, 128, 21)
    for p in X[:100]:
        hnsw.insert(p)                                  # the per-point call ...
    hnsw.insert_batch(X[100:])                          # ... and its batched form
    query = _uniform(1, 128, 22)[0]
    k = 10
    results = hnsw.search(query, k)
    assert len(results) == k
    last = 0.0
    for r in results:
        d = euclidean_distance(query, r.point)
        assert d >= last
        last = d
    hnsw.deinit()


def test_edge_cases(zv):                                # test_hnsw.zig:104-126
    hnsw = zv.HNSW(16, 200)
    point = np.array([1, 2, 3], np.float32)
    hnsw.insert(point)
    hnsw.insert(point)
    results = hnsw.search(point, 2)
    assert len(results) == 2
    assert np.array_equal(results[0].point, point) and np.array_equal(results[1].point, point)
    assert [r.id for r in results] == GOLDEN["G6_duplicates_k2"]["ids"]
    large_k = hnsw.search(point, 100)
    assert len(large_k) == 2
    hnsw.deinit()


def test_index_owns_its_points(zv):                     # test_hnsw.zig:128-152 (Memory Leaks): insert copies
    hnsw = zv.HNSW(16, 200)
    X = _uniform(1000, 64, 23)
    for i in range(1000):
        p = X[i].copy()
        hnsw.insert(p)
        p[:] = -1.0                                     # caller's buffer is dead after insert
    res = hnsw.search(_uniform(1, 64, 24)[0], 10)
    assert len(res) == 10
    assert np.array_equal(hnsw.point(17), X[17])
    hnsw.deinit()


def test_concurrent_access(zv):                         # test_hnsw.zig:154-209
    hnsw = zv.HNSW(16, 200)
    num_threads, per_thread, dim = 8, 1000, 128

    def work(t):
        pts = _uniform(per_thread, dim, 100 + t)
        for p in pts:
            hnsw.insert(p)

    ts = [threading.Thread(target=work, args=(t,)) for t in range(num_threads)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert hnsw.nodes.count() == num_threads * per_thread
    assert len(hnsw.search(_uniform(1, dim, 99)[0], 10)) == 10
    hnsw.deinit()


def test_stress(zv):                                    # test_hnsw.zig:211-237
    hnsw = zv.HNSW(16, 200)
    hnsw.insert_batch(_uniform(100000, 128, 25))
    Q = _uniform(100, 128, 26)
    for q in Q[:10]:
        assert len(hnsw.search(q, 10)) == 10
    ids, dist, counts = hnsw.search_batch(Q, 10)
    assert np.all(counts == 10)
    assert np.all(np.diff(dist, axis=1) >= 0)
    hnsw.deinit()


def test_consistency(zv):                               # test_hnsw.zig:275-317
    hnsw = zv.HNSW(16, 200)
    hnsw.insert_batch(_uniform(10000, 128, 27))
    query = _uniform(1, 128, 28)[0]
    first = None
    for i in range(10):
        res = hnsw.search(query, 10)
        pts = np.stack([r.point for r in res])
        if i == 0:
            first = pts
        else:
            assert np.array_equal(first, pts)
    hnsw.deinit()


# ------------------------------------------------------------------------------------------------
# golden vectors through the CUDA path
# ------------------------------------------------------------------------------------------------

@pytest.mark.parametrize("name", [k for k, v in GOLDEN.items() if not k.startswith("_") and v.get("dtype") == "f32"])
def test_reference_golden_on_gpu(zv, name):
    g = GOLDEN[name]
    hnsw = zv.HNSW(g["m"], 200)
    for p in g["points"]:
        hnsw.insert(p)
    if "layer0" in g:
        adj, deg = hnsw.export_layer(0)
        assert [[int(x) for x in adj[i, :deg[i]]] for i in range(hnsw.count())] == g["layer0"]
    q = np.asarray(g["query"], np.float32)[None]
    if hnsw.count() == 0:
        assert hnsw.search(g["query"], g["k"]) == []
        return
    ids, dist, counts, pops, evals = hnsw.search_batch(q, g["k"], counters=True)
    c = int(counts[0])
    # on exact distance ties the kernel's order is (distance, id) = 'ids_det' (north_star exempts exact ties; G11)
    assert [int(x) for x in ids[0, :c]] == g.get("ids_det", g["ids"])
    assert [float(x) for x in dist[0, :c]] == [float(x) for x in g["dist"]]
    if "evals" in g:
        assert (int(pops[0]), int(evals[0])) == (g["pops"], g["evals"])
    hnsw.deinit()


# ------------------------------------------------------------------------------------------------
# producer parity: the graph built by zvdb_insert is the oracle's graph
# ------------------------------------------------------------------------------------------------

@pytest.mark.parametrize("n,dim,m", [(3000, 128, 16), (2000, 24, 4), (1500, 200, 32), (500, 3, 2)])
def test_insert_builds_the_reference_graph(zv, oracle, n, dim, m):
    X = _gauss(n, dim, 31)
    levels = np.random.default_rng(32).geometric(0.5, n).astype(np.int32) - 1
    o = oracle.OracleHNSW(m, 200)
    o.insert_batch(X, levels=levels)
    h = zv.HNSW(m, 200)
    h.insert_batch(X, levels=levels)
    assert h.count() == o.count() and h.max_level == o.max_level and h.entry_point == o.entry_point == 0
    for layer in range(0, min(o.max_level, 4) + 1):
        ao, do_ = o.export_layer(layer)
        ah, dh = h.export_layer(layer)
        assert np.array_equal(do_, dh), layer
        assert np.array_equal(ao, ah), layer
    h.deinit()


# ------------------------------------------------------------------------------------------------
# search parity
# ------------------------------------------------------------------------------------------------

def _build_pair(zv, oracle, n, dim, m, seed, metric=0):
    X = _gauss(n, dim, seed)
    h = zv.HNSW(m, 200, metric=metric)
    h.insert_batch(X)
    adj, _ = h.export_layer(0)
    return X, h, adj


@pytest.mark.parametrize("n,dim,m,k,ef", [
    (10000, 128, 16, 10, 10),     # C1: the reference call, ef = k
    (10000, 128, 16, 10, 64),     # C1 with the config's "ef_search = 64" = search(q,64)[:10]
    (10000, 128, 16, 10, 512),    # bitmap mode, popped keys in global scratch
    (10000, 128, 16, 300, 600),   # ... with a long result list read back from it
    (4000, 3, 16, 5, 20),         # tiny dim (padding lanes)
    (4000, 200, 16, 10, 40),      # dim not a multiple of 128 floats
    (3000, 768, 32, 100, 128),    # C3 shape: 6 chunks per lane, M = 32, k = 100
    (3000, 1024, 8, 10, 16),
    (2000, 64, 40, 10, 30),       # m > 32: two adjacency passes per pop
])
def test_search_bit_exact_vs_oracle(zv, oracle, n, dim, m, k, ef):
    X, h, adj = _build_pair(zv, oracle, n, dim, m, 41)
    Q = _gauss(257, dim, 42)
    ids, dist, counts, pops, evals = h.search_batch(Q, k, ef, counters=True)
    ref = oracle.search_graph(X, adj, Q, ef, k, dist_mode=oracle.DIST_TREE, heap_mode=oracle.HEAP_DET)
    assert np.array_equal(counts, ref["counts"])
    assert np.array_equal(pops, ref["pops"])
    assert np.array_equal(evals, ref["evals"])
    mask = np.arange(k)[None, :] < counts[:, None]
    assert np.array_equal(ids[mask], ref["ids"].astype(np.uint64)[mask])
    assert np.array_equal(dist.view(np.uint32)[mask], ref["dist"].view(np.uint32)[mask])
    assert np.all(ids[~mask] == zv.INVALID_ID)
    # and against the reference's own arithmetic and heap, within the stated tolerance
    faithful = oracle.search_graph(X, adj, Q, ef, k, dist_mode=oracle.DIST_SEQ, heap_mode=oracle.HEAP_ZIG)
    assert_reference_faithful(ids, dist, counts, faithful)
    h.deinit()


@pytest.mark.parametrize("warps", [0b0100, 0b1000, 0b1100,             # visited set: shared-memory hash, global bitmap, global hash
                                   0b0101, 0b1010,                     # (bits 0-1, round 1's load widths: accepted, ignored)
                                   0x104, 0x304, 0x404,                # x L2 prefetch {off, rows, adjacency, both}
                                   0x108, 0x208, 0x308, 0x408, 0x10C, 0x20C, 0x30C, 0x40C])
def test_kernel_variant_does_not_change_results(zv, oracle, warps):
    X, h, adj = _build_pair(zv, oracle, 6000, 128, 16, 43)
    Q = _gauss(300, 128, 44)
    base = h.search_batch(Q, 10, 96, counters=True)
    h.set_kernel_variant(warps)
    got = h.search_batch(Q, 10, 96, counters=True)
    for a, b in zip(base, got):
        assert np.array_equal(a.view(np.uint8), b.view(np.uint8))
    h.deinit()


@pytest.mark.parametrize("vis", [0b1000, 0b1100])                      # global bitmap, global hash
@pytest.mark.parametrize("n,dim,m,k,ef", [
    (20000, 128, 16, 10, 512),    # the large-ef path of C2: persistent CTAs, popped keys in global scratch
    (3000, 128, 16, 10, 700),     # ef * m > n: the table is sized by n, nearly every row is visited
    (6000, 64, 40, 10, 200),      # m > 32: two adjacency passes per pop
    (300, 16, 4, 5, 300),         # tiny index, ef = n: the search runs dry
])
def test_global_visited_modes_bit_exact_vs_oracle(zv, oracle, vis, n, dim, m, k, ef):
    """The two global-memory visited sets (the per-CTA hash table sized by ef*m that round 2 made the default beyond
    ef = 64, and round 1's n-bit bitmap) against the oracle, with many more queries than resident CTAs' worth of
    table reuse would need to expose a table that is not left clean."""
    X, h, adj = _build_pair(zv, oracle, n, dim, m, 45)
    h.set_kernel_variant(vis)
    Q = _gauss(700, dim, 46)
    for rep in range(2):                                # second pass reuses every CTA's table
        ids, dist, counts, pops, evals = h.search_batch(Q, k, ef, counters=True)
        ref = oracle.search_graph(X, adj, Q, ef, k, dist_mode=oracle.DIST_TREE, heap_mode=oracle.HEAP_DET)
        assert np.array_equal(counts, ref["counts"]) and np.array_equal(pops, ref["pops"]) and np.array_equal(evals, ref["evals"])
        mask = np.arange(k)[None, :] < counts[:, None]
        assert np.array_equal(ids[mask], ref["ids"].astype(np.uint64)[mask])
        assert np.array_equal(dist.view(np.uint32)[mask], ref["dist"].view(np.uint32)[mask])
    # a different ef on the same handle: another table pitch over the same scratch
    ids, dist, counts = h.search_batch(Q, k, max(k, ef // 3))
    ref = oracle.search_graph(X, adj, Q, max(k, ef // 3), k, dist_mode=oracle.DIST_TREE, heap_mode=oracle.HEAP_DET)
    mask = np.arange(k)[None, :] < counts[:, None]
    assert np.array_equal(ids[mask], ref["ids"].astype(np.uint64)[mask])
    h.deinit()


def test_batch_size_does_not_change_results(zv):
    """Determinism across batch sizes (a property the reference never checks)."""
    h = zv.HNSW(16, 200)
    h.insert_batch(_gauss(8000, 128, 45))
    Q = _gauss(512, 128, 46)
    ids, dist, counts = h.search_batch(Q, 10, 48)
    for bs in (1, 7, 64):
        for s in range(0, 128, bs):
            i2, d2, c2 = h.search_batch(Q[s:s + bs], 10, 48)
            assert np.array_equal(i2, ids[s:s + bs]) and np.array_equal(d2.view(np.uint32), dist[s:s + bs].view(np.uint32))
    one = h.search(Q[5], 10)   # search(q, k) == search_batch(ef = k)
    i3, d3, c3 = h.search_batch(Q[5:6], 10, 10)
    assert [r.id for r in one] == [int(x) for x in i3[0, :c3[0]]]
    h.deinit()


def test_search_on_external_quality_graph(zv, oracle):
    """Throughput track: the same kernel on a graph from another builder -- here an
    exact k-NN graph plus reverse edges, <= m per node -- still equals the oracle on that graph."""
    n, dim, m = 5000, 64, 16
    X = _gauss(n, dim, 47)
    nn, _ = oracle.bruteforce(X, X, 9)
    adj = np.full((n, m), 0xFFFFFFFF, np.uint32)
    deg = np.zeros(n, np.int64)
    for i in range(n):
        for j in nn[i, 1:9]:
            adj[i, deg[i]] = j
            deg[i] += 1
    for i in range(n):
        for j in nn[i, 1:9]:
            if deg[j] < m and i not in adj[j, :deg[j]]:
                adj[j, deg[j]] = i
                deg[j] += 1
    h = zv.HNSW(m, 200)
    h.load_padded_graph(X, adj, entry=0)
    Q = _gauss(200, dim, 48)
    for ef in (10, 100, 400):
        ids, dist, counts, pops, evals = h.search_batch(Q, 10, ef, counters=True)
        ref = oracle.search_graph(X, adj, Q, ef, 10, dist_mode=oracle.DIST_TREE, heap_mode=oracle.HEAP_DET)
        assert np.array_equal(ids, ref["ids"].astype(np.uint64))
        assert np.array_equal(dist.view(np.uint32), ref["dist"].view(np.uint32))
        assert np.array_equal(evals, ref["evals"]) and np.array_equal(pops, ref["pops"])
    gt, _ = oracle.bruteforce(X, Q, 10)
    rec = np.mean([len(set(ids[i].tolist()) & set(gt[i].tolist())) / 10 for i in range(len(Q))])
    assert rec > 0.9
    h.deinit()


def test_incremental_insert_between_searches(zv, oracle):
    """insert -> search -> insert -> search: the device copy follows the host graph (scatter path)."""
    X = _gauss(6000, 32, 49)
    Q = _gauss(64, 32, 50)
    h = zv.HNSW(8, 200)
    o = oracle.OracleHNSW(8, 200)
    done = 0
    for upto in (1, 2, 50, 3000, 3010, 6000):
        h.insert_batch(X[done:upto])
        o.insert_batch(X[done:upto])
        done = upto
        adj, _ = o.export_layer(0)
        ids, dist, counts = h.search_batch(Q, 5, 12)
        ref = oracle.search_graph(X[:upto], adj, Q, 12, 5, dist_mode=oracle.DIST_TREE, heap_mode=oracle.HEAP_DET)
        assert np.array_equal(counts, ref["counts"])
        mask = np.arange(5)[None, :] < counts[:, None]
        assert np.array_equal(ids[mask], ref["ids"].astype(np.uint64)[mask])
    h.deinit()


def test_dim_mismatch_is_an_error(zv):
    h = zv.HNSW(16, 200)
    h.insert([1, 2, 3])
    with pytest.raises(zv.ZvdbError) as e:
        h.insert([1, 2, 3, 4])
    assert e.value.code == zv._lib.ERR_DIM_MISMATCH and "Mismatched dimensions" in str(e.value)
    with pytest.raises(zv.ZvdbError):
        h.search([1, 2], 1)
    h.deinit()


def test_device_buffer_entry_point_and_id_mapping(zv):
    import torch
    h = zv.HNSW(16, 200)
    h.insert_batch(_gauss(5000, 128, 51))
    Q = _gauss(100, 128, 52)
    ids, dist, counts = h.search_batch(Q, 10, 32)
    dq = torch.from_numpy(Q).cuda()
    d_ids = torch.empty((100, 10), dtype=torch.int64, device="cuda")
    d_dist = torch.empty((100, 10), dtype=torch.float32, device="cuda")
    d_cnt = torch.empty(100, dtype=torch.int32, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    h.search_batch_device(dq.data_ptr(), 100, 10, 32, d_ids.data_ptr(), d_dist.data_ptr(), d_cnt.data_ptr(), stream=s)
    h.search_batch_device(dq.data_ptr(), 100, 10, 32, d_ids.data_ptr(), d_dist.data_ptr(), d_cnt.data_ptr(),
                          id_stride=8, id_base=3, stream=s)
    torch.cuda.synchronize()
    assert np.array_equal(d_ids.cpu().numpy().astype(np.uint64), ids * 8 + 3)
    assert np.array_equal(d_dist.cpu().numpy(), dist)
    h.deinit()


def test_shard_merge_kernel(zv, oracle):
    import torch
    rng = np.random.default_rng(53)
    for G, nq, k in ((2, 300, 10), (8, 200, 10), (8, 50, 100), (4, 64, 1)):
        d = np.sort(rng.random((G, nq, k), dtype=np.float32), axis=2)
        d[:, : nq // 4, :] = np.round(d[:, : nq // 4, :], 1)          # force cross-shard distance ties
        d = np.sort(d, axis=2)
        ids = rng.permutation(G * nq * k).astype(np.uint64).reshape(G, nq, k)
        cnt = rng.integers(0, k + 1, (G, nq)).astype(np.uint32)
        do, io, co = oracle.merge_topk(d, ids, cnt)
        td, ti, tc = (torch.from_numpy(a).cuda() for a in (d, ids.view(np.int64), cnt.view(np.int32)))
        od = torch.empty((nq, k), dtype=torch.float32, device="cuda")
        oi = torch.empty((nq, k), dtype=torch.int64, device="cuda")
        oc = torch.empty(nq, dtype=torch.int32, device="cuda")
        zv.merge_topk_device(td.data_ptr(), ti.data_ptr(), tc.data_ptr(), G, nq, k, od.data_ptr(), oi.data_ptr(),
                             oc.data_ptr(), torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        assert np.array_equal(oc.cpu().numpy().view(np.uint32), co)
        assert np.array_equal(oi.cpu().numpy().view(np.uint64), io)
        assert np.array_equal(od.cpu().numpy(), do)


# ------------------------------------------------------------------------------------------------
# the GPU graph builder (extension; checked against tests/builder_ref.py, not against hnsw.zig)
# ------------------------------------------------------------------------------------------------

@pytest.mark.parametrize("n,dim,m,K", [(1500, 32, 8, 24), (1200, 128, 16, 48), (600, 200, 4, 128)])
def test_builder_matches_cpu_statement(zv, oracle, n, dim, m, K):
    from builder_ref import build_ref
    X = _gauss(n, dim, 61)
    nn, _ = oracle.bruteforce(X, X, K)          # includes self at rank 0: the builder must drop it
    cand = nn.astype(np.uint32)
    cand[::7, 3] = 0xFFFFFFFF                   # padding and duplicates are tolerated
    cand[::5, 5] = cand[::5, 4]
    h = zv.HNSW(m, 200)
    h.build_from_candidates(X, cand)
    adj, deg = h.export_layer(0)
    ref = build_ref(oracle, X, cand, m)
    assert np.array_equal(adj, ref)
    assert h.entry_point == 0 and h.count() == n
    # the built graph is searched by the same kernel with the same parity
    Q = _gauss(64, dim, 62)
    ids, dist, counts, pops, evals = h.search_batch(Q, 5, 30, counters=True)
    r = oracle.search_graph(X, adj, Q, 30, 5, dist_mode=oracle.DIST_TREE, heap_mode=oracle.HEAP_DET)
    assert np.array_equal(ids, r["ids"].astype(np.uint64)) and np.array_equal(evals, r["evals"])
    h.deinit()


def test_builder_graph_reaches_useful_recall(zv, oracle):
    n, dim, m = 20000, 64, 16
    X = _gauss(n, dim, 63)
    nn, _ = oracle.bruteforce(X, X, 49)
    h = zv.HNSW(m, 200)
    h.build_from_candidates(X, nn[:, 1:].astype(np.uint32))
    Q = _gauss(200, dim, 64)
    gt, _ = oracle.bruteforce(X, Q, 10)
    ids, _, _ = h.search_batch(Q, 10, 256)
    rec = np.mean([len(set(ids[i].tolist()) & set(gt[i].tolist())) / 10 for i in range(len(Q))])
    assert rec > 0.9, rec
    h.deinit()


def test_incremental_builder_is_close_to_the_exact_candidate_builder(zv, oracle):
    """The search-driven builder (candidates = the index's own search on the growing graph) against the same builder
    fed exact k-NN candidates: same layout, every row linked, recall in the same range."""
    from zvdb_b200 import builder
    n, dim, m, K = 20000, 64, 16, 48
    X = _gauss(n, dim, 65)
    Q = _gauss(200, dim, 66)
    gt, _ = oracle.bruteforce(X, Q, 10)

    def recall(h):
        ids, _, _ = h.search_batch(Q, 10, 256)
        return np.mean([len(set(ids[i].tolist()) & set(gt[i].tolist())) / 10 for i in range(len(Q))])

    nn, _ = oracle.bruteforce(X, X, K + 1)
    exact = zv.HNSW(m, 200)
    exact.build_from_candidates(X, nn[:, 1:].astype(np.uint32))
    inc = zv.HNSW(m, 200)
    stats = builder.build_quality_graph_incremental(inc, X, m, K=K, seed_rows=2048, ef=256)
    assert stats["phases"] >= 4 and inc.count() == n and inc.entry_point == 0
    adj, deg = inc.export_layer(0)
    assert adj.shape == (n, m) and deg.min() >= 1
    r_exact, r_inc = recall(exact), recall(inc)
    print(f"recall@10 at ef=256: exact candidates {r_exact:.3f}, incremental {r_inc:.3f}")
    assert r_inc >= r_exact - 0.15, (r_exact, r_inc)      # measured: 0.903 vs 0.887 (256 pops per construction search; 0.82 at 64)
    # searched by the same kernel with the same parity
    ids, dist, counts, pops, evals = inc.search_batch(Q, 5, 30, counters=True)
    r = oracle.search_graph(X, adj, Q, 30, 5, dist_mode=oracle.DIST_TREE, heap_mode=oracle.HEAP_DET)
    assert np.array_equal(ids, r["ids"].astype(np.uint64)) and np.array_equal(evals, r["evals"])
    exact.deinit(); inc.deinit()
