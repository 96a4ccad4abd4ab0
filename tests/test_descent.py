"""K2, the upper-layer descent (extension). CPU: the oracle's descent against a
hand-derived case. GPU (-m gpu): the CUDA descent + layer-0 search through the C ABI against the
oracle on the same graph, bit-exact in the kernel's arithmetic and within 1e-5 in the reference's."""
import numpy as np
import pytest

RTOL = 1e-5
INV = 0xFFFFFFFF


def _tiny():
    """Five 1-d points at x = 0, 10, 20, 30, 40. Layer 0 is a path. Levels 0,1,0,2,1; layer 1:
    1:[3] 3:[1,4] 4:[3]; layer 2: 3:[] (the reference leaves a new top layer empty, hnsw.zig:114-116)."""
    X = np.array([[0.0], [10.0], [20.0], [30.0], [40.0]], np.float32)
    adj0 = np.array([[1, INV], [0, 2], [1, 3], [2, 4], [3, INV]], np.uint32)
    levels = np.array([0, 1, 0, 2, 1], np.uint8)
    base = np.array([INV, 0, INV, 1, 3], np.uint32)
    upper = np.array([[3, INV], [1, 4], [INV, INV], [3, INV]], np.uint32)   # lists: 1@L1, 3@L1, 3@L2, 4@L1
    return X, adj0, (levels, base, upper, 2, 3)


def test_oracle_descent_hand_case(oracle):
    """By hand: q = 12. Start 3 (d = 324). Layer 2: empty list, stay. Layer 1: scan [1 (4), 4 (784)] ->
    move to 1; rescan 1's list [3 (324)]: nothing closer. Lands on node 1 with 1 + 2 + 1 = 4 evaluations.
    Layer-0 search(q, 3) from node 1: pops 1 (4), 2 (64), 0 (144)."""
    O = oracle
    X, adj0, up = _tiny()
    node, d, ev = O.descend_one(X, up, np.array([12.0], np.float32))
    assert (node, d, ev) == (1, 4.0, 4)
    r = O.search_graph(X, adj0, np.array([[12.0]], np.float32), 3, 3, upper=up)
    assert r["ids"][0].tolist() == [1, 2, 0] and r["dist"][0].tolist() == [4.0, 64.0, 144.0]
    # evals: descent 4 (start included) + beam pushes 0,2 (pop 1), 3 (pop 2), none (pop 0) = 4 + 3
    assert r["evals"][0] == 7 and r["pops"][0] == 3
    # without the descent the same search starts at node 0 (entry_point)
    r0 = O.search_graph(X, adj0, np.array([[12.0]], np.float32), 3, 3)
    assert r0["ids"][0].tolist() == [1, 2, 0] and r0["evals"][0] == 4   # 0; 1; 2; 3
    # q = 39: layer 1 from 3 (81): [1 (841), 4 (1)] -> 4; rescan [3]: no. Lands on 4.
    assert O.descend_one(X, up, np.array([39.0], np.float32))[0] == 4


def test_oracle_descent_strict_less_keeps_first_minimum(oracle):
    """Two equidistant neighbours: the scan moves on strict < only (hnsw.zig:97), so the first one wins."""
    O = oracle
    X = np.array([[0.0], [5.0], [-5.0]], np.float32)
    levels = np.array([1, 1, 1], np.uint8)
    base = np.array([0, 1, 2], np.uint32)
    upper = np.array([[2, 1], [0, INV], [0, INV]], np.uint32)
    up = (levels, base, upper, 1, 0)
    # q = 0: nobody beats the start. q at +-4.9..: tie impossible; craft the tie with the start far away
    X2 = np.array([[100.0], [5.0], [-5.0]], np.float32)
    node, d, ev = O.descend_one(X2, up, np.array([0.0], np.float32))
    assert node == 2 and d == 25.0 and ev == 1 + 2 + 1     # list order [2, 1]: 2 is met first, 1 ties and stays out
    assert O.descend_one(X, up, np.array([0.0], np.float32))[0] == 0


def test_oracle_export_upper_matches_layers(oracle):
    O = oracle
    rng = np.random.default_rng(5)
    X = rng.standard_normal((2000, 8), dtype=np.float32)
    o = O.OracleHNSW(6, 200, seed=9)
    o.insert_batch(X)
    levels, base, adj, mx, start = o.export_upper()
    assert mx == o.max_level and levels[start] == mx and not np.any(levels[:start] == mx)
    for layer in range(1, mx + 1):
        tab, deg = o.export_layer(layer)
        for i in np.nonzero(levels >= layer)[0][:200]:
            assert np.array_equal(adj[base[i] + layer - 1], tab[i])


# ------------------------------------------------------------------------------------------------
gpu = pytest.mark.gpu


def _check(zv, O, h, X, Q, k, ef, up):
    adj, _ = h.export_layer(0)
    ids, dist, counts, pops, evals = h.search_batch(Q, k, ef, counters=True)
    ref = O.search_graph(X, adj, Q, ef, k, dist_mode=O.DIST_TREE, heap_mode=O.HEAP_DET, upper=up)
    assert np.array_equal(counts, ref["counts"])
    assert np.array_equal(ids, ref["ids"].astype(np.uint64))
    assert np.array_equal(dist.view(np.uint32), ref["dist"].view(np.uint32))
    assert np.array_equal(pops, ref["pops"]) and np.array_equal(evals, ref["evals"])
    seq = O.search_graph(X, adj, Q, ef, k, upper=up)
    mask = np.arange(k)[None, :] < counts[:, None]
    np.testing.assert_allclose(dist[mask], seq["dist"][mask], rtol=RTOL, atol=1e-30)
    return ids, evals


@gpu
def test_gpu_descent_tiny_hand_case(zv, oracle):
    X, adj0, up = _tiny()
    h = zv.HNSW(2, 200)
    h.load_padded_graph(X, adj0, 0)
    h.load_upper_layers(up[0], up[2], up[4])
    assert h.max_level == 2 and h.descent_start == 3
    lv, base, ua = h.export_upper_layers()
    assert np.array_equal(lv, up[0]) and np.array_equal(base, up[1]) and np.array_equal(ua, up[2])
    h.set_descent(True)
    ids, dist, counts, pops, evals = h.search_batch(np.array([[12.0]], np.float32), 3, 3, counters=True)
    assert ids[0].tolist() == [1, 2, 0] and dist[0].tolist() == [4.0, 64.0, 144.0] and evals[0] == 7
    h.set_descent(False)
    ids, dist, counts, pops, evals = h.search_batch(np.array([[12.0]], np.float32), 3, 3, counters=True)
    assert ids[0].tolist() == [1, 2, 0] and evals[0] == 4
    h.deinit()


@gpu
@pytest.mark.parametrize("dim,m", [(32, 8), (128, 16), (200, 40)])
def test_gpu_descent_on_the_reference_graph(zv, oracle, dim, m):
    """insert builds layers >= 1 exactly like the reference (hnsw.zig:88-108); the descent walks them."""
    O = oracle
    rng = np.random.default_rng(11)
    n = 6000
    X = rng.standard_normal((n, dim), dtype=np.float32)
    Q = rng.standard_normal((300, dim), dtype=np.float32)
    lv = np.minimum(rng.geometric(0.5, n) - 1, 31).astype(np.int32)
    h = zv.HNSW(m, 200)
    h.insert_batch(X, levels=lv)
    o = O.OracleHNSW(m, 200)
    o.insert_batch(X, levels=lv)
    up = o.export_upper()
    got = h.export_upper_layers()
    assert np.array_equal(got[0], up[0]) and np.array_equal(got[1], up[1]) and np.array_equal(got[2], up[2])
    assert h.descent_start == up[4] and h.max_level == up[3]
    h.set_descent(True)
    for k, ef in ((10, 10), (10, 64), (5, 200)):
        _check(zv, O, h, X, Q, k, ef, up)
    # inserting more nodes re-sends the upper layers
    X2 = rng.standard_normal((500, dim), dtype=np.float32)
    lv2 = np.minimum(rng.geometric(0.5, 500) - 1, 31).astype(np.int32)
    h.insert_batch(X2, levels=lv2)
    o.insert_batch(X2, levels=lv2)
    _check(zv, O, h, np.concatenate([X, X2]), Q, 10, 32, o.export_upper())
    h.set_descent(False)                          # back to the reference's search
    adj, _ = h.export_layer(0)
    ids, dist, counts = h.search_batch(Q, 10, 32)
    ref = O.search_graph(np.concatenate([X, X2]), adj, Q, 32, 10, dist_mode=O.DIST_TREE, heap_mode=O.HEAP_DET)
    assert np.array_equal(ids, ref["ids"].astype(np.uint64))
    h.deinit()


def _knn_layers(X, m, rng):
    """A small navigable hierarchy: random levels, exact m-NN lists among the members of each layer."""
    n = len(X)
    levels = np.minimum(rng.geometric(0.5, n) - 1, 31).astype(np.uint8)
    def knn(idx):
        P = X[idx].astype(np.float64)
        d = ((P[:, None, :] - P[None, :, :]) ** 2).sum(-1)
        np.fill_diagonal(d, np.inf)
        kk = min(m, len(idx) - 1)
        out = np.full((len(idx), m), INV, np.uint32)
        if kk > 0:
            out[:, :kk] = idx[np.argsort(d, axis=1, kind="stable")[:, :kk]]
        return out
    adj0 = knn(np.arange(n))
    base = np.full(n, INV, np.uint32)
    has = levels > 0
    base[has] = (np.cumsum(levels.astype(np.int64)) - levels)[has]
    upper = np.full((int(levels.sum()), m), INV, np.uint32)
    for layer in range(1, int(levels.max()) + 1):
        idx = np.nonzero(levels >= layer)[0]
        upper[base[idx].astype(np.int64) + layer - 1] = knn(idx)
    mx = int(levels.max())
    return adj0, (levels, base, upper, mx, int(np.argmax(levels == mx)))


@gpu
def test_gpu_descent_on_a_loaded_hierarchy(zv, oracle):
    O = oracle
    rng = np.random.default_rng(3)
    n, dim, m = 1500, 24, 12
    X = rng.standard_normal((n, dim), dtype=np.float32)
    Q = rng.standard_normal((256, dim), dtype=np.float32)
    adj0, up = _knn_layers(X, m, rng)
    h = zv.HNSW(m, 200)
    h.load_padded_graph(X, adj0, 0)
    h.load_upper_layers(up[0], up[2], up[4])
    h.set_descent(True)
    ids, ev_desc = _check(zv, O, h, X, Q, 10, 16, up)
    # the descent lands near the query: fewer pops are needed for the same answer quality than from node 0
    gt, _ = O.bruteforce(X, Q, 10)
    h.set_descent(False)
    ids0, _, _ = h.search_batch(Q, 10, 16)
    rec = lambda a: np.mean([len(set(a[i].tolist()) & set(gt[i].tolist())) / 10 for i in range(len(Q))])
    assert rec(ids) >= rec(ids0)
    with pytest.raises(zv.ZvdbError):
        h.load_upper_layers(up[0], up[2], int(np.argmin(up[0])))     # start node must have the maximum level
    h.deinit()


@gpu
def test_gpu_descent_large_batch_and_bitmap_mode(zv, oracle):
    """nq >= 4096 takes the two-stream chunk pipeline; ef = 300 takes the global-bitmap visited set."""
    O = oracle
    rng = np.random.default_rng(8)
    n, dim, m = 20000, 64, 16
    X = rng.standard_normal((n, dim), dtype=np.float32)
    Q = rng.standard_normal((5000, dim), dtype=np.float32)
    lv = np.minimum(rng.geometric(0.5, n) - 1, 31).astype(np.int32)
    h = zv.HNSW(m, 200)
    h.insert_batch(X, levels=lv)
    h.set_descent(True)
    adj, _ = h.export_layer(0)
    got = h.export_upper_layers()
    up = (got[0], got[1], got[2], h.max_level, h.descent_start)
    _check(zv, O, h, X, Q, 10, 32, up)
    _check(zv, O, h, X, Q[:600], 10, 300, up)
    h.deinit()


@pytest.mark.gpu
@pytest.mark.parametrize("metric_name", ["l2", "cos"])
def test_gpu_descent_on_a_quality_graph_with_a_built_hierarchy(zv, oracle, metric_name):
    """Round 2: builder graphs get upper layers too -- levels drawn like randomLevel
    (hnsw.zig:172-180), layer l = the same builder on the rows of level >= l (zvdb_b200/builder.py build_hierarchy) --
    so K2 has something worth walking. Descent + layer-0 search through the C ABI == orc_descend + orc_search on the
    exported hierarchy, bit for bit; and at a small pop budget the descent finds more true neighbours than node 0 does."""
    from zvdb_b200 import builder
    zm, om = {"l2": (zv.METRIC_L2, oracle.METRIC_L2), "cos": (zv.METRIC_COSINE, oracle.METRIC_COS)}[metric_name]
    n, dim, m = 20000, 32, 8
    X = np.random.default_rng(71).standard_normal((n, dim), dtype=np.float32)
    Q = np.random.default_rng(72).standard_normal((300, dim), dtype=np.float32)
    h = zv.HNSW(m, 200, metric=zm)
    builder.build_quality_graph(h, X, m, K=32)
    levels, upper, start = builder.build_hierarchy(h, X, m, seed=3, K=32)
    assert h.max_level == int(levels.max()) >= 8 and h.descent_start == start and h.entry_point == 0
    lv, ub, ua = h.export_upper_layers()
    assert np.array_equal(lv, levels) and np.array_equal(ua, upper)
    # every list of layer l holds only nodes that have layer l, and never the node itself
    for layer in (1, 2, 5):
        S = np.nonzero(levels >= layer)[0]
        lists = ua[ub[S].astype(np.int64) + (layer - 1)]
        valid = lists != INV
        assert np.all(levels[lists[valid]] >= layer)
        assert not np.any(lists == S[:, None].astype(np.uint32))
    Xs = np.stack([h.point(i) for i in range(n)]).astype(np.float32)
    adj, _ = h.export_layer(0)
    up = (lv, ub, ua, h.max_level, h.descent_start)
    gt, _ = oracle.bruteforce(Xs, Q, 10, metric=zm)

    def recall(ids):
        return float(np.mean([len(set(ids[i].tolist()) & set(gt[i].tolist())) / 10 for i in range(len(Q))]))

    plain = h.search_batch(Q, 10, 16, counters=True)
    h.set_descent(True)
    for k, ef in ((10, 16), (10, 64), (5, 300)):
        ids, dist, counts, pops, evals = h.search_batch(Q, k, ef, counters=True)
        ref = oracle.search_graph(Xs, adj, Q, ef, k, dist_mode=oracle.DIST_TREE | om, heap_mode=oracle.HEAP_DET, upper=up)
        assert np.array_equal(counts, ref["counts"]) and np.array_equal(pops, ref["pops"]) and np.array_equal(evals, ref["evals"])
        assert np.array_equal(ids, ref["ids"].astype(np.uint64))
        assert np.array_equal(dist.view(np.uint32), ref["dist"].view(np.uint32))
        seq = oracle.search_graph(Xs, adj, Q, ef, k, dist_mode=oracle.DIST_SEQ | om, heap_mode=oracle.HEAP_ZIG, upper=up)
        np.testing.assert_allclose(dist, seq["dist"], rtol=RTOL, atol=1e-30)
    with_descent = h.search_batch(Q, 10, 16)
    r0, r1 = recall(plain[0]), recall(with_descent[0])
    print(f"recall@10 at 16 pops ({metric_name}): from node 0 {r0:.3f}, after the descent {r1:.3f}")
    assert r1 > r0
    h.deinit()
