"""On-disk format: save -> load gives the same index bit for bit -- same layers,
same search results, and the same graph growth under further inserts. The reference has no
persistence, so the oracle here is the index itself before the round trip (+ the CPU oracle for
the searches)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _gauss(n, dim, seed):
    return np.random.default_rng(seed).standard_normal((n, dim), dtype=np.float32)


def test_save_load_round_trip(zv, oracle, tmp_path):
    X, Q = _gauss(5000, 48, 31), _gauss(200, 48, 32)
    a = zv.HNSW(12, 200, level_seed=7)
    a.insert_batch(X[:4000])
    path = os.path.join(tmp_path, "index.zvdb")
    a.save(path)
    b = zv.HNSW(12, 200)
    b.load(path)
    assert b.count() == 4000 and b.dim == 48 and b.max_level == a.max_level and b.entry_point == 0
    assert b.descent_start == a.descent_start
    for layer in range(a.max_level + 1):
        assert np.array_equal(a.export_layer(layer)[0], b.export_layer(layer)[0])
    ua, ub = a.export_upper_layers(), b.export_upper_layers()
    assert all(np.array_equal(x, y) for x, y in zip(ua, ub))
    assert np.array_equal(a.point(3999), b.point(3999))
    for descent in (False, True):
        a.set_descent(descent); b.set_descent(descent)
        ra, rb = a.search_batch(Q, 10, 64, counters=True), b.search_batch(Q, 10, 64, counters=True)
        assert all(np.array_equal(x, y) for x, y in zip(ra, rb))
    a.set_descent(False); b.set_descent(False)
    adj, _ = b.export_layer(0)
    ref = oracle.search_graph(X[:4000], adj, Q, 64, 10, dist_mode=oracle.DIST_TREE, heap_mode=oracle.HEAP_DET)
    ids, dist, counts = b.search_batch(Q, 10, 64)
    assert np.array_equal(ids, ref["ids"].astype(np.uint64)) and np.array_equal(dist.view(np.uint32), ref["dist"].view(np.uint32))
    # the level generator state travels: both indexes grow the same graph from here on
    a.insert_batch(X[4000:]); b.insert_batch(X[4000:])
    assert a.max_level == b.max_level
    for layer in range(a.max_level + 1):
        assert np.array_equal(a.export_layer(layer)[0], b.export_layer(layer)[0])
    assert np.array_equal(a.bruteforce_knn(Q[:20], 5)[0], b.bruteforce_knn(Q[:20], 5)[0])
    a.deinit(); b.deinit()


def test_load_rejects_bad_files(zv, tmp_path):
    a = zv.HNSW(8, 200)
    a.insert_batch(_gauss(300, 16, 33))
    path = os.path.join(tmp_path, "i.zvdb")
    a.save(path)
    raw = bytearray(open(path, "rb").read())
    with pytest.raises(zv.ZvdbError):                       # wrong m
        zv.HNSW(16, 200).load(path)
    with pytest.raises(zv.ZvdbError):                       # wrong metric
        zv.HNSW(8, 200, metric=zv.METRIC_DOT).load(path)
    with pytest.raises(zv.ZvdbError):                       # missing file
        zv.HNSW(8, 200).load(os.path.join(tmp_path, "nope"))
    flipped = bytearray(raw); flipped[len(raw) // 2] ^= 0x40
    open(path + ".flip", "wb").write(flipped)
    with pytest.raises(zv.ZvdbError, match="checksum|corrupt|out of range"):
        zv.HNSW(8, 200).load(path + ".flip")
    open(path + ".cut", "wb").write(raw[: len(raw) - 100])
    with pytest.raises(zv.ZvdbError):
        zv.HNSW(8, 200).load(path + ".cut")
    open(path + ".magic", "wb").write(b"NOTZVDB!" + raw[8:])
    with pytest.raises(zv.ZvdbError, match="magic"):
        zv.HNSW(8, 200).load(path + ".magic")
    # a failed load leaves the target untouched
    b = zv.HNSW(8, 200)
    b.insert([1.0] * 16)
    with pytest.raises(zv.ZvdbError):
        b.load(path + ".cut")
    assert b.count() == 1
    # empty index round trip
    e = zv.HNSW(8, 200)
    e.save(path + ".empty")
    b.load(path + ".empty")
    assert b.count() == 0 and b.search([0.0] * 16, 3) == []
    a.deinit(); b.deinit(); e.deinit()


def test_handle_reused_with_a_larger_dim(zv, oracle, tmp_path):
    """A handle whose device buffers were sized for 32-d rows is given 128-d rows by load_graph, by
    build_from_candidates and by load: the device copy must be re-sized for the new row pitch (capacity is
    rows x pitch, not rows), and results must equal a fresh index's and the oracle's."""
    small, big = _gauss(1000, 32, 41), _gauss(900, 128, 42)
    Q = _gauss(64, 128, 43)
    fresh = zv.HNSW(8, 200)
    fresh.insert_batch(big)
    adj, _ = fresh.export_layer(0)
    want = fresh.search_batch(Q, 5, 40, counters=True)
    ref = oracle.search_graph(big, adj, Q, 40, 5, dist_mode=oracle.DIST_TREE, heap_mode=oracle.HEAP_DET)
    assert np.array_equal(want[0], ref["ids"].astype(np.uint64))
    path = os.path.join(tmp_path, "big.zvdb")
    fresh.save(path)

    def used_handle():
        h = zv.HNSW(8, 200)
        h.insert_batch(small)
        assert np.all(h.search_batch(small[:16], 3, 12)[2] == 3)     # device buffers now exist, 32 floats per row
        return h

    h = used_handle()
    h.load_padded_graph(big, adj, entry=0)
    assert h.dim == 128 and h.count() == 900
    got = h.search_batch(Q, 5, 40, counters=True)
    assert all(np.array_equal(a, b) for a, b in zip(want, got))
    assert np.array_equal(h.bruteforce_knn(Q[:8], 5)[0], fresh.bruteforce_knn(Q[:8], 5)[0])
    h.deinit()

    h = used_handle()
    h.load(path)
    got = h.search_batch(Q, 5, 40, counters=True)
    assert all(np.array_equal(a, b) for a, b in zip(want, got))
    h.deinit()

    h = used_handle()
    nn, _ = oracle.bruteforce(big, big, 24)
    h.build_from_candidates(big, nn.astype(np.uint32))
    g2 = zv.HNSW(8, 200)
    g2.build_from_candidates(big, nn.astype(np.uint32))
    assert np.array_equal(h.export_layer(0)[0], g2.export_layer(0)[0])
    assert all(np.array_equal(a, b) for a, b in zip(h.search_batch(Q, 5, 40), g2.search_batch(Q, 5, 40)))
    # and back to a smaller dim on the same handle
    h.load_padded_graph(small, np.full((1000, 8), 0xFFFFFFFF, np.uint32), entry=0)
    assert h.dim == 32 and np.all(h.search_batch(small[:4], 3, 3)[2] == 1)
    h.deinit(); g2.deinit(); fresh.deinit()


def test_wrapper_refuses_an_index_of_another_element_type(zv, tmp_path):
    """HNSW(T) wrappers are typed at construction: loading a file of another T, or a (float32) graph into an
    HNSW(f64), raises instead of reading rows with the wrong element size."""
    a = zv.HNSW(8, 200, dtype=np.float64)
    a.insert_batch(np.random.default_rng(44).standard_normal((200, 16)))
    path = os.path.join(tmp_path, "f64.zvdb")
    a.save(path)
    b = zv.HNSW(8, 200)                       # an f32 wrapper
    with pytest.raises(TypeError):
        b.load(path)
    c = zv.HNSW(8, 200, dtype=np.float64)
    c.load(path)
    assert np.array_equal(c.point(7), a.point(7))
    with pytest.raises(TypeError):
        c.load_padded_graph(_gauss(10, 16, 45), np.full((10, 8), 0xFFFFFFFF, np.uint32))
    a.deinit(); b.deinit(); c.deinit()
