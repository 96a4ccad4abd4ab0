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
