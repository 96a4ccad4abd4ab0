"""HNSW(T) for T = i32 / f64 (the reference is generic, src/hnsw.zig:8; its "Different Data Types" test,
test_hnsw.zig:239-273). The caller's rows are kept in T, the graph is built comparing distances in T's
arithmetic (so it is the reference's graph), the search runs on the f32 conversion."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_tests.json")))


def test_different_data_types(zv):                      # test_hnsw.zig:239-273
    for name, dt in (("G2_i32", np.int32), ("G3_f64", np.float64)):
        g = GOLDEN[name]
        hnsw = zv.HNSW(16, 200, dtype=dt)
        for p in g["points"]:
            hnsw.insert(p)
        results = hnsw.search(g["query"], g["k"])
        assert len(results) == 2                                            # the reference's assertion
        assert [r.id for r in results] == g["ids"]
        np.testing.assert_allclose([r.distance for r in results], g["dist"], rtol=1e-6)
        assert results[0].point.dtype == np.dtype(dt) and results[0].point.tolist() == g["points"][g["ids"][0]]
        assert [hnsw.connections(i, 0) for i in range(3)] == g["layer0"]
        hnsw.deinit()


@pytest.mark.parametrize("dtype,odt", [(np.float64, "f64"), (np.int32, "i32")])
def test_typed_index_builds_the_reference_graph_and_finds_the_reference_results(zv, oracle, dtype, odt, tmp_path):
    rng = np.random.default_rng(17)
    n, dim, m = 3000, 24, 8
    if dtype == np.int32:
        X = rng.integers(-1000, 1000, (n, dim)).astype(np.int32)
        Q = rng.integers(-1000, 1000, (100, dim)).astype(np.int32)
    else:
        X = rng.standard_normal((n, dim))          # f64 values that are NOT representable in f32
        Q = rng.standard_normal((100, dim))
    lv = np.minimum(rng.geometric(0.5, n) - 1, 31).astype(np.int32)
    h = zv.HNSW(m, 200, dtype=dtype)
    h.insert_batch(X, levels=lv)
    o = oracle.OracleHNSW(m, 200, dtype=odt)
    o.insert_batch(X, levels=lv)
    for layer in range(0, 3):
        assert np.array_equal(h.export_layer(layer)[0], o.export_layer(layer)[0]), f"layer {layer} differs from the reference's"
    adj, _ = o.export_layer(0)
    ids, dist, counts = h.search_batch(Q, 10, 32)
    ref = oracle.search_graph(X, adj, Q, 32, 10, dtype=odt)                  # the reference's arithmetic in T
    assert np.array_equal(counts, ref["counts"])
    mask = np.arange(10)[None, :] < counts[:, None]
    np.testing.assert_allclose(dist[mask], ref["dist"][mask].astype(np.float64), rtol=1e-5)
    assert ((ids != ref["ids"].astype(np.uint64)) & mask).sum() <= 2         # f32 near-ties only
    assert h.point(5).dtype == np.dtype(dtype) and np.array_equal(h.point(5), X[5])
    # one HNSW(T) holds one T
    import ctypes as C
    other = np.zeros(dim, np.float32)
    rc = zv.lib().zvdb_insert(h._h, other.ctypes.data_as(C.POINTER(C.c_float)), dim)
    assert rc == zv._lib.ERR_INVALID and h.count() == n
    # the typed rows travel through save / load
    path = os.path.join(tmp_path, "typed.zvdb")
    h.save(path)
    b = zv.HNSW(m, 200, dtype=dtype)
    b.load(path)
    assert np.array_equal(b.point(n - 1), X[n - 1]) and np.array_equal(b.export_layer(0)[0], h.export_layer(0)[0])
    assert np.array_equal(b.search_batch(Q, 10, 32)[0], ids)
    b.insert_batch(X[:50]); h.insert_batch(X[:50])
    assert np.array_equal(b.export_layer(0)[0], h.export_layer(0)[0])
    h.deinit(); b.deinit()
