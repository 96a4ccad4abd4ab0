"""The C oracle (oracle/) against a second, independently written restatement of src/hnsw.zig in plain Python
(tests/pyref_hnsw.py): graph, search results (ids and distance BITS), pop and evaluation counters must be equal.
The reference itself cannot be run in this image (Zig), so two restatements that agree are the strongest check of
the restatement that can be had here; what neither can pin is the Zig standard library's own behaviour on exact ties."""
import numpy as np
import pytest

from pyref_hnsw import PyHNSW


def _levels(n, rng):
    lv = np.zeros(n, np.int32)
    for i in range(n):                                          # hnsw.zig:172-180: geometric, p = 1/2, cap 31
        while lv[i] < 31 and rng.random() < 0.5:
            lv[i] += 1
    return lv


@pytest.mark.parametrize("n,dim,m,seed", [(300, 6, 4, 1), (250, 16, 8, 2), (120, 3, 16, 3)])
def test_c_oracle_equals_python_restatement(oracle, n, dim, m, seed):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n, dim)).astype(np.float32)
    lv = _levels(n, rng)
    py = PyHNSW(m)
    for p, l in zip(X, lv):
        py.insert(p, int(l))
    o = oracle.OracleHNSW(m, 200)
    o.insert_batch(X, levels=lv)
    assert o.max_level == py.max_level == int(lv.max()) and o.entry_point == py.entry_point == 0
    for layer in range(py.max_level + 1):
        assert np.array_equal(o.export_layer(layer)[0], py.layer(layer)), f"layer {layer} differs"
    Q = rng.standard_normal((40, dim)).astype(np.float32)
    for q in Q:
        for k in (1, 7, 40, n + 5):
            ids, d, pops, evals = o.search(q, k, counters=True)
            pids, pd, ppops, pevals = py.search(q, k)
            assert np.array_equal(ids, pids) and np.array_equal(d.view(np.uint32), pd.view(np.uint32))
            assert (pops, evals) == (ppops, pevals)


def test_c_oracle_equals_python_restatement_on_exact_ties(oracle):
    """Duplicate points and a tiny integer lattice: many exact distance ties, so the heap's structural tie order and
    the stability of both sorts are exercised (test_hnsw.zig:104-126 is the reference's own duplicate case)."""
    rng = np.random.default_rng(7)
    X = rng.integers(0, 3, size=(200, 4)).astype(np.float32)     # 81 distinct points among 200
    lv = _levels(len(X), rng)
    py = PyHNSW(5)
    for p, l in zip(X, lv):
        py.insert(p, int(l))
    o = oracle.OracleHNSW(5, 200)
    o.insert_batch(X, levels=lv)
    for layer in range(py.max_level + 1):
        assert np.array_equal(o.export_layer(layer)[0], py.layer(layer))
    for q in rng.integers(0, 3, size=(30, 4)).astype(np.float32):
        for k in (3, 25, 300):
            ids, d, pops, evals = o.search(q, k, heap_mode=oracle.HEAP_ZIG, counters=True)
            pids, pd, ppops, pevals = py.search(q, k)
            assert np.array_equal(ids, pids) and np.array_equal(d.view(np.uint32), pd.view(np.uint32))
            assert (pops, evals) == (ppops, pevals)


def test_c_oracle_descent_equals_python_restatement(oracle):
    rng = np.random.default_rng(11)
    n, dim, m = 400, 8, 6
    X = rng.standard_normal((n, dim)).astype(np.float32)
    lv = _levels(n, rng)
    py = PyHNSW(m)
    for p, l in zip(X, lv):
        py.insert(p, int(l))
    o = oracle.OracleHNSW(m, 200)
    o.insert_batch(X, levels=lv)
    upper = o.export_upper()
    assert upper[3] == py.max_level >= 2
    for q in rng.standard_normal((60, dim)).astype(np.float32):
        node, d, ev = oracle.descend_one(X, upper, q)
        pnode, pd, pev = py.descend(q)
        assert (node, ev) == (pnode, pev) and np.float32(d).view(np.uint32) == np.float32(pd).view(np.uint32)


@pytest.mark.parametrize("dtype,npdt", [("f64", np.float64), ("i32", np.int32)])
def test_c_oracle_equals_python_restatement_for_other_element_types(oracle, dtype, npdt):
    """HNSW(f64) and HNSW(i32) (test_hnsw.zig:239-273): distances summed in T's own arithmetic."""
    rng = np.random.default_rng(13)
    n, dim, m = 200, 5, 4
    X = (rng.standard_normal((n, dim)) * 3).astype(npdt) if dtype == "f64" else rng.integers(-20, 20, size=(n, dim)).astype(npdt)
    lv = _levels(n, rng)
    py = PyHNSW(m, npdt)
    for p, l in zip(X, lv):
        py.insert(p, int(l))
    o = oracle.OracleHNSW(m, 200, dtype=dtype)
    o.insert_batch(X, levels=lv)
    for layer in range(py.max_level + 1):
        assert np.array_equal(o.export_layer(layer)[0], py.layer(layer))
    Q = (rng.standard_normal((25, dim)) * 3).astype(npdt) if dtype == "f64" else rng.integers(-20, 20, size=(25, dim)).astype(npdt)
    for q in Q:
        for k in (2, 30):
            ids, d, pops, evals = o.search(q, k, counters=True)
            pids, pd, ppops, pevals = py.search(q, k)
            assert np.array_equal(ids, pids) and np.array_equal(d, pd) and (pops, evals) == (ppops, pevals)


def test_python_restatement_reproduces_the_hand_worked_heap_tie_golden():
    """G11 (tests/golden/reference_tests.json): eight exact ties, pop order derived by hand from Zig's
    PriorityQueue sift-down rule; the second restatement must land on the same order."""
    import json
    import os
    G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_tests.json")))
    for name in ("G11_heap_tie_rule_k9", "G11_heap_tie_rule_k5"):
        g = G[name]
        py = PyHNSW(g["m"])
        for p in g["points"]:
            py.insert(np.asarray(p, np.float32), 0)
        assert [[int(x) for x in row if x != 0xFFFFFFFF] for row in py.layer(0)] == g["layer0"]
        ids, d, pops, evals = py.search(np.asarray(g["query"], np.float32), g["k"])
        assert [int(x) for x in ids] == g["ids"] and [float(x) for x in d] == g["dist"]
        assert (pops, evals) == (g["pops"], g["evals"])
