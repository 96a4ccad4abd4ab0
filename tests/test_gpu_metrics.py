"""GPU parity tests for the metrics the reference does not have (cosine, dot): K1 and the producer, called
through the C ABI, against the CPU oracle on the same seeded inputs.

The reference's `distance` is squared L2 only (src/hnsw.zig:182-192). BASELINE configs[2] (1M x 768 cosine,
M=32, k=100) and north_star subsystem (2) name cosine and dot, so the ORACLE defines them
(oracle/oracle_impl.h `orc_dist`): cosine = 1 - dot on rows L2-normalised once at insert, dot = -dot, summed
sequentially and unfused like hnsw.zig:186-190 (ORC_DIST_SEQ | metric), or in the kernel's lane order
(ORC_DIST_TREE | metric). Everything else -- insert, connect, shrink, search, the heap -- is the reference's.

Levels of comparison, as in test_gpu_parity.py:
  * producer: the graph zvdb_insert builds under the metric equals the oracle's, layer by layer;
  * bit-exact: ids, order, distance bits, pop/eval counters vs the oracle in the kernel's arithmetic and
    tie order (TREE | metric, HEAP_DET);
  * reference-faithful: vs the oracle in the reference's arithmetic and heap (SEQ | metric, HEAP_ZIG):
    distances within 1e-5 relative (north_star's tolerance), ids identical except at such near-ties.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RTOL = 1e-5   # north_star: "distances within 1e-5 relative"


def _gauss(n, dim, seed):
    return np.random.default_rng(seed).standard_normal((n, dim), dtype=np.float32)


def _metric(zv, oracle, name):
    return {"cos": (zv.METRIC_COSINE, oracle.METRIC_COS), "dot": (zv.METRIC_DOT, oracle.METRIC_DOT),
            "l2": (zv.METRIC_L2, oracle.METRIC_L2)}[name]


def _normalised(X):
    """Rows L2-normalised the way the index does it at insert (host_graph.hpp insert / set_points_locked): squared
    norm summed sequentially in double, one reciprocal square root, product rounded to f32."""
    X64 = X.astype(np.float64)
    s = np.zeros(len(X), np.float64)
    for t in range(X.shape[1]):
        s += X64[:, t] * X64[:, t]
    inv = 1.0 / np.sqrt(s)
    return (X64 * inv[:, None]).astype(np.float32)


def _stored_rows(h):
    """Rows as the index keeps them (normalised once at insert for cosine): what the oracle is given."""
    return np.stack([h.point(i) for i in range(h.count())]).astype(np.float32)


def _faithful(ids, dist, counts, ref, scale):
    """Distances within RTOL of the reference-arithmetic oracle -- relative to the value, or for sums that
    cancel (a dot product near zero) to the magnitude of the summed terms `scale` -- and ids/order identical
    except at near-ties."""
    assert np.array_equal(counts, ref["counts"])
    mask = np.arange(ids.shape[1])[None, :] < counts[:, None]
    np.testing.assert_allclose(dist[mask], ref["dist"][mask], rtol=RTOL, atol=RTOL * scale)
    differ = (ids != ref["ids"].astype(np.uint64)) & mask
    assert differ.sum() <= max(2, 0.002 * mask.sum()), f"{differ.sum()} of {mask.sum()} ids differ"


def _bit_exact(zv, got, ref, k):
    ids, dist, counts, pops, evals = got
    assert np.array_equal(counts, ref["counts"])
    assert np.array_equal(pops, ref["pops"])
    assert np.array_equal(evals, ref["evals"])
    mask = np.arange(k)[None, :] < counts[:, None]
    assert np.array_equal(ids[mask], ref["ids"].astype(np.uint64)[mask])
    assert np.array_equal(dist.view(np.uint32)[mask], ref["dist"].view(np.uint32)[mask])
    assert np.all(ids[~mask] == zv.INVALID_ID)


# ------------------------------------------------------------------------------------------------
# producer parity under the metric: insert / connect / shrink (hnsw.zig:73-170) with the oracle's distance
# ------------------------------------------------------------------------------------------------

@pytest.mark.parametrize("name,n,dim,m", [
    ("cos", 2500, 768, 32),      # C3 shape: 768-d cosine, M = 32
    ("cos", 3000, 24, 4),        # small dim, small m: many shrinks
    ("dot", 2500, 128, 16),
    ("dot", 1500, 200, 8),       # dim not a multiple of 128 floats
])
def test_insert_builds_the_oracle_graph_under_metric(zv, oracle, name, n, dim, m):
    zm, om = _metric(zv, oracle, name)
    X = _gauss(n, dim, 131)
    levels = np.random.default_rng(132).geometric(0.5, n).astype(np.int32) - 1
    h = zv.HNSW(m, 200, metric=zm)
    h.insert_batch(X, levels=levels)
    Xs = _stored_rows(h)
    if name == "cos":           # rows are normalised once, at insert (in double, rounded to f32)
        assert np.array_equal(Xs, _normalised(X))
    else:
        assert np.array_equal(Xs, X)
    o = oracle.OracleHNSW(m, 200, dist_mode=oracle.DIST_SEQ | om)
    o.insert_batch(Xs, levels=levels)
    assert h.count() == o.count() and h.max_level == o.max_level and h.entry_point == o.entry_point == 0
    for layer in range(0, min(o.max_level, 4) + 1):
        ao, do_ = o.export_layer(layer)
        ah, dh = h.export_layer(layer)
        assert np.array_equal(do_, dh), layer
        assert np.array_equal(ao, ah), layer
    h.deinit()


# ------------------------------------------------------------------------------------------------
# K1 under the metric, on the graph the producer built
# ------------------------------------------------------------------------------------------------

@pytest.mark.parametrize("name,n,dim,m,k,ef", [
    ("cos", 3000, 768, 32, 100, 128),    # C3 shape: 6 chunks per lane, M = 32, k = 100, ef >= 128
    ("cos", 3000, 768, 32, 100, 400),    # ... the large-ef visited path
    ("cos", 5000, 32, 16, 10, 64),       # small dim
    ("cos", 4000, 200, 16, 10, 40),      # dim not a multiple of 128 floats
    ("dot", 5000, 128, 16, 10, 64),
    ("dot", 3000, 768, 32, 100, 128),
    ("dot", 4000, 48, 8, 5, 600),
])
def test_search_under_metric_vs_oracle(zv, oracle, name, n, dim, m, k, ef):
    zm, om = _metric(zv, oracle, name)
    X = _gauss(n, dim, 141)
    h = zv.HNSW(m, 200, metric=zm)
    h.insert_batch(X)
    Xs = _stored_rows(h)
    adj, _ = h.export_layer(0)
    Q = _gauss(257, dim, 142)
    if name == "cos":            # embedding-like queries (unit norm) for half the batch, raw for the rest:
        Q[:128] /= np.linalg.norm(Q[:128], axis=1, keepdims=True)   # the kernel computes 1 - <q, row> either way
    got = h.search_batch(Q, k, ef, counters=True)
    ref = oracle.search_graph(Xs, adj, Q, ef, k, dist_mode=oracle.DIST_TREE | om, heap_mode=oracle.HEAP_DET)
    _bit_exact(zv, got, ref, k)
    faithful = oracle.search_graph(Xs, adj, Q, ef, k, dist_mode=oracle.DIST_SEQ | om, heap_mode=oracle.HEAP_ZIG)
    scale = float(np.abs(Q).mean() * np.abs(Xs).mean() * dim) if name == "dot" else 1.0   # ~ sum_i |q_i x_i|
    _faithful(got[0], got[1], got[2], faithful, scale)
    h.deinit()


@pytest.mark.parametrize("name", ["cos", "dot"])
def test_kernel_variants_agree_under_metric(zv, oracle, name):
    zm, om = _metric(zv, oracle, name)
    X = _gauss(6000, 128, 143)
    h = zv.HNSW(16, 200, metric=zm)
    h.insert_batch(X)
    Q = _gauss(200, 128, 144)
    base = h.search_batch(Q, 10, 96, counters=True)
    for variant in (0b0100, 0b1000, 0b1100):        # shared-memory hash, global bitmap, global hash
        h.set_kernel_variant(variant)
        got = h.search_batch(Q, 10, 96, counters=True)
        for a, b in zip(base, got):
            assert np.array_equal(a.view(np.uint8), b.view(np.uint8)), variant
    h.deinit()


# ------------------------------------------------------------------------------------------------
# K1 under the metric on a builder graph (the throughput track of C3), and the builder itself
# ------------------------------------------------------------------------------------------------

@pytest.mark.parametrize("name,n,dim,m,K,k,ef", [
    ("cos", 1500, 768, 32, 64, 100, 128),   # C3 shape
    ("cos", 2000, 64, 16, 48, 10, 64),
    ("dot", 2000, 96, 16, 48, 10, 64),
])
def test_builder_graph_under_metric(zv, oracle, name, n, dim, m, K, k, ef):
    from builder_ref import build_ref
    zm, om = _metric(zv, oracle, name)
    X = _gauss(n, dim, 151)
    if name == "dot":
        X *= np.random.default_rng(152).uniform(0.5, 1.5, (n, 1)).astype(np.float32)   # row norms matter for dot
    h = zv.HNSW(m, 200, metric=zm)
    # candidate ids from the oracle's exact search under the same metric (self at rank 0 for cosine; tolerated)
    Xn = X if name == "dot" else _normalised(X)
    nn, _ = oracle.bruteforce(Xn, Xn, K, metric=zm)
    h.build_from_candidates(X, nn.astype(np.uint32))
    Xs = _stored_rows(h)
    assert np.array_equal(Xs, Xn)
    adj, deg = h.export_layer(0)
    assert np.array_equal(adj, build_ref(oracle, Xs, nn.astype(np.uint32), m, metric=om))
    Q = _gauss(129, dim, 153)
    got = h.search_batch(Q, k, ef, counters=True)
    ref = oracle.search_graph(Xs, adj, Q, ef, k, dist_mode=oracle.DIST_TREE | om, heap_mode=oracle.HEAP_DET)
    _bit_exact(zv, got, ref, k)
    faithful = oracle.search_graph(Xs, adj, Q, ef, k, dist_mode=oracle.DIST_SEQ | om, heap_mode=oracle.HEAP_ZIG)
    scale = float(np.abs(Q).mean() * np.abs(Xs).mean() * dim) if name == "dot" else 1.0   # ~ sum_i |q_i x_i|
    _faithful(got[0], got[1], got[2], faithful, scale)
    if name == "cos":           # and it is a useful graph: recall against the exact cosine neighbours
        gt, _ = oracle.bruteforce(Xs, Q, 10, metric=zm)
        ids = h.search_batch(Q, 10, 200)[0]
        rec = np.mean([len(set(ids[i].tolist()) & set(gt[i].tolist())) / 10 for i in range(len(Q))])
        assert rec > 0.85, rec
    h.deinit()


def test_metric_is_part_of_the_index_not_the_call(zv):
    """An L2 index and a cosine index over the same rows disagree (sanity: the METRIC template is live)."""
    X, Q = _gauss(3000, 64, 161), _gauss(50, 64, 162)
    Q /= np.linalg.norm(Q, axis=1, keepdims=True)
    a = zv.HNSW(16, 200); a.insert_batch(X)
    b = zv.HNSW(16, 200, metric=zv.METRIC_COSINE); b.insert_batch(X)
    da, db = a.search_batch(Q, 5, 20)[1], b.search_batch(Q, 5, 20)[1]
    assert not np.array_equal(da, db)
    assert np.all(db >= -1e-6) and np.all(db <= 2.0 + 1e-6)      # 1 - cos of unit vectors
    a.deinit(); b.deinit()
