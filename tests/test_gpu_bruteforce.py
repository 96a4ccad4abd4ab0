"""GPU parity tests for K4, the exact brute-force k-NN (tcgen05 3xTF32 GEMM + fused top-k + exact
re-rank), called through the C ABI. No reference counterpart exists: the oracle
is orc_bruteforce_f32 (distances accumulated in double), and the distance bits are additionally
pinned to the search kernel's own arithmetic (oracle ORC_DIST_TREE)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RTOL = 1e-5   # north_star: distances within 1e-5 relative; ids may differ only at such near-ties


def _gauss(n, dim, seed):
    return np.random.default_rng(seed).standard_normal((n, dim), dtype=np.float32)


def _check_against_oracle(oracle, X, Q, k, ids, dist, counts, metric=0, Xn=None):
    n = len(X)
    kk = min(k, n)
    assert np.all(counts == kk)
    ref_ids, ref_d = oracle.bruteforce(X if Xn is None else Xn, Q, kk, metric=metric)
    np.testing.assert_allclose(dist[:, :kk], ref_d, rtol=RTOL, atol=2e-6)
    differ = ids[:, :kk] != ref_ids.astype(np.uint64)
    # an id may differ only where the two candidates' distances tie within RTOL (checked by the
    # allclose above, position by position); such positions must be rare
    assert differ.sum() <= max(2, 0.002 * differ.size), f"{differ.sum()} of {differ.size} ids differ"
    assert np.all(ids[:, kk:] == 0xFFFFFFFFFFFFFFFF)
    assert np.all(np.diff(dist[:, :kk], axis=1) >= 0)


BF_SINGLE, BF_PAIR = 1 << 4, 2 << 4   # zvdb_set_kernel_variant bits 4-5: cta_group::1 / cta_group::2 GEMM
BF_SORTED = 1 << 7                    # bit 7: sorted lists + cooperative insertion instead of append-and-compact


@pytest.mark.parametrize("shape", [BF_SINGLE, BF_PAIR, BF_PAIR | BF_SORTED])
@pytest.mark.parametrize("n,dim,nq,k", [
    (5000, 128, 300, 10),      # several row tiles, 3 query tiles (last one ragged)
    (1000, 3, 17, 5),          # tiny dim: one K chunk, zero padding
    (3000, 200, 129, 100),     # dim not a multiple of 32; k = 100 -> per-thread lists in global memory
    (50, 16, 40, 100),         # k > n
    (129, 64, 1, 1),           # single query, k = 1, ragged last row tile
    (20000, 128, 1000, 10),    # more work items than one wave of splits
    (70000, 32, 5000, 10),     # whole waves plus a refined last wave (several segments per CTA)
    (30000, 96, 700, 300),     # k + slack > 256: few result slots, several refined items per query tile
])
def test_bruteforce_matches_oracle(zv, oracle, n, dim, nq, k, shape):
    X, Q = _gauss(n, dim, 71), _gauss(nq, dim, 72)
    h = zv.HNSW(16, 200)
    h.set_kernel_variant(shape)
    h.insert_batch(X)
    ids, dist, counts = h.bruteforce_knn(Q, k)
    _check_against_oracle(oracle, X, Q, k, ids, dist, counts)
    # distance bits are the search kernel's own (difference form, lane order + butterfly)
    for qi in (0, nq - 1):
        c = int(counts[qi])
        d = oracle.dist_many(Q[qi], X, ids[qi, :c].astype(np.uint32), oracle.DIST_TREE)
        assert np.array_equal(d.view(np.uint32), dist[qi, :c].view(np.uint32))
    h.deinit()


@pytest.mark.parametrize("metric_name,dim,m", [("cos", 768, 32), ("dot", 96, 16)])
def test_bruteforce_other_metrics(zv, oracle, metric_name, dim, m):
    metric = {"cos": zv.METRIC_COSINE, "dot": zv.METRIC_DOT}[metric_name]
    X, Q = _gauss(2500, dim, 73), _gauss(200, dim, 74)
    h = zv.HNSW(m, 200, metric=metric)
    h.insert_batch(X)
    Xs = np.stack([h.point(i) for i in range(len(X))])   # rows as stored (normalised for cosine)
    ids, dist, counts = h.bruteforce_knn(Q, 20)
    _check_against_oracle(oracle, X, Q, 20, ids, dist, counts, metric=metric, Xn=Xs)
    h.deinit()


def test_bruteforce_agrees_with_search_distances(zv):
    """Both kernels return bit-identical distances for the same (query, id)."""
    X, Q = _gauss(8000, 128, 75), _gauss(64, 128, 76)
    h = zv.HNSW(16, 200)
    h.insert_batch(X)
    sid, sdist, scnt = h.search_batch(Q, 10, 64)
    bid, bdist, bcnt = h.bruteforce_knn(Q, 8000 if False else 64)
    hits = 0
    for q in range(len(Q)):
        lut = {int(i): float(d) for i, d in zip(bid[q], bdist[q])}
        for i, d in zip(sid[q, :scnt[q]], sdist[q, :scnt[q]]):
            if int(i) in lut:
                assert np.float32(lut[int(i)]).view(np.uint32) == np.float32(d).view(np.uint32)
                hits += 1
    h.deinit()


def test_bruteforce_pair_and_single_cta_agree_bit_for_bit(zv):
    X, Q = _gauss(30000, 96, 81), _gauss(700, 96, 82)
    h = zv.HNSW(16, 200)
    h.insert_batch(X)
    for k in (25, 1, 60, 120, 300):             # append-and-compact with 1, 1, 4, 8 keys per lane; k = 300 -> sorted lists
        h.set_kernel_variant(BF_SINGLE)
        a = h.bruteforce_knn(Q, k)
        for variant in (BF_PAIR, BF_PAIR | BF_SORTED):
            h.set_kernel_variant(variant)
            b = h.bruteforce_knn(Q, k)
            assert np.array_equal(a[0], b[0]) and np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32)) and np.array_equal(a[2], b[2])
    h.deinit()


def test_bruteforce_filter_mode_is_near_exact_and_returns_exact_distances(zv, oracle):
    """Bit 6 of the variant: single-product TF32 GEMM as a filter + exact re-rank of k+24 candidates."""
    X, Q = _gauss(60000, 128, 83), _gauss(500, 128, 84)
    h = zv.HNSW(16, 200)
    h.insert_batch(X)
    exact = h.bruteforce_knn(Q, 10)
    h.set_kernel_variant(1 << 6)
    ids, dist, counts = h.bruteforce_knn(Q, 10)
    recall = np.mean([len(set(ids[i].tolist()) & set(exact[0][i].tolist())) / 10 for i in range(len(Q))])
    assert recall >= 0.999, recall
    for qi in (0, 250, 499):                     # whatever is returned carries the exact distance
        d = oracle.dist_many(Q[qi], X, ids[qi].astype(np.uint32), oracle.DIST_TREE)
        assert np.array_equal(d.view(np.uint32), dist[qi].view(np.uint32))
    assert np.all(np.diff(dist, axis=1) >= 0) and np.all(counts == 10)
    h.set_kernel_variant(0)
    again = h.bruteforce_knn(Q, 10)
    assert np.array_equal(again[0], exact[0])
    h.deinit()


def test_bruteforce_follows_inserts_and_empty_index(zv, oracle):
    h = zv.HNSW(16, 200)
    ids, dist, counts = h.bruteforce_knn(np.zeros((3, 8), np.float32), 4)
    assert np.all(counts == 0) and np.all(ids == zv.INVALID_ID)
    X = _gauss(700, 8, 77)
    Q = _gauss(9, 8, 78)
    h.insert_batch(X[:300])
    ids, dist, counts = h.bruteforce_knn(Q, 4)
    _check_against_oracle(oracle, X[:300], Q, 4, ids, dist, counts)
    h.insert_batch(X[300:])                      # the operand split must be refreshed
    ids, dist, counts = h.bruteforce_knn(Q, 4)
    _check_against_oracle(oracle, X, Q, 4, ids, dist, counts)
    with pytest.raises(zv.ZvdbError):
        h.bruteforce_knn(np.zeros((1, 9), np.float32), 4)
    h.deinit()


def test_bruteforce_device_entry_point(zv, oracle):
    import torch
    X, Q = _gauss(4000, 64, 79), _gauss(130, 64, 80)
    h = zv.HNSW(16, 200)
    h.insert_batch(X)
    ids, dist, counts = h.bruteforce_knn(Q, 10)
    dq = torch.from_numpy(Q).cuda()
    d_ids = torch.empty((130, 10), dtype=torch.int64, device="cuda")
    d_dist = torch.empty((130, 10), dtype=torch.float32, device="cuda")
    d_cnt = torch.empty(130, dtype=torch.int32, device="cuda")
    h.bruteforce_knn_device(dq.data_ptr(), 130, 10, d_ids.data_ptr(), d_dist.data_ptr(), d_cnt.data_ptr(),
                            id_stride=4, id_base=1, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert np.array_equal(d_ids.cpu().numpy().astype(np.uint64), ids * 4 + 1)
    assert np.array_equal(d_dist.cpu().numpy(), dist)
    h.deinit()
