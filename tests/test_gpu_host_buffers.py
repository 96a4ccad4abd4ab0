"""zvdb_search_batch on host buffers: pageable memory is staged through device buffers; page-locked memory
(zvdb_alloc_host) is read and written by the search kernel itself (zero-copy), or -- variant bit 11 -- staged
through the chunked copy pipeline. All three give the same results."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_pinned_and_pageable_buffers_agree(zv, oracle):
    rng = np.random.default_rng(41)
    X = rng.standard_normal((20000, 64), dtype=np.float32)
    nq, k, ef = 6000, 10, 32                      # >= 4096 queries: eligible for the pipeline
    Q = rng.standard_normal((nq, 64), dtype=np.float32)
    h = zv.HNSW(16, 200)
    h.insert_batch(X)
    ids_p, dist_p, cnt_p = h.search_batch(Q, k, ef)             # pageable numpy arrays
    pq, pi, pd, pc = (zv.PinnedArray((nq, 64), np.float32), zv.PinnedArray((nq, k), np.uint64),
                      zv.PinnedArray((nq, k), np.float32), zv.PinnedArray(nq, np.uint32))
    pq.array[:] = Q
    pi.array[:] = 0; pd.array[:] = 0; pc.array[:] = 0
    h.search_batch_ptr(pq.array.ctypes.data, nq, 64, k, ef, pi.array.ctypes.data, pd.array.ctypes.data, pc.array.ctypes.data)
    assert np.array_equal(pi.array, ids_p) and np.array_equal(pd.array.view(np.uint32), dist_p.view(np.uint32))
    assert np.array_equal(pc.array, cnt_p)
    # the staged pipeline on the same page-locked buffers (variant bit 11), with the counters
    pp, pe = zv.PinnedArray(nq, np.uint32), zv.PinnedArray(nq, np.uint32)
    zero_copy = (pi.array.copy(), pd.array.copy(), pc.array.copy())
    for variant in (0x800, 0):
        h.set_kernel_variant(variant)
        pi.array[:] = 0; pd.array[:] = 0; pc.array[:] = 0; pp.array[:] = 0; pe.array[:] = 0
        zv._lib.check(zv.lib().zvdb_search_batch(h._h, pq.array.ctypes.data, nq, 64, k, ef, pi.array.ctypes.data, pd.array.ctypes.data,
                                            pc.array.ctypes.data, pp.array.ctypes.data, pe.array.ctypes.data))
        assert np.array_equal(pi.array, zero_copy[0]) and np.array_equal(pd.array.view(np.uint32), zero_copy[1].view(np.uint32))
        assert np.array_equal(pc.array, zero_copy[2]) and 1 <= pp.array.min() and pp.array.max() <= ef and pe.array.min() >= 1
    adj, _ = h.export_layer(0)
    ref = oracle.search_graph(X, adj, Q, ef, k, dist_mode=oracle.DIST_TREE, heap_mode=oracle.HEAP_DET)
    assert np.array_equal(ids_p, ref["ids"].astype(np.uint64))
    for a in (pq, pi, pd, pc, pp, pe):
        a.free()
    h.deinit()
