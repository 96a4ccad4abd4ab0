import sys, numpy as np
sys.path.insert(0, ".")
import zvdb_b200
from zvdb_b200.sharded import ShardedHNSW
from oracle import oracle as O
O.build()
X = np.random.default_rng(111).standard_normal((5000, 32), dtype=np.float32)
Q = np.random.default_rng(112).standard_normal((12000, 32), dtype=np.float32)
for vis in (0, 8, 12):
    sh = ShardedHNSW(16, 200, rank=0, world=1, device=0, exchange="p2p")
    if vis: sh.index.set_kernel_variant(vis)
    sh.insert_batch(X)
    adj, _ = sh.index.export_layer(0)
    for k, ef in ((100, 100), (100, 128), (10, 100), (33, 100), (64, 64)):
        ref = O.search_graph(X, adj, Q, ef, k, dist_mode=O.DIST_TREE, heap_mode=O.HEAP_DET)
        want = sh.index.search_batch(Q, k, ef)
        mask = np.arange(k)[None, :] < ref["counts"][:, None]
        plain_ok = np.array_equal(want[0][mask], ref["ids"].astype(np.uint64)[mask]) and np.array_equal(want[2], ref["counts"])
        got = sh.search_batch(Q, k, ef)
        bad = np.nonzero((got[0] != want[0]).any(axis=1) | (got[2] != want[2]))[0]
        print(f"vis={vis} k={k} ef={ef}: plain==oracle {plain_ok}; fused!=plain rows {len(bad)} first {bad[:8].tolist()}", flush=True)
        if len(bad):
            q = bad[0]
            print("   counts", got[2][q], want[2][q], "ids got", got[0][q][:12].tolist(), "want", want[0][q][:12].tolist())
            print("   dist got", got[1][q][:6].tolist(), "want", want[1][q][:6].tolist())
    sh.deinit()
