"""CPU tests (gloo, world_size 2) of the id-sharded path's HOST logic: row partition, local/global
id mapping, packed block layout, gather order and merge order. The shard-local compute
is injected: here the CPU oracle stands in for the CUDA library (the product has no CPU backend),
so what is under test is zvdb_b200/sharded.py and the collective plumbing, not a kernel."""
import os
import sys
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleBackend:
    """Test double for CudaBackend: same hooks, CPU oracle underneath, torch CPU tensors."""

    def __init__(self, O, m, rank, world):
        self.O, self.rank, self.world = O, rank, world
        self.ix = O.OracleHNSW(m, 200)

    def insert_batch(self, pts):
        self.ix.insert_batch(pts)

    def to_device(self, q):
        import torch
        return torch.from_numpy(np.ascontiguousarray(q, np.float32))

    def to_host(self, ids, dist, counts):
        return ids.numpy().view(np.uint64), dist.numpy(), counts.numpy().view(np.uint32)

    def search_packed(self, d_queries, nq, k, ef):
        import torch
        from zvdb_b200.sharded import pack_block
        O = self.O
        ids = np.full((nq, k), 0xFFFFFFFFFFFFFFFF, np.uint64)
        dist = np.zeros((nq, k), np.float32)
        counts = np.zeros(nq, np.uint32)
        if self.ix.count():
            adj, _ = self.ix.export_layer(0)
            r = O.search_graph(self.ix.points(), adj, d_queries.numpy(), ef, k, dist_mode=O.DIST_TREE, heap_mode=O.HEAP_DET)
            counts = r["counts"]
            mask = np.arange(k)[None, :] < counts[:, None]
            ids[mask] = r["ids"].astype(np.uint64)[mask] * self.world + self.rank
            dist[mask] = r["dist"][mask]
        return torch.from_numpy(pack_block(ids, dist, counts))

    def merge_packed(self, gathered, nq, k):
        import torch
        from zvdb_b200.sharded import block_bytes, unpack_block
        bb = block_bytes(nq, k)
        g = gathered.numpy()
        parts = [unpack_block(g[i * bb:(i + 1) * bb], nq, k) for i in range(self.world)]
        d, i, c = self.O.merge_topk(np.stack([p[1] for p in parts]), np.stack([p[0] for p in parts]),
                                    np.stack([p[2] for p in parts]))
        return torch.from_numpy(i.view(np.int64)), torch.from_numpy(d), torch.from_numpy(c.view(np.int32))


def _worker(rank, world, port, outdir, n, dim, m, nq, k, ef):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from oracle import oracle as O
    from zvdb_b200.sharded import ShardedHNSW
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    X = np.random.default_rng(91).standard_normal((n, dim), dtype=np.float32)
    Q = np.random.default_rng(92).standard_normal((nq, dim), dtype=np.float32)
    sh = ShardedHNSW(m, 200, exchange="nccl", backend=OracleBackend(O, m, rank, world))
    assert (sh.rank, sh.world) == (rank, world)
    sh.insert_batch(X[:7])           # ragged batches: ownership must follow the global row counter
    sh.insert_batch(X[7:1000])
    sh.insert_batch(X[1000:])
    assert sh.count() == n and sh.backend.ix.count() == len(range(rank, n, world))
    assert np.array_equal(sh.backend.ix.points(), X[rank::world])
    ids, d, c = sh.search_batch(Q, k, ef)
    np.savez(os.path.join(outdir, f"r{rank}.npz"), ids=ids, dist=d, counts=c)
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("n,dim,m,nq,k,ef", [(3001, 32, 8, 65, 10, 40), (500, 16, 4, 9, 5, 5)])
def test_two_rank_sharded_search_matches_the_sharded_oracle(oracle, n, dim, m, nq, k, ef):
    import torch.multiprocessing as mp
    from zvdb_b200.sharded import per_shard_ef
    world = 2
    with tempfile.TemporaryDirectory() as tmp:
        mp.spawn(_worker, args=(world, _free_port(), tmp, n, dim, m, nq, k, ef), nprocs=world, join=True)
        got = [np.load(os.path.join(tmp, f"r{r}.npz")) for r in range(world)]
    # oracle for this path: G independent reference indexes + the same merge
    O = oracle
    X = np.random.default_rng(91).standard_normal((n, dim), dtype=np.float32)
    Q = np.random.default_rng(92).standard_normal((nq, dim), dtype=np.float32)
    e = per_shard_ef(ef, k, world)
    D, I, Cn = [], [], []
    for r in range(world):
        ix = O.OracleHNSW(m, 200)
        ix.insert_batch(X[r::world])
        adj, _ = ix.export_layer(0)
        res = O.search_graph(X[r::world], adj, Q, e, k, dist_mode=O.DIST_TREE, heap_mode=O.HEAP_DET)
        ids = np.full((nq, k), 0xFFFFFFFFFFFFFFFF, np.uint64)
        mask = np.arange(k)[None, :] < res["counts"][:, None]
        ids[mask] = res["ids"].astype(np.uint64)[mask] * world + r
        D.append(res["dist"]); I.append(ids); Cn.append(res["counts"])
    d, i, c = O.merge_topk(np.stack(D), np.stack(I), np.stack(Cn))
    for r in range(world):               # every rank holds the merged result (all-gather semantics)
        assert np.array_equal(got[r]["counts"], c)
        assert np.array_equal(got[r]["ids"], i)
        assert np.array_equal(got[r]["dist"].view(np.uint32), d.view(np.uint32))
    # global ids map back to rows: distance of (query, X[global id]) is the reported one
    q0 = 0
    dd = O.dist_many(Q[q0], X, i[q0, :c[q0]].astype(np.uint32), O.DIST_TREE)
    assert np.array_equal(dd.view(np.uint32), d[q0, :c[q0]].view(np.uint32))


def test_block_layout_round_trip():
    from zvdb_b200.sharded import block_bytes, pack_block, unpack_block, shard_rows, per_shard_ef
    rng = np.random.default_rng(3)
    for nq, k in ((1, 1), (7, 10), (300, 100)):
        ids = rng.integers(0, 2**63, (nq, k)).astype(np.uint64)
        dist = rng.random((nq, k), dtype=np.float32)
        cnt = rng.integers(0, k + 1, nq).astype(np.uint32)
        b = pack_block(ids, dist, cnt)
        assert len(b) == block_bytes(nq, k) and len(b) % 256 == 0
        i2, d2, c2 = unpack_block(b, nq, k)
        assert np.array_equal(i2, ids) and np.array_equal(d2, dist) and np.array_equal(c2, cnt)
    assert shard_rows(10, 1, 4).tolist() == [1, 5, 9]
    assert per_shard_ef(64, 10, 8) == 10 and per_shard_ef(64, 10, 2) == 32 and per_shard_ef(10, 10, 1) == 10


def test_block_bytes_matches_the_library(zv):
    from zvdb_b200.sharded import block_bytes
    for nq, k in ((1, 1), (10000, 10), (333, 77)):
        assert zv.lib().zvdb_shard_block_bytes(nq, k) == block_bytes(nq, k)
