#!/usr/bin/env python
"""One C4 shard (12.5M x 128 rows = shard 0 of the 100M-row index over 8 GPUs) on one GPU: what the upper-layer descent (K2)
buys at the small per-shard pop budgets of the sharded runs. Incremental quality graph on layer 0, hierarchy from
builder.build_hierarchy with levels drawn geometrically at p = 1/16 (one node in 16 reaches layer 1, one in 256 layer 2, ...),
K4 ground truth inside the shard, pop budgets 10..512 with the descent off / on. One JSON line per point."""
import json, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import bench, zvdb_b200
from zvdb_b200 import builder

rows_total = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
shards, dim, m, nq, k = 8, 128, 16, 10_000, 10
X = bench.make_shard(rows_total, dim, 0, shards)
n = len(X)
Q = np.random.default_rng(2).standard_normal((nq, dim), dtype=np.float32)
dev = torch.device("cuda", 0)
stream = torch.cuda.current_stream().cuda_stream
h = zvdb_b200.HNSW(m, 200)
t0 = time.time()
builder.build_quality_graph_incremental(h, X, m)
t_graph = time.time() - t0
t0 = time.time()
levels = builder.draw_levels(n, seed=3, p=1.0 / 16)
builder.build_hierarchy(h, X, m, levels=levels)
t_hier = time.time() - t0
h.sync_device()
print(f"# {n} rows: layer 0 in {t_graph:.1f}s, hierarchy (max level {int(levels.max())}, {int((levels > 0).sum())} nodes above layer 0) in {t_hier:.1f}s", flush=True)
dq = torch.from_numpy(Q).to(dev)
ids = torch.empty((nq, k), dtype=torch.int64, device=dev); dist = torch.empty((nq, k), dtype=torch.float32, device=dev)
cnt = torch.empty(nq, dtype=torch.int32, device=dev); pops = torch.empty(nq, dtype=torch.int32, device=dev); evals = torch.empty(nq, dtype=torch.int32, device=dev)
h.bruteforce_knn_device(dq.data_ptr(), nq, k, ids.data_ptr(), dist.data_ptr(), cnt.data_ptr(), stream=stream)
torch.cuda.synchronize()
gt = ids.cpu().numpy().copy()
peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]
for descent in (False, True):
    h.set_descent(descent)
    for ef in (10, 16, 32, 64, 128, 256, 512):
        ms = []
        for r in range(4):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            h.search_batch_device(dq.data_ptr(), nq, k, ef, ids.data_ptr(), dist.data_ptr(), cnt.data_ptr(), pops.data_ptr(), evals.data_ptr(), stream=stream)
            b.record(); torch.cuda.synchronize()
            if r: ms.append(a.elapsed_time(b))
        t = float(np.median(ms))
        got = ids.cpu().numpy()
        rec = float(np.mean([len(set(got[i].tolist()) & set(gt[i].tolist())) / k for i in range(nq)]))
        ev = float(evals.cpu().numpy().astype(np.int64).mean())
        print(json.dumps({"rows": n, "m": m, "descent": descent, "pops": ef, "ms": round(t, 4), "qps": round(nq / t * 1e3), "recall_at_10_in_shard": round(rec, 4),
                          "evals_per_query": round(ev, 1), "frac_of_hbm_peak": round((ev * 512 + ef * m * 4 + 632) * nq / (t * 1e-3) / 1e9 / peak, 3)}), flush=True)
h.deinit()
