#!/usr/bin/env python
"""Recall/QPS exploration of the quality track on one B200 (not part of the product or tests):
graph degree m and candidate count K vs recall@10 at several ef, 1M x 128 Gaussian."""
import json, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import zvdb_b200
from zvdb_b200 import builder

n, dim, nq, k = 1_000_000, 128, 10_000, 10
X = np.random.default_rng(1).standard_normal((n, dim), dtype=np.float32)
Q = np.random.default_rng(2).standard_normal((nq, dim), dtype=np.float32)
dev = torch.device("cuda", 0)
dq = torch.from_numpy(Q).to(dev)
d_ids = torch.empty((nq, k), dtype=torch.int64, device=dev)
d_dist = torch.empty((nq, k), dtype=torch.float32, device=dev)
d_cnt = torch.empty(nq, dtype=torch.int32, device=dev)
d_ev = torch.empty(nq, dtype=torch.int32, device=dev)
stream = torch.cuda.current_stream().cuda_stream
gt = None
for m, K in ((16, 64), (32, 64), (32, 128), (64, 128)):
    h = zvdb_b200.HNSW(m, 200)
    t0 = time.time()
    builder.build_quality_graph(h, X, m, K=K)
    h.sync_device()
    bt = time.time() - t0
    if gt is None:
        g_ids = torch.empty((nq, k), dtype=torch.int64, device=dev)
        h.bruteforce_knn_device(dq.data_ptr(), nq, k, g_ids.data_ptr(), d_dist.data_ptr(), d_cnt.data_ptr(), stream=stream)
        torch.cuda.synchronize()
        gt = g_ids.cpu().numpy()
    adj, deg = h.export_layer(0)
    for ef in (32, 64, 128, 256, 512, 1024):
        try:
            for _ in range(2):
                h.search_batch_device(dq.data_ptr(), nq, k, ef, d_ids.data_ptr(), d_dist.data_ptr(), d_cnt.data_ptr(), 0, d_ev.data_ptr(), stream=stream)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(2):
                h.search_batch_device(dq.data_ptr(), nq, k, ef, d_ids.data_ptr(), d_dist.data_ptr(), d_cnt.data_ptr(), 0, d_ev.data_ptr(), stream=stream)
            b.record(); torch.cuda.synchronize()
            ms = a.elapsed_time(b) / 2
            got = d_ids.cpu().numpy()
            rec = np.mean([len(set(got[i].tolist()) & set(gt[i].tolist())) / k for i in range(nq)])
            ev = float(d_ev.cpu().numpy().mean())
            print(json.dumps({"m": m, "K": K, "build_s": round(bt, 1), "mean_deg": round(float(deg.mean()), 2), "ef": ef, "ms": round(ms, 3),
                              "qps": round(nq / ms * 1e3), "recall": round(float(rec), 4), "evals": round(ev, 1),
                              "GBps": round(ev * 512 * nq / ms / 1e6, 1)}), flush=True)
        except zvdb_b200.ZvdbError as e:
            print(json.dumps({"m": m, "ef": ef, "error": str(e)[:100]}), flush=True)
    h.deinit()
