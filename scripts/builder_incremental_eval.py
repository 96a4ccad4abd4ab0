#!/usr/bin/env python
"""Search-driven incremental builder vs the exact-candidate (GEMM) builder on one B200: build seconds, mean degree,
recall@10 / QPS over an ef sweep against K4's exact ground truth. JSON lines. Not product code.
    python scripts/builder_incremental_eval.py [rows=1000000] [skip_exact=0]"""
import json, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import zvdb_b200
from zvdb_b200 import builder

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
skip_exact = len(sys.argv) > 2 and sys.argv[2] == "1"
dim, nq, k, m = 128, 10_000, 10, 16
X = np.random.default_rng(1).standard_normal((n, dim), dtype=np.float32)
Q = np.random.default_rng(2).standard_normal((nq, dim), dtype=np.float32)
dev = torch.device("cuda", 0)
dq = torch.from_numpy(Q).to(dev)
mk = lambda dt, *s: torch.empty(s, dtype=dt, device=dev)
d_ids, d_dist, d_cnt, d_pops, d_evals = mk(torch.int64, nq, k), mk(torch.float32, nq, k), mk(torch.int32, nq), mk(torch.int32, nq), mk(torch.int32, nq)
stream = torch.cuda.current_stream().cuda_stream
gt = None
configs = [("exact", {})] if not skip_exact else []
import os
build_efs = [int(x) for x in os.environ.get("ZVDB_BUILD_EF", "128,256,512").split(",")]
build_K = int(os.environ.get("ZVDB_BUILD_K", "64"))
configs += [("incremental", dict(refine_rounds=2, ef=e, K=build_K)) for e in build_efs]
for name, kw in configs:
    h = zvdb_b200.HNSW(m, 200)
    t0 = time.time()
    if name == "exact":
        builder.build_quality_graph(h, X, m)
        stats = {}
    else:
        stats = builder.build_quality_graph_incremental(h, X, m, log=lambda s: print("#", s, flush=True), **kw)
    h.sync_device()
    build_s = time.time() - t0
    if gt is None:      # exact ground truth from the tensor-core brute force (K4)
        g_ids = mk(torch.int64, nq, k); g_dist = mk(torch.float32, nq, k); g_cnt = mk(torch.int32, nq)
        h.bruteforce_knn_device(dq.data_ptr(), nq, k, g_ids.data_ptr(), g_dist.data_ptr(), g_cnt.data_ptr(), stream=stream)
        torch.cuda.synchronize()
        gt = g_ids.cpu().numpy()
    adj, deg = h.export_layer(0)
    for ef in (64, 128, 256, 512):
        run = lambda: h.search_batch_device(dq.data_ptr(), nq, k, ef, d_ids.data_ptr(), d_dist.data_ptr(), d_cnt.data_ptr(),
                                            d_pops.data_ptr(), d_evals.data_ptr(), stream=stream)
        for _ in range(2): run()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3): run()
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 3
        ids = d_ids.cpu().numpy()
        rec = float(np.mean([len(set(ids[i].tolist()) & set(gt[i].tolist())) / k for i in range(nq)]))
        print(json.dumps({"builder": name, "n": n, "m": m, "build_s": round(build_s, 1), "mean_deg": round(float(deg.mean()), 2), **{("build_" + a): b for a, b in stats.items()},
                          "ef": ef, "ms": round(ms, 3), "qps": round(nq / ms * 1e3), "recall_at_10": round(rec, 4),
                          "evals_per_query": round(float(d_evals.cpu().numpy().view(np.uint32).mean()), 1)}), flush=True)
    h.deinit()
