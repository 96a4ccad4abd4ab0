#!/usr/bin/env python
"""Round 2, VERDICT r1 item 7: recall@10 vs row evaluations per query on builder graphs over 1M x 128 Gaussian rows,
M = 16 / 32, with and without the upper-layer descent (hierarchy from builder.build_hierarchy), next to the QPS the
HBM gather roofline allows at that many evaluations (measured peak / (evals * 512 B)). One JSON line per point.
Candidates come from the repo's own exact k-NN (K4), ground truth too. Not part of the product or the tests."""
import json, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import zvdb_b200
from zvdb_b200 import builder

n, dim, nq, k = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000, 128, 10_000, 10
peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if len(sys.argv) < 3 else float(sys.argv[2])
X = np.random.default_rng(1).standard_normal((n, dim), dtype=np.float32)
Q = np.random.default_rng(2).standard_normal((nq, dim), dtype=np.float32)
dev = torch.device("cuda", 0)
dq = torch.from_numpy(Q).to(dev)
ids = torch.empty((nq, k), dtype=torch.int64, device=dev)
dist = torch.empty((nq, k), dtype=torch.float32, device=dev)
cnt = torch.empty(nq, dtype=torch.int32, device=dev)
pops = torch.empty(nq, dtype=torch.int32, device=dev)
evals = torch.empty(nq, dtype=torch.int32, device=dev)
stream = torch.cuda.current_stream().cuda_stream
gt = None
for m in (16, 32):
    for source in ("k4_exact", "incremental"):
        h = zvdb_b200.HNSW(m, 200)
        t0 = time.time()
        if source == "k4_exact":
            builder.build_quality_graph(h, X, m, K=64)
        else:
            builder.build_quality_graph_incremental(h, X, m)
        t_graph = time.time() - t0
        t0 = time.time()
        levels, upper, start = builder.build_hierarchy(h, X, m, seed=3)
        t_hier = time.time() - t0
        h.sync_device()
        if gt is None:
            h.bruteforce_knn_device(dq.data_ptr(), nq, k, ids.data_ptr(), dist.data_ptr(), cnt.data_ptr(), stream=stream)
            torch.cuda.synchronize()
            gt = ids.cpu().numpy().copy()
        adj, deg = h.export_layer(0)
        print(f"# m={m} {source}: layer 0 in {t_graph:.1f}s (mean degree {deg.mean():.2f}), hierarchy ({int(levels.max())} levels) in {t_hier:.1f}s", flush=True)
        for descent in (False, True):
            h.set_descent(descent)
            for ef in (16, 32, 64, 128, 256, 512, 1024):
                ms = []
                for r in range(4):
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record()
                    h.search_batch_device(dq.data_ptr(), nq, k, ef, ids.data_ptr(), dist.data_ptr(), cnt.data_ptr(), pops.data_ptr(), evals.data_ptr(), stream=stream)
                    b.record(); torch.cuda.synchronize()
                    if r:
                        ms.append(a.elapsed_time(b))
                t = float(np.median(ms))
                got = ids.cpu().numpy()
                rec = float(np.mean([len(set(got[i].tolist()) & set(gt[i].tolist())) / k for i in range(nq)]))
                ev = float(evals.cpu().numpy().astype(np.int64).mean())
                print(json.dumps({"m": m, "candidates": source, "descent": descent, "ef": ef, "ms": round(t, 4), "qps": round(nq / t * 1e3),
                                  "recall_at_10": round(rec, 4), "evals_per_query": round(ev, 1),
                                  "roofline_qps_at_these_evals": round(peak * 1e9 / (ev * 512 + 64 * ef / 1 + 512 + 120)),
                                  "frac_of_hbm_peak": round((ev * 512 + ef * m * 4 + 512 + 120) * nq / (t * 1e-3) / 1e9 / peak, 3)}), flush=True)
        h.deinit()
