#!/usr/bin/env python
"""K1 tuning sweep on one B200: time every visited-set variant (shared-memory hash, global bitmap, global hash) per ef on both
graphs. Prints one JSON line per (graph, ef, variant). Not part of the product or the tests."""
import json, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import zvdb_b200
from zvdb_b200 import builder

n, dim, nq, k, m = 1_000_000, 128, 10_000, 10, int(sys.argv[1]) if len(sys.argv) > 1 else 16
X = np.random.default_rng(1).standard_normal((n, dim), dtype=np.float32)
Q = np.random.default_rng(2).standard_normal((nq, dim), dtype=np.float32)
dev = torch.device("cuda", 0)
dq = torch.from_numpy(Q).to(dev)
d_ids = torch.empty((nq, k), dtype=torch.int64, device=dev)
d_dist = torch.empty((nq, k), dtype=torch.float32, device=dev)
d_cnt = torch.empty(nq, dtype=torch.int32, device=dev)
stream = torch.cuda.current_stream().cuda_stream
for graph in ("reference", "quality"):
    h = zvdb_b200.HNSW(m, 200)
    t0 = time.time()
    if graph == "reference": h.insert_batch(X)
    else: builder.build_quality_graph_incremental(h, X, m)
    h.sync_device()
    print(f"# {graph} built in {time.time()-t0:.1f}s", flush=True)
    for ef in (16, 32, 64, 128, 256, 512, 1024):
        for variant in (0b0100, 0b1000, 0b1100):
            h.set_kernel_variant(variant)
            try:
                for _ in range(2):
                    h.search_batch_device(dq.data_ptr(), nq, k, ef, d_ids.data_ptr(), d_dist.data_ptr(), d_cnt.data_ptr(), stream=stream)
                torch.cuda.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(3):
                    h.search_batch_device(dq.data_ptr(), nq, k, ef, d_ids.data_ptr(), d_dist.data_ptr(), d_cnt.data_ptr(), stream=stream)
                b.record(); torch.cuda.synchronize()
                ms = a.elapsed_time(b) / 3
                print(json.dumps({"graph": graph, "m": m, "ef": ef, "vis": {1: "smem_hash", 2: "global_bitmap", 3: "global_hash"}[variant >> 2], "ms": round(ms, 4), "qps": round(nq / ms * 1e3)}), flush=True)
            except zvdb_b200.ZvdbError as e:
                print(json.dumps({"graph": graph, "ef": ef, "variant": variant, "error": str(e)[:80]}), flush=True)
    h.deinit()
