"""A/B timing of K1 builds: python scripts/dev/ab_time.py <tag>   (library chosen by ZVDB_B200_LIB). One JSON line per point."""
import json, sys, time
import numpy as np, torch
sys.path.insert(0, __import__("os").environ.get("ZVDB_TREE", "."))
import zvdb_b200
from zvdb_b200 import builder
tag = sys.argv[1]
n, dim, nq, k, m = 1_000_000, 128, 10_000, 10, 16
X = np.random.default_rng(1).standard_normal((n, dim), dtype=np.float32)
Q = np.random.default_rng(2).standard_normal((nq, dim), dtype=np.float32)
dev = torch.device("cuda", 0)
dq = torch.from_numpy(Q).to(dev)
ids = torch.empty((nq, k), dtype=torch.int64, device=dev); dist = torch.empty((nq, k), dtype=torch.float32, device=dev); cnt = torch.empty(nq, dtype=torch.int32, device=dev)
stream = torch.cuda.current_stream().cuda_stream
for graph in ("reference", "incremental"):
    h = zvdb_b200.HNSW(m, 200)
    if graph == "reference": h.insert_batch(X)
    else: builder.build_quality_graph_incremental(h, X, m)
    h.sync_device()
    for ef in (64, 128, 256, 512):
        for variant in ((0,) if ef == 64 else (8, 12)):
            try:
                h.set_kernel_variant(variant)
            except Exception:
                continue
            ms = []
            for r in range(7):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); h.search_batch_device(dq.data_ptr(), nq, k, ef, ids.data_ptr(), dist.data_ptr(), cnt.data_ptr(), stream=stream); b.record(); torch.cuda.synchronize()
                if r >= 2: ms.append(a.elapsed_time(b))
            print(json.dumps({"tag": tag, "graph": graph, "ef": ef, "variant": variant, "ms_median": round(float(np.median(ms)), 4), "ms_min": round(min(ms), 4)}), flush=True)
    h.deinit()
