#!/usr/bin/env python
"""K1 A/B on one B200: L2 prefetch off / rows / adjacency / both (variant bits 8-10), per ef, on the
reference-insert and the quality graph (1M x 128) and, with `c3`, on 1M x 768 cosine M=32. Checks that ids,
distances and counters are identical across modes. One JSON line per (graph, ef, mode). Not product code."""
import json, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import zvdb_b200
from zvdb_b200 import builder

c3 = len(sys.argv) > 1 and sys.argv[1] == "c3"
n, dim, nq, k, m = (1_000_000, 768, 10_000, 100, 32) if c3 else (1_000_000, 128, 10_000, 10, 16)
X = np.random.default_rng(1).standard_normal((n, dim), dtype=np.float32)
Q = np.random.default_rng(2).standard_normal((nq, dim), dtype=np.float32)
if c3:
    X /= np.linalg.norm(X, axis=1, keepdims=True); Q /= np.linalg.norm(Q, axis=1, keepdims=True)
dev = torch.device("cuda", 0)
dq = torch.from_numpy(Q).to(dev)
mk = lambda dt, *s: torch.empty(s, dtype=dt, device=dev)
d_ids, d_dist, d_cnt, d_pops, d_evals = mk(torch.int64, nq, k), mk(torch.float32, nq, k), mk(torch.int32, nq), mk(torch.int32, nq), mk(torch.int32, nq)
stream = torch.cuda.current_stream().cuda_stream
for graph in (("quality",) if c3 else ("reference", "quality")):
    h = zvdb_b200.HNSW(m, 200, metric=zvdb_b200._lib.METRIC_COSINE if c3 else zvdb_b200._lib.METRIC_L2)
    t0 = time.time()
    if graph == "reference": h.insert_batch(X)
    else: builder.build_quality_graph(h, X, m)
    h.sync_device()
    print(f"# {graph} built in {time.time()-t0:.1f}s", flush=True)
    for ef in ((100, 200) if c3 else (10, 16, 32, 64, 128, 256, 512)):
        base = None
        for mode in (1, 2, 3, 4):
            h.set_kernel_variant(mode << 8)
            run = lambda: h.search_batch_device(dq.data_ptr(), nq, k, ef, d_ids.data_ptr(), d_dist.data_ptr(), d_cnt.data_ptr(),
                                                d_pops.data_ptr(), d_evals.data_ptr(), stream=stream)
            for _ in range(2): run()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(5): run()
            b.record(); torch.cuda.synchronize()
            ms = a.elapsed_time(b) / 5
            sig = (d_ids.cpu().numpy().tobytes(), d_dist.cpu().numpy().tobytes(), d_evals.cpu().numpy().tobytes(), d_pops.cpu().numpy().tobytes())
            same = True if base is None else sig == base
            base = base or sig
            ev = d_evals.cpu().numpy().view(np.uint32)
            rowb = ((dim + 31) // 32) * 128
            print(json.dumps({"graph": graph, "dim": dim, "m": m, "ef": ef, "prefetch": ["", "off", "rows", "adj", "rows+adj"][mode], "ms": round(ms, 4),
                              "qps": round(nq / ms * 1e3), "alg_gbs": round(float(ev.sum()) * rowb / ms / 1e6, 1), "identical_to_off": same}), flush=True)
    h.deinit()
