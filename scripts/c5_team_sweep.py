#!/usr/bin/env python
"""C5 (BASELINE configs[4]) A/B: the one-warp-per-query kernel K1 against the one-CTA-per-query latency form K1L
(zvdb_set_kernel_variant bits 14-15 = 1 / 2; 3 = K1L with teams of 4 warps) over the batch sizes where they compete, 1M x 128, on the reference
graph and on the incremental quality graph. Device time per launch (CUDA events, successive launches take successive
slices of a 65 536-query pool so a launch never repeats the previous one's queries) and the host-call time of
search_batch on pageable numpy buffers. Results of both kernels are compared row for row. JSON lines.
usage: python scripts/c5_team_sweep.py [reference|incremental|both]"""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import zvdb_b200
from zvdb_b200 import builder

dev = torch.device("cuda", 0)
stream = torch.cuda.current_stream().cuda_stream
which = sys.argv[1] if len(sys.argv) > 1 else "both"
NEVER, ALWAYS, HALF = 1 << 14, 2 << 14, 3 << 14
# shape overrides (defaults = BASELINE configs[4]): C5_DIM, C5_M, C5_K, C5_EFS ("10,64"), C5_METRIC (0 L2 / 1 cosine / 2 dot), C5_NQ ("1,64,...")
n, dim, k, m = 1_000_000, int(os.environ.get("C5_DIM", 128)), int(os.environ.get("C5_K", 10)), int(os.environ.get("C5_M", 16))
metric = int(os.environ.get("C5_METRIC", 0))
efs = [int(x) for x in os.environ.get("C5_EFS", "10,64").split(",")]
X = np.random.default_rng(1).standard_normal((n, dim), dtype=np.float32)
Qall = np.random.default_rng(2).standard_normal((65536, dim), dtype=np.float32)
dq = torch.from_numpy(Qall).to(dev)
d_ids = torch.empty((65536, k), dtype=torch.int64, device=dev); d_dist = torch.empty((65536, k), dtype=torch.float32, device=dev)
d_cnt = torch.empty(65536, dtype=torch.int32, device=dev)


def device_ms(h, nq, ef, reps):
    slices = max(1, min(64, 65536 // nq))
    def launch(i):
        off = (i % slices) * nq
        h.search_batch_device(dq.data_ptr() + off * dim * 4, nq, k, ef, d_ids.data_ptr() + off * k * 8, d_dist.data_ptr() + off * k * 4,
                              d_cnt.data_ptr() + off * 4, stream=stream)
    for i in range(5): launch(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(reps): launch(i)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


for graph in (("reference", "incremental") if which == "both" else (which,)):
    h = zvdb_b200.HNSW(m, 200, metric=metric)
    if graph == "reference": h.insert_batch(X)
    else: builder.build_quality_graph_incremental(h, X, m)
    h.sync_device()
    for ef in efs:
        for nq in ([int(x) for x in os.environ["C5_NQ"].split(",")] if os.environ.get("C5_NQ") else (1, 8, 64, 148, 296, 592, 1024, 2048, 4096)):
            row = {"config": "C5", "graph": graph, "dim": dim, "m": m, "k": k, "metric": metric, "ef": ef, "nq": nq}
            keep = None
            for name, variant in (("k1", NEVER), ("k1l", ALWAYS), ("k1l_half", HALF)):
                h.set_kernel_variant(variant)
                row[name + "_device_ms"] = round(device_ms(h, nq, ef, 200 if nq <= 296 else 50), 5)
                torch.cuda.synchronize()
                got = (d_ids[:nq].cpu().numpy(), d_dist[:nq].cpu().numpy(), d_cnt[:nq].cpu().numpy())
                if keep is None: keep = got
                else: row["rows_equal"] = bool(all(np.array_equal(x.view(np.uint8), y.view(np.uint8)) for x, y in zip(keep, got)))
                if nq <= 296:
                    hq = Qall[:nq].copy()
                    h.search_batch(hq, k, ef)
                    best = 1e9
                    for _ in range(20):
                        t0 = time.perf_counter(); h.search_batch(hq, k, ef); best = min(best, (time.perf_counter() - t0) * 1e3)
                    row[name + "_host_call_ms"] = round(best, 5)
            row["speedup_device"] = round(row["k1_device_ms"] / row["k1l_device_ms"], 3)
            print(json.dumps(row), flush=True)
    h.deinit()
