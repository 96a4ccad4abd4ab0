#!/usr/bin/env python
"""What the sharded step costs on top of the shard-local search, WITHOUT NVLink or rank skew: one GPU, world = 1, the shard of
the N=8 default line (125 000 rows = every 8th row of the 1M index, 10 pops, 10 000 queries): plain search vs the fused
step in its forms. One JSON line per form."""
import json, sys
import numpy as np, torch
sys.path.insert(0, ".")
import zvdb_b200
from zvdb_b200.sharded import CudaBackend
X = np.random.default_rng(1).standard_normal((1_000_000, 128), dtype=np.float32)[0::8]
Q = np.random.default_rng(2).standard_normal((10_000, 128), dtype=np.float32)
dev = torch.device("cuda", 0); stream = torch.cuda.current_stream().cuda_stream
nq, k = len(Q), 10
dq = torch.from_numpy(Q).to(dev)
ids = torch.empty((nq, k), dtype=torch.int64, device=dev); dist = torch.empty((nq, k), dtype=torch.float32, device=dev); cnt = torch.empty(nq, dtype=torch.int32, device=dev)
def timed(fn, reps=20):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
for ef in (10, 16, 32):
    for form, variant in (("plain search", None), ("fused, records (p2p)", 0), ("fused, blocks + flags (p2pb)", 0x2000), ("three launches (p2p3)", 0x1000)):
        h = zvdb_b200.HNSW(16, 200)
        h.insert_batch(X); h.sync_device()
        if variant is None:
            ms = timed(lambda: h.search_batch_device(dq.data_ptr(), nq, k, ef, ids.data_ptr(), dist.data_ptr(), cnt.data_ptr(), id_stride=8, id_base=0, stream=stream))
        else:
            h.set_kernel_variant(variant)
            be = CudaBackend(h, 0, 1); be.open_exchange(nq, k, None)
            ms = timed(lambda: be.search_exchange(dq, nq, k, ef, out=(ids, dist, cnt)))
            be.close()
        print(json.dumps({"rows": len(X), "pops": ef, "form": form, "ms_per_step": round(ms, 4)}), flush=True)
        h.deinit()
