#!/usr/bin/env python
"""C4 strong-scaling baseline: the SAME id-sharded index the 8-GPU run searches (S per-shard indexes over
n rows, shard s = global ids s, s+S, ...; rows from bench.make_shard(n, dim, s, S), so bit-identical to
`bench.py --gpus S --rows n --shard-gen`), held and searched by ONE GPU.

A step = S shard-local searches of the full query batch (zvdb_search_batch_packed_device, each writing
block s of the gather buffer with global ids) + the merge kernel (zvdb_merge_topk_packed_device): the
8-GPU step minus the exchange, serialised on one device. One JSON line per ef on stdout.

    python scripts/c4_shards_one_gpu.py --rows 100000000 --shards 8 [--build-threads 8]
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (make_shard, algorithmic_bytes, load_peaks)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=100_000_000)
    ap.add_argument("--dim", type=int, default=128)
    ap.add_argument("--shards", type=int, default=8)
    ap.add_argument("--nq", type=int, default=10_000)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--m", type=int, default=16)
    ap.add_argument("--efs", default="32,64,128,256,512")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--build-threads", type=int, default=8)
    ap.add_argument("--graph", default="reference", choices=["reference", "incremental"],
                    help="per-shard graph: the reference's insert (host, shards in parallel) or the search-driven incremental builder (GPU, one shard after the other)")
    ap.add_argument("--min-free-gb", type=float, default=0.0, help="refuse to start with less free host memory than this")
    args = ap.parse_args()

    if args.min_free_gb > 0:
        free_kb = next(int(l.split()[1]) for l in open("/proc/meminfo") if l.startswith("MemAvailable"))
        if free_kb / 1e6 < args.min_free_gb:
            print(json.dumps({"skipped": f"host MemAvailable {free_kb / 1e6:.0f} GB < {args.min_free_gb} GB"}))
            return

    import torch
    import zvdb_b200
    from zvdb_b200 import _lib as L
    from zvdb_b200.sharded import block_bytes, per_shard_ef

    S, nq, k, dim = args.shards, args.nq, args.k, args.dim
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    stream = torch.cuda.current_stream().cuda_stream
    lib = L.lib()

    # ---- build: one reference-insert index per shard, shards in parallel on host threads (insert is
    # serial per index, hnsw.zig:74-75; distinct indexes share nothing) --------------------------------
    hs = [zvdb_b200.HNSW(args.m, 200, device=0) for _ in range(S)]
    t0 = time.time()
    gate = threading.Semaphore(args.build_threads)
    errs = []

    def build_one(s):
        with gate:
            try:
                Xs = bench.make_shard(args.rows, dim, s, S)
                hs[s].insert_batch(Xs)
                del Xs
            except Exception as e:   # noqa: BLE001
                errs.append((s, repr(e)))

    if args.graph == "incremental":
        from zvdb_b200 import builder
        for s in range(S):
            Xs = bench.make_shard(args.rows, dim, s, S)
            builder.build_quality_graph_incremental(hs[s], Xs, args.m)
            del Xs
            print(f"shard {s} built ({time.time() - t0:.0f}s)", file=sys.stderr, flush=True)
    else:
        th = [threading.Thread(target=build_one, args=(s,)) for s in range(S)]
        [t.start() for t in th]
        [t.join() for t in th]
    if errs:
        raise RuntimeError(f"shard build failed: {errs}")
    for h in hs:
        h.sync_device()
    build_s = time.time() - t0
    print(f"built {S} shards of {args.rows // S} x {dim} in {build_s:.1f}s", file=sys.stderr, flush=True)

    Qs = [np.random.default_rng(2 + b).standard_normal((nq, dim), dtype=np.float32) for b in range(bench.QUERY_BATCHES)]
    dq = [torch.from_numpy(q).to(dev) for q in Qs]
    bb = block_bytes(nq, k)
    gathered = torch.empty(S * bb, dtype=torch.uint8, device=dev)
    m_ids = torch.empty((nq, k), dtype=torch.int64, device=dev)
    m_dist = torch.empty((nq, k), dtype=torch.float32, device=dev)
    m_cnt = torch.empty(nq, dtype=torch.int32, device=dev)
    d_ids = torch.empty((nq, k), dtype=torch.int64, device=dev)
    d_dist = torch.empty((nq, k), dtype=torch.float32, device=dev)
    d_cnt = torch.empty(nq, dtype=torch.int32, device=dev)
    d_pops = torch.empty(nq, dtype=torch.int32, device=dev)
    d_evals = torch.empty(nq, dtype=torch.int32, device=dev)
    row_bytes = ((dim + 31) // 32) * 128
    peak, peak_src = bench.load_peaks()

    def step(b, e):
        for s in range(S):
            L.check(lib.zvdb_search_batch_packed_device(hs[s]._h, dq[b].data_ptr(), nq, k, e, gathered.data_ptr() + s * bb, S, s, stream))
        L.check(lib.zvdb_merge_topk_packed_device(gathered.data_ptr(), S, nq, k, m_dist.data_ptr(), m_ids.data_ptr(), m_cnt.data_ptr(), stream))

    for ef in [int(x) for x in args.efs.split(",")]:
        e = per_shard_ef(ef, k, S)
        # roofline numerator: the kernel's own counters, every shard, query batch 0
        job_bytes = 0
        evals = 0.0
        for s in range(S):
            hs[s].search_batch_device(dq[0].data_ptr(), nq, k, e, d_ids.data_ptr(), d_dist.data_ptr(), d_cnt.data_ptr(),
                                      d_pops.data_ptr(), d_evals.data_ptr(), id_stride=S, id_base=s, stream=stream)
            torch.cuda.synchronize()
            ev, po = d_evals.cpu().numpy().view(np.uint32), d_pops.cpu().numpy().view(np.uint32)
            job_bytes += bench.algorithmic_bytes(ev, po, row_bytes, args.m, dim, k)
            evals += float(ev.mean())
        for w in range(args.warmup):
            step(w % bench.QUERY_BATCHES, e)
        torch.cuda.synchronize()
        a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for s_ in range(args.steps):
            step(s_ % bench.QUERY_BATCHES, e)
        z.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(z) / args.steps
        step(0, e)                       # the checksum below is of batch 0, like bench.py's sweep lines (round 2's committed run took it
        torch.cuda.synchronize()         # from the last timed step = batch 1, so its sums do not compare with the 8-GPU file's)
        line = {"metric": "batched search QPS (id-sharded index held by one GPU)", "value": nq / (ms * 1e-3), "unit": "queries/s",
                "n_gpus": 1, "shards": S, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
                "config": {"workload": f"{args.rows}x{dim} fp32 L2 synthetic Gaussian, M={args.m}, {nq}-query batch, k={k}, ef={ef} "
                                       f"({e} pops/shard, {S} per-shard {args.graph}-graph indexes on one GPU, merge kernel, no exchange)",
                           "evals_per_query_all_shards": evals, "build_seconds": build_s},
                "roofline": {"bound": "hbm", "achieved": job_bytes / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                             "frac": job_bytes / (ms * 1e-3) / 1e9 / peak, "peak_source": peak_src,
                             "scope": f"the whole step ({S} search launches + merge)", "algorithmic_bytes_per_step": job_bytes},
                "gpu_launches_per_step": S + 1,
                "merged_ids_checksum": int(m_ids.sum().item())}
        print(json.dumps(line), flush=True)
    for h in hs:
        h.deinit()


if __name__ == "__main__":
    main()
