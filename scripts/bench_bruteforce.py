#!/usr/bin/env python
"""K4 measurement on one B200: exact k-NN of nq queries against n rows; reports algorithmic
TFLOP/s (2*nq*n*dim) against the measured dense bf16 peak / 2 (TF32 dense) and recall
vs a torch fp32 matmul reference on a sample. One JSON line."""
import argparse, json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zvdb_b200

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1_000_000); ap.add_argument("--dim", type=int, default=128)
ap.add_argument("--nq", type=int, default=10_000); ap.add_argument("--k", type=int, default=10)
ap.add_argument("--metric", default="l2"); ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--filter", action="store_true", help="single-product TF32 filter + exact re-rank (approximate) instead of 3xTF32")
a = ap.parse_args()
dev = torch.device("cuda", 0)
X = np.random.default_rng(1).standard_normal((a.n, a.dim), dtype=np.float32)
Q = np.random.default_rng(2).standard_normal((a.nq, a.dim), dtype=np.float32)
metric = {"l2": 0, "cos": 1, "dot": 2}[a.metric]
h = zvdb_b200.HNSW(16, 200, metric=metric)
# rows only: an edge-less graph (the brute-force path never reads adjacency)
h.load_graph(X, np.zeros(a.n + 1, np.uint64), np.zeros(0, np.uint32), 0)
h.sync_device()
if a.filter: h.set_kernel_variant(1 << 6)
dq = torch.from_numpy(Q).to(dev)
d_ids = torch.empty((a.nq, a.k), dtype=torch.int64, device=dev)
d_dist = torch.empty((a.nq, a.k), dtype=torch.float32, device=dev)
d_cnt = torch.empty(a.nq, dtype=torch.int32, device=dev)
s = torch.cuda.current_stream().cuda_stream
def step():
    h.bruteforce_knn_device(dq.data_ptr(), a.nq, a.k, d_ids.data_ptr(), d_dist.data_ptr(), d_cnt.data_ptr(), stream=s)
for _ in range(2): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.steps): step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.steps
flops = 2.0 * a.nq * a.n * a.dim
# reference on a sample of queries: fp32 matmul (TF32 off) + topk
torch.backends.cuda.matmul.allow_tf32 = False
Xd = torch.from_numpy(X).to(dev) if metric != 1 else torch.nn.functional.normalize(torch.from_numpy(X).to(dev).double(), dim=1).float()
ns = min(a.nq, 512)
qs = dq[:ns]
if metric == 0: sc = (Xd * Xd).sum(1)[None, :] - 2.0 * (qs @ Xd.T)
else: sc = -(qs @ Xd.T)
ref = torch.topk(sc, a.k, dim=1, largest=False).indices.cpu().numpy()
got = d_ids[:ns].cpu().numpy()
rec = np.mean([len(set(got[i].tolist()) & set(ref[i].tolist())) / a.k for i in range(ns)])
peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json"))) if os.path.exists("MEASURED_PEAKS.json") else {"bf16_tflops": 1590.0}
tf32_peak = peaks["bf16_tflops"] / 2
tf32_sustained = peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"]) / 2
terms = 1 if a.filter else 3
print(json.dumps({"kernel": "bf_gemm_topk_kernel (+split, +finalize)", "n": a.n, "dim": a.dim, "nq": a.nq, "k": a.k, "metric": a.metric,
                  "ms": ms, "qps": a.nq / ms * 1e3, "algorithmic_tflops": flops / ms / 1e9, "mode": "1xTF32 filter + exact re-rank of k+24 (approximate)" if a.filter else "3xTF32 (exact)",
                  "issued_tflops": terms * flops / ms / 1e9,
                  "tf32_dense_peak_tflops(bf16_measured/2)": tf32_peak, "tf32_sustained_peak_tflops(bf16_sustained/2)": tf32_sustained,
                  "frac_algorithmic": flops / ms / 1e9 / tf32_peak,
                  "frac_issued": terms * flops / ms / 1e9 / tf32_peak, "frac_issued_of_sustained": terms * flops / ms / 1e9 / tf32_sustained, "agreement_with_fp32_matmul_topk": rec,
                  "kernel_launches": h.kernel_launches()}))
