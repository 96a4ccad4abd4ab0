"""CPU statement of the GPU graph builder (zvdb_b200/csrc/builder.cuh), for parity tests only.

The builder has no reference counterpart, so this file -- not hnsw.zig -- is
what the CUDA kernels are checked against, bit for bit: all distances in the kernel's summation
order (oracle ORC_DIST_TREE), all orderings by (distance, id)."""
import numpy as np

UNION_CAP = 96   # builder.cuh kUnionCap


def _sorted_keys(O, X, i, ids, metric):
    ids = np.array(sorted(set(int(x) for x in ids if 0 <= int(x) < len(X) and int(x) != i)), np.uint32)
    if len(ids) == 0:
        return []
    d = O.dist_many(X[i], X, ids, O.DIST_TREE | metric)
    keys = sorted(zip(d.tolist(), ids.tolist()))
    return keys


def _rng_select(O, X, keys, m, metric):
    sel, taken = [], []
    for idx, (d_ic, c) in enumerate(keys):
        if len(sel) == m:
            break
        ok = True
        if sel:
            ds = O.dist_many(X[c], X, np.array(sel, np.uint32), O.DIST_TREE | metric)
            ok = not bool((ds < np.float32(d_ic)).any())
        if ok:
            sel.append(c)
            taken.append(idx)
    return sel, set(taken)


def build_ref(O, X, cand, m, metric=0):
    X = np.ascontiguousarray(X, np.float32)
    n = len(X)
    fw = []
    for i in range(n):
        keys = _sorted_keys(O, X, i, cand[i], metric)
        fw.append(_rng_select(O, X, keys, m, metric)[0])
    rev = [[] for _ in range(n)]
    for i in range(n):
        for j in fw[i]:
            rev[j].append(i)
    adj = np.full((n, m), 0xFFFFFFFF, np.uint32)
    for i in range(n):
        keys = _sorted_keys(O, X, i, fw[i] + rev[i], metric)[:UNION_CAP]
        sel, taken = _rng_select(O, X, keys, m, metric)
        for idx, (_, c) in enumerate(keys):
            if len(sel) >= m:
                break
            if idx not in taken:
                sel.append(c)
        adj[i, :len(sel)] = sel
    return adj
