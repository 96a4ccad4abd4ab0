"""ctypes front-end for the CPU oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE, NOT PRODUCT: only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module. Nothing under zvdb_b200/ does.

The oracle restates /root/reference/src/hnsw.zig (see zvdb_oracle.c for the line map).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

DIST_SEQ, DIST_TREE = 0, 1
METRIC_L2, METRIC_COS, METRIC_DOT = 0x00, 0x10, 0x20
HEAP_ZIG, HEAP_DET = 0, 1

_NP = {"f32": np.float32, "f64": np.float64, "i32": np.int32}
_CT = {"f32": C.c_float, "f64": C.c_double, "i32": C.c_int32}


def build(force: bool = False) -> str:
    """Compile oracle/liboracle.so with the committed Makefile (gcc, -ffp-contract=off)."""
    src_m = max(os.path.getmtime(os.path.join(_HERE, f)) for f in ("zvdb_oracle.c", "oracle_impl.h", "Makefile"))
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < src_m:
        subprocess.run(["make", "-C", _HERE, "-s", "-B", "liboracle.so"], check=True)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        vp, sz, i32, u64, lg = C.c_void_p, C.c_size_t, C.c_int, C.c_uint64, C.c_long
        for s, ct in _CT.items():
            p = C.POINTER(ct)
            g = lambda name: getattr(L, f"{name}_{s}")
            g("orc_create").restype = vp
            g("orc_create").argtypes = [i32, i32, u64]
            g("orc_destroy").argtypes = [vp]
            g("orc_destroy").restype = None
            g("orc_set_dist_mode").argtypes = [vp, i32]
            g("orc_insert").argtypes = [vp, p, i32, i32]
            g("orc_insert_batch").argtypes = [vp, p, sz, i32, C.POINTER(C.c_int)]
            g("orc_count").argtypes = [vp]
            g("orc_count").restype = sz
            g("orc_dim").argtypes = [vp]
            g("orc_max_level").argtypes = [vp]
            g("orc_entry").argtypes = [vp]
            g("orc_entry").restype = lg
            g("orc_level").argtypes = [vp, sz]
            g("orc_points").argtypes = [vp]
            g("orc_points").restype = p
            g("orc_export_layer").argtypes = [vp, i32, sz, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
            g("orc_export_layer").restype = None
            g("orc_search").argtypes = [vp, p, sz, i32, i32, C.POINTER(C.c_uint32), p,
                                        C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
            g("orc_search").restype = lg
            g("orc_search_graph").argtypes = [p, i32, sz, C.POINTER(C.c_uint32), sz, lg, p, sz, sz, sz,
                                              i32, i32, i32, i32, C.POINTER(C.c_uint32), p,
                                              C.POINTER(C.c_uint32), C.POINTER(C.c_uint32),
                                              C.POINTER(C.c_uint32)]
            pu8, pu32 = C.POINTER(C.c_uint8), C.POINTER(C.c_uint32)
            g("orc_search_graph_desc").argtypes = [p, i32, sz, pu32, sz, lg, p, sz, sz, sz, i32, i32, i32, i32,
                                                   pu8, pu32, pu32, i32, lg, pu32, p, pu32, pu32, pu32]
            g("orc_descend_one").argtypes = [p, i32, pu8, pu32, pu32, sz, i32, lg, p, i32, pu32, p]
            g("orc_descend_one").restype = lg
        L.orc_max_threads.restype = i32
        L.orc_distance_f32.restype = C.c_float
        L.orc_distance_f32.argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_float), i32, i32]
        L.orc_dist_many_f32.restype = None
        L.orc_dist_many_f32.argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_float), i32, C.POINTER(C.c_uint32), sz,
                                        i32, C.POINTER(C.c_float)]
        L.orc_bruteforce_f32.argtypes = [C.POINTER(C.c_float), sz, i32, C.POINTER(C.c_float), sz, sz, i32, i32,
                                         C.POINTER(C.c_uint32), C.POINTER(C.c_float)]
        L.orc_merge_topk.argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_uint64), C.POINTER(C.c_uint32), i32,
                                     sz, sz, C.POINTER(C.c_float), C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)]
        L.orc_merge_topk.restype = None
        _lib = L
    return _lib


def _ptr(a, ct):
    return a.ctypes.data_as(C.POINTER(ct))


def max_threads() -> int:
    return int(lib().orc_max_threads())


def distance(a, b, mode=DIST_SEQ) -> np.float32:
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    return np.float32(lib().orc_distance_f32(_ptr(a, C.c_float), _ptr(b, C.c_float), a.size, mode))


def dist_many(a, points, ids, mode=DIST_SEQ) -> np.ndarray:
    """Distances from vector `a` to rows points[ids] (points must be C-contiguous float32)."""
    a = np.ascontiguousarray(a, np.float32)
    ids = np.ascontiguousarray(ids, np.uint32)
    out = np.empty(len(ids), np.float32)
    lib().orc_dist_many_f32(_ptr(a, C.c_float), _ptr(points, C.c_float), points.shape[1], _ptr(ids, C.c_uint32),
                            len(ids), mode, _ptr(out, C.c_float))
    return out


class OracleHNSW:
    """Restatement of `HNSW(T)` (hnsw.zig:8): init(m, ef_construction), insert(point), search(query, k)."""

    def __init__(self, m: int = 16, ef_construction: int = 200, dtype: str = "f32", seed: int = 0,
                 dist_mode: int = DIST_SEQ):
        self.s = dtype
        self.np = _NP[dtype]
        self.ct = _CT[dtype]
        self.m = m
        self._f = lambda name: getattr(lib(), f"{name}_{dtype}")
        self.h = self._f("orc_create")(m, ef_construction, seed)
        self.dist_mode = dist_mode
        self._f("orc_set_dist_mode")(self.h, dist_mode)

    def __del__(self):
        if getattr(self, "h", None):
            self._f("orc_destroy")(self.h)
            self.h = None

    def insert(self, point, level: int = -1) -> None:
        p = np.ascontiguousarray(point, self.np)
        rc = self._f("orc_insert")(self.h, _ptr(p, self.ct), p.size, level)
        if rc == -2:
            raise ValueError("Mismatched dimensions in distance calculation")
        if rc:
            raise MemoryError

    def insert_batch(self, points, levels=None) -> None:
        p = np.ascontiguousarray(points, self.np)
        lv = None
        if levels is not None:
            lv = np.ascontiguousarray(levels, np.int32)
        rc = self._f("orc_insert_batch")(self.h, _ptr(p, self.ct), p.shape[0], p.shape[1],
                                         _ptr(lv, C.c_int) if lv is not None else None)
        if rc:
            raise RuntimeError(f"oracle insert failed rc={rc}")

    def count(self) -> int:
        return int(self._f("orc_count")(self.h))

    @property
    def dim(self) -> int:
        return int(self._f("orc_dim")(self.h))

    @property
    def max_level(self) -> int:
        return int(self._f("orc_max_level")(self.h))

    @property
    def entry_point(self):
        e = int(self._f("orc_entry")(self.h))
        return None if e < 0 else e

    def level(self, i: int) -> int:
        return int(self._f("orc_level")(self.h, i))

    def points(self) -> np.ndarray:
        n, d = self.count(), self.dim
        if n == 0:
            return np.zeros((0, max(d, 0)), self.np)
        p = self._f("orc_points")(self.h)
        return np.ctypeslib.as_array(p, shape=(n, d)).copy()

    def export_layer(self, layer: int = 0, pitch: int | None = None):
        """Padded adjacency table adj[n, pitch] (0xFFFFFFFF padding) and degrees deg[n]."""
        n = self.count()
        pitch = pitch or self.m
        adj = np.full((max(n, 1), pitch), 0xFFFFFFFF, np.uint32)
        deg = np.zeros(max(n, 1), np.uint32)
        if n:
            self._f("orc_export_layer")(self.h, layer, pitch, _ptr(adj, C.c_uint32), _ptr(deg, C.c_uint32))
        return adj[:n], deg[:n]

    def export_upper(self):
        """Layers >= 1 in the flat form of zvdb_export_upper_layers: (levels[n] u8, upper_base[n] u32,
        upper_adj[n_lists, m] u32, max_level, start) with start = the first node of maximum level."""
        n = self.count()
        levels = np.array([self.level(i) for i in range(n)], np.uint8)
        base = np.full(n, 0xFFFFFFFF, np.uint32)
        has = levels > 0
        base[has] = (np.cumsum(levels.astype(np.uint64)) - levels)[has].astype(np.uint32)
        nl = int(levels.astype(np.uint64).sum())
        adj = np.full((nl, self.m), 0xFFFFFFFF, np.uint32)
        for layer in range(1, int(levels.max(initial=0)) + 1):
            tab, _ = self.export_layer(layer)
            idx = np.nonzero(levels >= layer)[0]
            adj[base[idx].astype(np.int64) + (layer - 1)] = tab[idx]
        mx = int(levels.max(initial=0))
        start = int(np.argmax(levels == mx)) if n else -1
        return levels, base, adj, mx, start

    def search(self, query, k: int, heap_mode: int = HEAP_ZIG, dist_mode: int | None = None,
               counters: bool = False):
        """The reference call `search(query, k)` (hnsw.zig:194): (ids, distances[, pops, evals])."""
        q = np.ascontiguousarray(query, self.np)
        ids = np.zeros(k + 1, np.uint32)
        d = np.zeros(k + 1, self.np)
        pops, evals = C.c_uint32(0), C.c_uint32(0)
        dm = self.dist_mode if dist_mode is None else dist_mode
        r = self._f("orc_search")(self.h, _ptr(q, self.ct), k, dm, heap_mode, _ptr(ids, C.c_uint32),
                                  _ptr(d, self.ct), C.byref(pops), C.byref(evals))
        if r < 0:
            raise MemoryError
        if counters:
            return ids[:r].copy(), d[:r].copy(), pops.value, evals.value
        return ids[:r].copy(), d[:r].copy()


def search_graph(points, adj, queries, ef: int, k: int | None = None, entry: int = 0,
                 dist_mode: int = DIST_SEQ, heap_mode: int = HEAP_ZIG, nthreads: int = 0,
                 global_lock: bool = False, dtype: str = "f32", upper=None):
    """Batched `search(q, ef)[0..k]` on a supplied padded graph. Returns dict(ids, dist, counts, pops, evals).
    upper = (levels, upper_base, upper_adj, max_level, start): first descend layers max_level..1 from
    `start` (extension, see orc_descend in oracle_impl.h) and begin the layer-0 search where that lands."""
    npdt, ct = _NP[dtype], _CT[dtype]
    pts = np.ascontiguousarray(points, npdt)
    adj = np.ascontiguousarray(adj, np.uint32)
    q = np.ascontiguousarray(queries, npdt).reshape(-1, pts.shape[1])
    k = ef if k is None else min(k, ef)
    nq = q.shape[0]
    ids = np.zeros((nq, k), np.uint32)
    d = np.zeros((nq, k), npdt)
    counts = np.zeros(nq, np.uint32)
    pops = np.zeros(nq, np.uint32)
    evals = np.zeros(nq, np.uint32)
    n = pts.shape[0]
    if upper is not None:
        lv, ub, ua, mx, start = upper
        lv = np.ascontiguousarray(lv, np.uint8); ub = np.ascontiguousarray(ub, np.uint32)
        ua = np.ascontiguousarray(ua, np.uint32)
        if ua.size == 0:
            ua = np.full((1, adj.shape[1]), 0xFFFFFFFF, np.uint32)
        assert ua.shape[1] == adj.shape[1], "upper lists and layer 0 share the pitch m"
        rc = getattr(lib(), f"orc_search_graph_desc_{dtype}")(
            _ptr(pts, ct), pts.shape[1], n, _ptr(adj, C.c_uint32), adj.shape[1], entry if n else -1, _ptr(q, ct), nq,
            ef, k, dist_mode, heap_mode, nthreads, int(global_lock), _ptr(lv, C.c_uint8), _ptr(ub, C.c_uint32),
            _ptr(ua, C.c_uint32), int(mx), int(start), _ptr(ids, C.c_uint32), _ptr(d, ct), _ptr(counts, C.c_uint32),
            _ptr(pops, C.c_uint32), _ptr(evals, C.c_uint32))
    else:
        rc = getattr(lib(), f"orc_search_graph_{dtype}")(
            _ptr(pts, ct), pts.shape[1], n, _ptr(adj, C.c_uint32), adj.shape[1] if adj.ndim == 2 else 1,
            entry if n else -1, _ptr(q, ct), nq, ef, k, dist_mode, heap_mode, nthreads, int(global_lock),
            _ptr(ids, C.c_uint32), _ptr(d, ct), _ptr(counts, C.c_uint32), _ptr(pops, C.c_uint32),
            _ptr(evals, C.c_uint32))
    if rc:
        raise MemoryError
    return {"ids": ids, "dist": d, "counts": counts, "pops": pops, "evals": evals}


def descend_one(points, upper, query, dist_mode: int = DIST_SEQ, dtype: str = "f32"):
    """Where the descent lands for one query: (node id, its distance, distance evaluations)."""
    npdt, ct = _NP[dtype], _CT[dtype]
    pts = np.ascontiguousarray(points, npdt)
    q = np.ascontiguousarray(query, npdt)
    lv, ub, ua, mx, start = upper
    lv = np.ascontiguousarray(lv, np.uint8); ub = np.ascontiguousarray(ub, np.uint32); ua = np.ascontiguousarray(ua, np.uint32)
    ev = C.c_uint32(0)
    dd = ct(0)
    node = getattr(lib(), f"orc_descend_one_{dtype}")(_ptr(pts, ct), pts.shape[1], _ptr(lv, C.c_uint8), _ptr(ub, C.c_uint32),
                                                      _ptr(ua, C.c_uint32), ua.shape[1] if ua.ndim == 2 else 1, int(mx),
                                                      int(start), _ptr(q, ct), dist_mode, C.byref(ev), C.byref(dd))
    return int(node), dd.value, ev.value


def bruteforce(points, queries, k: int, metric: int = 0, nthreads: int = 0):
    pts = np.ascontiguousarray(points, np.float32)
    q = np.ascontiguousarray(queries, np.float32).reshape(-1, pts.shape[1])
    k = min(k, pts.shape[0])
    ids = np.zeros((q.shape[0], k), np.uint32)
    d = np.zeros((q.shape[0], k), np.float32)
    rc = lib().orc_bruteforce_f32(_ptr(pts, C.c_float), pts.shape[0], pts.shape[1], _ptr(q, C.c_float),
                                  q.shape[0], k, metric, nthreads, _ptr(ids, C.c_uint32), _ptr(d, C.c_float))
    if rc:
        raise MemoryError
    return ids, d


def merge_topk(dist, ids, counts):
    """dist/ids: [G, nq, k], counts: [G, nq] -> (dist[nq,k], ids[nq,k], counts[nq]) by (distance, id)."""
    dist = np.ascontiguousarray(dist, np.float32)
    ids = np.ascontiguousarray(ids, np.uint64)
    counts = np.ascontiguousarray(counts, np.uint32)
    G, nq, k = dist.shape
    do = np.zeros((nq, k), np.float32)
    io = np.zeros((nq, k), np.uint64)
    co = np.zeros(nq, np.uint32)
    lib().orc_merge_topk(_ptr(dist, C.c_float), _ptr(ids, C.c_uint64), _ptr(counts, C.c_uint32), G, nq, k,
                         _ptr(do, C.c_float), _ptr(io, C.c_uint64), _ptr(co, C.c_uint32))
    return do, io, co
