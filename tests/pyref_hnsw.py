"""A second, independent restatement of the reference (src/hnsw.zig), in plain Python, for small cases only.

Test infrastructure: it exists to cross-check the C oracle (oracle/zvdb_oracle.c) statement by statement -- two
restatements written apart from each other, from the same reference lines, must agree on the graph, on the search
results (ids, distance bits) and on the counters. Like the oracle it states the Zig standard library's
`PriorityQueue` (binary heap: siftUp moves a child up while it is strictly less than its parent; siftDown takes the
right child only if strictly less than the left one and stops once the moved element is strictly less than the
lesser child) and `sort.insertion` (stable) from Zig 0.13, which is not available in this image.
Arithmetic is the reference's: per-element f32 rounding, sequential sum, no fused multiply-add (hnsw.zig:182-192).
"""
import numpy as np

F = np.float32


def distance(a, b):                                            # hnsw.zig:182-192, in the element type T of the rows
    assert len(a) == len(b), "Mismatched dimensions in distance calculation"
    T = a.dtype.type
    s = T(0)
    for i in range(len(a)):
        diff = T(a[i] - b[i])
        s = T(s + T(diff * diff))
    return s


class ZigPriorityQueue:                                        # std.PriorityQueue(CandidateNode, void, lessThan), hnsw.zig:202
    """Min-heap on .distance alone (hnsw.zig:238-245); entries are (id, distance)."""

    def __init__(self):
        self.items = []

    def count(self):
        return len(self.items)

    @staticmethod
    def _lt(a, b):
        return a[1] < b[1]

    def add(self, e):
        self.items.append(e)
        child_index = len(self.items) - 1
        child = e
        while child_index > 0:                                 # siftUp
            parent_index = (child_index - 1) >> 1
            parent = self.items[parent_index]
            if not self._lt(child, parent):
                break
            self.items[child_index] = parent
            child_index = parent_index
        self.items[child_index] = child

    def remove(self):
        last = self.items[-1]
        item = self.items[0]
        self.items[0] = last
        self.items.pop()
        if self.items:                                         # siftDown(0)
            target = self.items[0]
            index = 0
            n = len(self.items)
            while True:
                lesser = (index * 2) | 1
                if not lesser < n:
                    break
                nxt = lesser + 1
                if nxt < n and self._lt(self.items[nxt], self.items[lesser]):
                    lesser = nxt
                if self._lt(target, self.items[lesser]):
                    break
                self.items[index] = self.items[lesser]
                index = lesser
            self.items[index] = target
        return item


def insertion_sort(items, less):                               # std.sort.insertion: stable
    for i in range(1, len(items)):
        x = items[i]
        j = i
        while j > 0 and less(x, items[j - 1]):
            items[j] = items[j - 1]
            j -= 1
        items[j] = x


class PyHNSW:
    def __init__(self, m, dtype=F):                            # HNSW(T).init, hnsw.zig:8, :52-62
        self.T = np.dtype(dtype).type
        self.m = m
        self.points = []
        self.conn = []                                         # conn[id][layer] = list of ids
        self.entry_point = None
        self.max_level = 0

    def insert(self, point, level):                            # hnsw.zig:73-117 (level: randomLevel's draw, :172-180)
        nid = len(self.points)
        p = np.asarray(point, self.T)
        self.points.append(p)
        self.conn.append([[] for _ in range(level + 1)])
        if self.entry_point is not None:
            ep = self.entry_point
            curr = distance(p, self.points[ep])
            for layer in range(self.max_level + 1):
                changed = True
                while changed:
                    changed = False
                    cur_conn = self.conn[ep]                   # captured before the scan (:92)
                    if layer < len(cur_conn):
                        for nb in list(cur_conn[layer]):
                            d = distance(p, self.points[nb])
                            if d < curr:
                                ep, curr, changed = nb, d, True
                if layer <= level:
                    self.connect(nid, ep, layer)
        else:
            self.entry_point = nid
        if level > self.max_level:
            self.max_level = level

    def connect(self, source, target, level):                  # hnsw.zig:119-141
        if level < len(self.conn[source]):
            self.conn[source][level].append(target)
        if level < len(self.conn[target]):
            self.conn[target][level].append(source)
        if level < len(self.conn[source]):
            self.shrink(source, level)
        if level < len(self.conn[target]):
            self.shrink(target, level)

    def shrink(self, node, level):                             # hnsw.zig:143-170
        c = self.conn[node][level]
        if len(c) <= self.m:
            return
        cand = list(c)
        p = self.points[node]
        insertion_sort(cand, lambda a, b: distance(p, self.points[a]) < distance(p, self.points[b]))
        self.conn[node][level] = cand[: self.m]

    def search(self, query, k):                                # hnsw.zig:194-236 -> (ids, distances, pops, evals)
        q = np.asarray(query, self.T)
        result, evals = [], 0
        if self.entry_point is not None:
            cands = ZigPriorityQueue()
            visited = set()
            e = self.entry_point
            cands.add((e, distance(q, self.points[e]))); evals += 1
            visited.add(e)
            while cands.count() > 0 and len(result) < k:
                cur = cands.remove()
                result.append(cur[0])
                for nb in self.conn[cur[0]][0]:
                    if nb not in visited:
                        cands.add((nb, distance(q, self.points[nb]))); evals += 1
                        visited.add(nb)
        pops = len(result)
        insertion_sort(result, lambda a, b: distance(q, self.points[a]) < distance(q, self.points[b]))
        return (np.array(result, np.uint32), np.array([distance(q, self.points[i]) for i in result], self.T), pops, evals)

    def layer(self, layer):
        """Padded adjacency [n, m] (0xFFFFFFFF) of one layer, like OracleHNSW.export_layer."""
        adj = np.full((len(self.points), self.m), 0xFFFFFFFF, np.uint32)
        for i, c in enumerate(self.conn):
            if layer < len(c):
                adj[i, : len(c[layer])] = c[layer]
        return adj

    def descend(self, query):
        """K2 (extension): insert's greedy walk (hnsw.zig:89-104) taken top down over layers max_level..1 from the
        first node that reached max_level -> (node, distance, evaluations)."""
        q = np.asarray(query, self.T)
        levels = [len(c) - 1 for c in self.conn]
        ep = levels.index(max(levels))
        curr, ev = distance(q, self.points[ep]), 1
        for layer in range(max(levels), 0, -1):
            changed = True
            while changed:
                changed = False
                cur_conn = self.conn[ep]
                if layer < len(cur_conn):
                    for nb in list(cur_conn[layer]):
                        d = distance(q, self.points[nb]); ev += 1
                        if d < curr:
                            ep, curr, changed = nb, d, True
        return ep, curr, ev
