"""The reference's benchmark harness over the B200 index.

Mirrors benchmarks/shared_benchmarks.zig: `BenchmarkResult` with the same text block (`format`,
:14-37) and CSV row (`toCsv`, :39-50), `randomPoint` (:53-59, uniform [0,1)), `runInsertionBenchmark`
(:61-88) and `runSearchBenchmark` (:90-125) with the reference's timed regions -- point/query
generation INSIDE the loop, one `insert` / `search(query, k)` call per iteration, index built with
untimed inserts before the search loop -- and the two mains' sweep (dims {128,512,768,1024} x k
{10,25,50,100}, 100 000 points, 10 000 queries; single_threaded_benchmarks.zig:28-33). As in the
reference, "threads" is only a label on the result (multi_threaded_benchmarks.zig spawns none).

`run_search_benchmark_batched` is the same measurement through the batch call (one kernel launch
for all queries), which is how this index is meant to be driven.

    python -m zvdb_b200.benchmarks single [--points N] [--queries Q] [--dims 128,512] [--ks 10,25] [--csv]
"""
from __future__ import annotations

import argparse
import sys
import time
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np


@dataclass
class BenchmarkResult:                       # shared_benchmarks.zig:4-12
    operation: str
    num_points: int
    dimensions: int
    num_queries: Optional[int]
    k: Optional[int]
    num_threads: Optional[int]
    total_time_ns: int
    operations_per_second: float

    def format(self) -> str:                 # shared_benchmarks.zig:14-37
        out = [f"{self.operation} Benchmark:", f"  Points: {self.num_points}", f"  Dimensions: {self.dimensions}"]
        if self.num_queries is not None:
            out.append(f"  Queries: {self.num_queries}")
        if self.k is not None:
            out.append(f"  k: {self.k}")
        if self.num_threads is not None:
            out.append(f"  Threads: {self.num_threads}")
        out.append(f"  Total time: {self.total_time_ns / 1e9:.2f} seconds")
        out.append(f"  {self.operation} per second: {self.operations_per_second:.2f}")
        return "\n".join(out) + "\n"

    __str__ = format

    def to_csv(self) -> str:                 # shared_benchmarks.zig:39-50
        return (f"{self.operation},{self.num_points},{self.dimensions},{self.num_queries or 0},{self.k or 0},"
                f"{self.num_threads or 1},{self.total_time_ns},{self.operations_per_second:.2f}")


@dataclass
class BenchmarkConfig:                       # shared_benchmarks.zig:127-132
    num_points: int = 100000
    dimensions: Sequence[int] = (128, 512, 768, 1024)
    num_queries: int = 10000
    k_values: Sequence[int] = (10, 25, 50, 100)


_rng = np.random.default_rng()               # unseeded, like std.crypto.random


def random_point(dim: int) -> np.ndarray:    # shared_benchmarks.zig:53-59
    return _rng.random(dim, dtype=np.float32)


def _new_index():
    from .hnsw import HNSW
    return HNSW(16, 200)                     # HNSW(f32).init(allocator, 16, 200), shared_benchmarks.zig:62,91


def run_insertion_benchmark(num_points: int, dim: int, num_threads: Optional[int] = None) -> BenchmarkResult:
    hnsw = _new_index()
    try:
        start = time.perf_counter_ns()
        for _ in range(num_points):          # :68-72: generation is inside the timed region
            hnsw.insert(random_point(dim))
        elapsed = time.perf_counter_ns() - start
    finally:
        hnsw.deinit()
    return BenchmarkResult("Insertion", num_points, dim, None, None, num_threads, elapsed, num_points / (elapsed / 1e9))


def run_search_benchmark(num_points: int, dim: int, num_queries: int, k: int,
                         num_threads: Optional[int] = None) -> BenchmarkResult:
    hnsw = _new_index()
    try:
        hnsw.insert_batch(_rng.random((num_points, dim), dtype=np.float32))      # untimed, :95-99
        hnsw.sync_device()
        start = time.perf_counter_ns()
        for _ in range(num_queries):         # :104-109: query generation and result hand-over inside the timed region
            results = hnsw.search(random_point(dim), k)
            del results
        elapsed = time.perf_counter_ns() - start
    finally:
        hnsw.deinit()
    return BenchmarkResult("Search", num_points, dim, num_queries, k, num_threads, elapsed, num_queries / (elapsed / 1e9))


def run_search_benchmark_batched(num_points: int, dim: int, num_queries: int, k: int,
                                 num_threads: Optional[int] = None) -> BenchmarkResult:
    """Same workload, all queries in ONE zvdb_search_batch call (host buffers, copies inside)."""
    hnsw = _new_index()
    try:
        hnsw.insert_batch(_rng.random((num_points, dim), dtype=np.float32))
        hnsw.sync_device()
        hnsw.search_batch(_rng.random((num_queries, dim), dtype=np.float32), k, k)   # sizes the device buffers, untimed
        start = time.perf_counter_ns()
        queries = _rng.random((num_queries, dim), dtype=np.float32)
        hnsw.search_batch(queries, k, k)
        elapsed = time.perf_counter_ns() - start
    finally:
        hnsw.deinit()
    return BenchmarkResult("Search", num_points, dim, num_queries, k, num_threads, elapsed, num_queries / (elapsed / 1e9))


def run_single_threaded_benchmarks(config: BenchmarkConfig, out=sys.stdout, csv: bool = False, batched: bool = False) -> list:
    """single_threaded_benchmarks.zig:4-21."""
    if not csv:
        out.write("Running Single-Threaded Benchmarks\n================================\n\n")
    search = run_search_benchmark_batched if batched else run_search_benchmark
    results = []
    for dim in config.dimensions:
        r = run_insertion_benchmark(config.num_points, dim, None)
        results.append(r)
        out.write((r.to_csv() if csv else r.format()) + "\n")
        for k in config.k_values:
            r = search(config.num_points, dim, config.num_queries, k, None)
            results.append(r)
            out.write((r.to_csv() if csv else r.format()) + "\n")
        if not csv:
            out.write("\n")
    return results


def run_multi_threaded_benchmarks(config: BenchmarkConfig, out=sys.stdout, csv: bool = False, batched: bool = False) -> list:
    """multi_threaded_benchmarks.zig:4-26: the same loops with a thread-count LABEL in {2, 4, 8}."""
    if not csv:
        out.write("Running Multi-Threaded Benchmarks\n================================\n\n")
    search = run_search_benchmark_batched if batched else run_search_benchmark
    results = []
    for dim in config.dimensions:
        for threads in (2, 4, 8):
            r = run_insertion_benchmark(config.num_points, dim, threads)
            results.append(r)
            out.write((r.to_csv() if csv else r.format()) + "\n")
            for k in config.k_values:
                r = search(config.num_points, dim, config.num_queries, k, threads)
                results.append(r)
                out.write((r.to_csv() if csv else r.format()) + "\n")
            if not csv:
                out.write("\n")
    return results


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("which", choices=["single", "multi"])
    ap.add_argument("--points", type=int, default=100000)
    ap.add_argument("--queries", type=int, default=10000)
    ap.add_argument("--dims", default="128,512,768,1024")
    ap.add_argument("--ks", default="10,25,50,100")
    ap.add_argument("--csv", action="store_true")
    ap.add_argument("--batched", action="store_true", help="one zvdb_search_batch call instead of a loop of search calls")
    a = ap.parse_args(argv)
    cfg = BenchmarkConfig(a.points, tuple(int(x) for x in a.dims.split(",")), a.queries, tuple(int(x) for x in a.ks.split(",")))
    fn = run_single_threaded_benchmarks if a.which == "single" else run_multi_threaded_benchmarks
    fn(cfg, csv=a.csv, batched=a.batched)
    return 0


if __name__ == "__main__":
    sys.exit(main())
