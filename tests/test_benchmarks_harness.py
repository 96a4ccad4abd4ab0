"""The reference's benchmark harness (benchmarks/shared_benchmarks.zig) mirrored in
zvdb_b200/benchmarks.py: result formatting on CPU, a small end-to-end run on the GPU."""
import io

import pytest


def test_benchmark_result_text_and_csv_follow_the_reference():
    from zvdb_b200.benchmarks import BenchmarkResult
    # the one measured result the reference publishes (benchmarks/benchmark.md:107-113)
    r = BenchmarkResult("Search", 100000, 128, 10000, 10, None, 3_733_950_000, 2678.13)
    assert r.format() == ("Search Benchmark:\n  Points: 100000\n  Dimensions: 128\n  Queries: 10000\n  k: 10\n"
                          "  Total time: 3.73 seconds\n  Search per second: 2678.13\n")
    assert r.to_csv() == "Search,100000,128,10000,10,1,3733950000,2678.13"      # threads default to 1, :46
    i = BenchmarkResult("Insertion", 100000, 128, None, None, 4, 11_920_000_000, 8392.22)
    assert i.format() == ("Insertion Benchmark:\n  Points: 100000\n  Dimensions: 128\n  Threads: 4\n"
                          "  Total time: 11.92 seconds\n  Insertion per second: 8392.22\n")
    assert i.to_csv() == "Insertion,100000,128,0,0,4,11920000000,8392.22"


def test_benchmark_config_defaults_are_the_reference_sweep():
    from zvdb_b200.benchmarks import BenchmarkConfig, random_point
    c = BenchmarkConfig()
    assert (c.num_points, tuple(c.dimensions), c.num_queries, tuple(c.k_values)) == \
        (100000, (128, 512, 768, 1024), 10000, (10, 25, 50, 100))      # single_threaded_benchmarks.zig:28-33
    p = random_point(64)
    assert p.dtype.name == "float32" and p.shape == (64,) and 0.0 <= p.min() and p.max() < 1.0


@pytest.mark.gpu
def test_benchmark_harness_runs(zv):
    from zvdb_b200 import benchmarks as B
    cfg = B.BenchmarkConfig(2000, (32,), 50, (5,))
    out = io.StringIO()
    res = B.run_single_threaded_benchmarks(cfg, out=out)
    assert [r.operation for r in res] == ["Insertion", "Search"]
    assert res[1].num_queries == 50 and res[1].k == 5 and res[1].operations_per_second > 0
    assert out.getvalue().startswith("Running Single-Threaded Benchmarks\n================================\n\nInsertion Benchmark:\n")
    out = io.StringIO()
    res = B.run_multi_threaded_benchmarks(B.BenchmarkConfig(500, (16,), 20, (3,)), out=out, csv=True, batched=True)
    assert len(res) == 6 and [r.num_threads for r in res] == [2, 2, 4, 4, 8, 8]
    assert all(line.count(",") == 7 for line in out.getvalue().strip().splitlines())
