#!/bin/bash
# re-entry pass: GPU tests, smoke, default bench line, reference arm
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_h.json 2> gpurun_out/bench_h.err; echo "bench rc=$?"; cut -c1-1200 gpurun_out/bench_h.json; tail -3 gpurun_out/bench_h.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_refarm_h.json 2> gpurun_out/bench_refarm_h.err; echo "refarm rc=$?"; cut -c1-600 gpurun_out/bench_refarm_h.json
