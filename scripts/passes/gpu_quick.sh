#!/bin/bash
# quick pass: GPU tests (optionally a subset via $1) + default bench line
mkdir -p gpurun_out
timeout 1200 python -m pytest ${1:-tests} -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench rc=$?"; cut -c1-900 gpurun_out/bench_quick.json; tail -3 gpurun_out/bench_quick.err
