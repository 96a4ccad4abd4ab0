#!/bin/bash
# first GPU pass: parity tests, smoke, bench on both graphs
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 --sweep > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref rc=$?"; cat gpurun_out/bench_ref.json; tail -5 gpurun_out/bench_ref.err
timeout 900 python bench.py --steps 10 --warmup 3 --graph quality --ef 128 --sweep > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err; echo "bench q rc=$?"; cat gpurun_out/bench_q.json; tail -5 gpurun_out/bench_q.err
