#!/bin/bash
# re-entry pass: GPU parity tests, smoke, bench on both graphs (+sweeps), reference arm, C5 and C3 full-size runs
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
timeout 600 python bench.py --steps 20 --warmup 3 --sweep > gpurun_out/bench_ref_c.json 2> gpurun_out/bench_ref_c.err; echo "bench ref rc=$?"; cat gpurun_out/bench_ref_c.json; tail -3 gpurun_out/bench_ref_c.err
timeout 600 python bench.py --steps 10 --warmup 3 --graph quality --ef 128 --sweep > gpurun_out/bench_q_c.json 2> gpurun_out/bench_q_c.err; echo "bench q rc=$?"; cat gpurun_out/bench_q_c.json; tail -3 gpurun_out/bench_q_c.err
timeout 600 python scripts/configs_c3_c5.py c5 > gpurun_out/c5.jsonl 2> gpurun_out/c5.err; echo "c5 rc=$?"; tail -40 gpurun_out/c5.jsonl; tail -3 gpurun_out/c5.err
timeout 900 python scripts/configs_c3_c5.py c3 > gpurun_out/c3.jsonl 2> gpurun_out/c3.err; echo "c3 rc=$?"; cat gpurun_out/c3.jsonl; tail -3 gpurun_out/c3.err
