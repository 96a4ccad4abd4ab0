#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 --sweep > gpurun_out/bench_ref_b.json 2> gpurun_out/bench_ref_b.err; echo "bench rc=$?"; cat gpurun_out/bench_ref_b.json; tail -3 gpurun_out/bench_ref_b.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_refarm.json 2> gpurun_out/bench_refarm.err; echo "refarm rc=$?"; cat gpurun_out/bench_refarm.json
timeout 900 python scripts/recall_explore.py > gpurun_out/recall_explore.jsonl 2> gpurun_out/recall_explore.err; echo "explore rc=$?"; cat gpurun_out/recall_explore.jsonl; tail -3 gpurun_out/recall_explore.err
