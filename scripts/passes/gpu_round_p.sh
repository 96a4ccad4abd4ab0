#!/bin/bash
# final verification of the tree: all GPU tests, smoke, default bench line, incremental-graph bench line, reference arm
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_p.json 2> gpurun_out/bench_p.err; echo "bench rc=$?"; python -c "import json; d=json.load(open('gpurun_out/bench_p.json')); print(d['value'], d['ms_per_step'], d['e2e'], d['roofline']['frac'], d['cpu_baseline'], d['clocks'])"
timeout 600 python bench.py --steps 10 --warmup 3 --graph incremental --ef 128 --sweep > gpurun_out/bench_inc_p.json 2> gpurun_out/bench_inc_p.err; echo "bench inc rc=$?"; python -c "import json; d=json.load(open('gpurun_out/bench_inc_p.json')); print(d['value'], d['config'], d['e2e'], d['roofline']['frac'], d.get('sweep'))"; tail -2 gpurun_out/bench_inc_p.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_refarm_p.json 2> gpurun_out/bench_refarm_p.err; echo "refarm rc=$?"; cut -c1-300 gpurun_out/bench_refarm_p.json
