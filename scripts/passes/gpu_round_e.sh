#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 --sweep > gpurun_out/bench_ref_e.json 2> gpurun_out/bench_ref_e.err; echo "bench ref rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/bench_ref_e.json'));print(d['value'],d['roofline']['frac'],d['e2e']['value'],d['cpu_baseline'],d['clocks']);[print(x) for x in d['sweep']]"
timeout 600 python bench.py --steps 10 --warmup 3 --graph quality --ef 128 --sweep > gpurun_out/bench_q_e.json 2> gpurun_out/bench_q_e.err; echo "bench q rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/bench_q_e.json'));print(d['value'],d['roofline']['frac'],d['e2e']['value']);[print(x) for x in d['sweep']]"
timeout 600 python bench.py --steps 20 --warmup 3 --descent > gpurun_out/bench_ref_descent.json 2> gpurun_out/bench_ref_descent.err; echo "bench descent rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/bench_ref_descent.json'));print(d['value'],d['config'],d['cpu_baseline'])"
