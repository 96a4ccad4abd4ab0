#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 --sweep > gpurun_out/bench_ref_f.json 2> gpurun_out/bench_ref_f.err; echo "bench ref rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/bench_ref_f.json'));print(d['value'],d['roofline']['frac'],d['e2e']['value']);[print(x) for x in d['sweep']]"
timeout 600 python bench.py --steps 10 --warmup 3 --graph quality --ef 128 --sweep > gpurun_out/bench_q_f.json 2> gpurun_out/bench_q_f.err; echo "bench q rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/bench_q_f.json'));print(d['value'],d['roofline']['frac'],d['e2e']['value']);[print(x) for x in d['sweep']]"
timeout 600 python scripts/configs_c3_c5.py c5 > gpurun_out/c5.jsonl 2> gpurun_out/c5.err; echo "c5 rc=$?"; python -c "
import json
for l in open('gpurun_out/c5.jsonl'):
    d=json.loads(l); print(d['graph'],d['nq'],round(d['device_ms'],3),int(d['device_qps']),round(d['host_call_ms'],3),int(d['host_call_qps']))"
