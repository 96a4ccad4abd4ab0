#!/bin/bash
# gate for the mapped small-batch path: all GPU tests, smoke, the reference harness at 128-d, default bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 600 python -m zvdb_b200.benchmarks single > gpurun_out/reference_harness_single.txt 2> gpurun_out/reference_harness_single.err; echo "harness rc=$?"; grep -E "Dimensions|  k:|per second" gpurun_out/reference_harness_single.txt | paste - - - - | head -30
timeout 600 python -m zvdb_b200.benchmarks single --batched --dims 128 --ks 10 > gpurun_out/reference_harness_batched.txt 2>> gpurun_out/reference_harness_single.err; echo "harness batched rc=$?"
timeout 600 python bench.py --steps 20 --warmup 3 --graph quality --ef 128 > gpurun_out/bench_q_m.json 2> gpurun_out/bench_q_m.err; echo "bench q rc=$?"; python -c "import json; d=json.load(open('gpurun_out/bench_q_m.json')); print(d['value'], d['e2e'], d['cpu_baseline'])"
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_m.json 2> gpurun_out/bench_m.err; echo "bench rc=$?"; python -c "import json; d=json.load(open('gpurun_out/bench_m.json')); print(d['value'], d['e2e'], d['cpu_baseline'])"
timeout 600 python scripts/configs_c3_c5.py c5 > gpurun_out/c5.jsonl 2> gpurun_out/c5.err; echo "c5 rc=$?"; grep '"reference"' gpurun_out/c5.jsonl | cut -c1-260 | head -8
