#!/bin/bash
# round 2, pass W (1 GPU): final tree: all GPU tests, smoke, default bench line, C5 batch sweep in automatic mode
mkdir -p gpurun_out; rm -f gpurun_out/r02w_*
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r02w_tests.log; tail -2 gpurun_out/r02w_tests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02w_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r02w_smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02w_bench_n1.json 2> gpurun_out/r02w_bench_n1.err; echo "bench rc=$?"
timeout 900 python scripts/configs_c3_c5.py c5 > gpurun_out/r02w_c5_batch_sweep.jsonl 2> gpurun_out/r02w_c5_batch_sweep.err; echo "c5 rc=$?"
python - <<'PY'
import json
for l in open("gpurun_out/r02w_c5_batch_sweep.jsonl"):
    r = json.loads(l); print(r["graph"], r["nq"], round(r["device_ms"], 4), round(r["host_call_ms"], 4))
PY
cut -c1-200 gpurun_out/r02w_bench_n1.json
