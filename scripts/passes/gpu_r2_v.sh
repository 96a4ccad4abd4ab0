#!/bin/bash
# round 2, pass V (1 GPU): K1L with half teams (128 threads): parity tests, then the C5 A/B sweep K1 / K1L-256 / K1L-128
mkdir -p gpurun_out; rm -f gpurun_out/r02v_*
timeout 600 python -m pytest tests/test_gpu_team_kernel.py tests/test_abi.py -q 2>&1 | tail -4 > gpurun_out/r02v_tests.log; tail -2 gpurun_out/r02v_tests.log
timeout 900 python scripts/c5_team_sweep.py ${1:-both} > gpurun_out/r02v_c5_team_sweep.jsonl 2> gpurun_out/r02v_c5_team_sweep.err; echo "sweep rc=$?"; tail -2 gpurun_out/r02v_c5_team_sweep.err
python - <<'PY'
import json
for l in open("gpurun_out/r02v_c5_team_sweep.jsonl"):
    r = json.loads(l); print(r["graph"], r["ef"], r["nq"], r["k1_device_ms"], r["k1l_device_ms"], r["k1l_half_device_ms"], r["rows_equal"])
PY
