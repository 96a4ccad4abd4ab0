#!/bin/bash
# round 2, pass P (1 GPU): K1L experiment: scan shared by idle warps (default) vs every warp scans (ZVDB_TEAM_NOSHARE=1), both graphs, nq in {1, 64, 296}
mkdir -p gpurun_out; rm -f gpurun_out/r02p_*
timeout 300 python -m pytest tests/test_gpu_team_kernel.py -m gpu -q 2>&1 | tail -3
C5_NQ=1,64,296 timeout 600 python scripts/c5_team_sweep.py both > gpurun_out/r02p_share.jsonl 2> gpurun_out/r02p_share.err; echo "share rc=$?"
ZVDB_TEAM_NOSHARE=1 C5_NQ=1,64,296 timeout 600 python scripts/c5_team_sweep.py both > gpurun_out/r02p_noshare.jsonl 2> gpurun_out/r02p_noshare.err; echo "noshare rc=$?"
for f in share noshare; do echo "== $f"; python - <<PY
import json
for l in open("gpurun_out/r02p_$f.jsonl"):
    r = json.loads(l); print(r["graph"], r["ef"], r["nq"], r["k1_device_ms"], r["k1l_device_ms"], r["rows_equal"])
PY
done
