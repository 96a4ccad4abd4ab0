#!/bin/bash
# round 2, pass Q (1 GPU): K1L A/B between two builds of the library (ZVDB_B200_LIB), nq in {1, 64, 296}
mkdir -p gpurun_out; rm -f gpurun_out/r02q_*
C5_NQ=1,64,296 timeout 600 python scripts/c5_team_sweep.py ${2:-both} > gpurun_out/r02q_head.jsonl 2> gpurun_out/r02q_head.err; echo "head rc=$?"
ZVDB_B200_LIB=$PWD/$1 C5_NQ=1,64,296 timeout 600 python scripts/c5_team_sweep.py ${2:-both} > gpurun_out/r02q_alt.jsonl 2> gpurun_out/r02q_alt.err; echo "alt rc=$?"; tail -2 gpurun_out/r02q_alt.err
for f in head alt; do echo "== $f"; python - <<PY
import json
for l in open("gpurun_out/r02q_$f.jsonl"):
    r = json.loads(l); print(r["graph"], r["ef"], r["nq"], r["k1_device_ms"], r["k1l_device_ms"], r["rows_equal"])
PY
done
