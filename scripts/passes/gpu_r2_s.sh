#!/bin/bash
# round 2, pass S (1 GPU): default bench line with the small_batch leg; K1L + ABI tests
mkdir -p gpurun_out; rm -f gpurun_out/r02s_*
timeout 600 python -m pytest tests/test_gpu_team_kernel.py tests/test_abi.py -q 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02s_bench_n1.json 2> gpurun_out/r02s_bench_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/r02s_bench_n1.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02s_bench_n1.json").read().strip().splitlines()[-1])
print(d["value"], d["e2e"]["value"], d["parity"]["ids"], d["parity"]["dist_bits"])
print(json.dumps(d.get("small_batch")))
PY
