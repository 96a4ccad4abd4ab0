#!/bin/bash
# round 2, pass U (2 GPUs): the final tree at N = 2: the sharded GPU tests (the two-rank test needs 2 devices), the N = 2 bench line and its reference arm
mkdir -p gpurun_out; rm -f gpurun_out/r02u_*
timeout 600 python -m pytest tests/test_gpu_sharded.py -m gpu -q 2>&1 | tail -4 > gpurun_out/r02u_tests.log; tail -2 gpurun_out/r02u_tests.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02u_n2.json 2> gpurun_out/r02u_n2.err; echo "bench N=2 rc=$?"; tail -1 gpurun_out/r02u_n2.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02u_n2.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["parity"])
PY
