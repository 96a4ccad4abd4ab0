#!/bin/bash
# zero-copy host-buffer path: host-buffer + parity tests, default bench line (e2e zero-copy vs staged), N=1 sweep
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_host_buffers.py tests/test_gpu_parity.py tests/test_descent.py -m gpu -x -q > gpurun_out/pytest_parity.log 2>&1; echo "parity rc=$?"; tail -3 gpurun_out/pytest_parity.log
timeout 600 python bench.py --steps 20 --warmup 3 --sweep > gpurun_out/bench_k.json 2> gpurun_out/bench_k.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_k.json'))
print({k:d[k] for k in ('value','ms_per_step','e2e','roofline','cpu_baseline','clocks')})
print(d.get('sweep'))
PY
tail -3 gpurun_out/bench_k.err
