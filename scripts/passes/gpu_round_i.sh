#!/bin/bash
# K1 change gate: parity tests first, then the L2-prefetch A/B (1M x 128, both graphs)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_descent.py -m gpu -x -q > gpurun_out/pytest_parity.log 2>&1; echo "parity rc=$?"; tail -3 gpurun_out/pytest_parity.log
timeout 600 python scripts/sweep_prefetch.py $1 > gpurun_out/prefetch_ab.jsonl 2> gpurun_out/prefetch_ab.err; echo "ab rc=$?"; cat gpurun_out/prefetch_ab.jsonl; tail -3 gpurun_out/prefetch_ab.err
