#!/bin/bash
# search-driven incremental builder: its test, then exact vs incremental at 1M x 128
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "incremental or builder" > gpurun_out/pytest_builder.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_builder.log
timeout 900 python scripts/builder_incremental_eval.py ${1:-1000000} ${2:-0} > gpurun_out/builder_incremental.jsonl 2> gpurun_out/builder_incremental.err; echo "eval rc=$?"; cat gpurun_out/builder_incremental.jsonl | cut -c1-400; tail -5 gpurun_out/builder_incremental.err
