#!/bin/bash
# round 2, pass Z (1 GPU): single-query completion mailbox: GPU tests that make single calls, then the reference's harness loop from a compiled C caller
mkdir -p gpurun_out; rm -f gpurun_out/r02z_*
timeout 600 python -m pytest tests/test_gpu_team_kernel.py tests/test_gpu_parity.py tests/test_gpu_dtypes.py tests/test_gpu_host_buffers.py tests/test_benchmarks_harness.py tests/test_gpu_persistence.py -m gpu -q 2>&1 | tail -3
gcc -O2 -std=c99 -Iinclude integration/harness.c -Lzvdb_b200/lib -lzvdb_b200 -Wl,-rpath,$PWD/zvdb_b200/lib -o gpurun_out/r02z_harness && \
  for d in 128 1024; do for k in 10 100; do gpurun_out/r02z_harness 100000 $d 10000 $k; done; done > gpurun_out/r02z_c_harness.txt 2>&1; rm -f gpurun_out/r02z_harness
grep "Search per second\|per call" gpurun_out/r02z_c_harness.txt
