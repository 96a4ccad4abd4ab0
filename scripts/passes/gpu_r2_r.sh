#!/bin/bash
# round 2, pass R (1 GPU): the tree with K1L: all GPU tests, smoke, default bench line, C5 A/B sweep K1 vs K1L, C5 batch sweep in automatic mode,
# the reference's harness loop from a compiled C caller
mkdir -p gpurun_out; rm -f gpurun_out/r02r_*
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r02r_tests.log; tail -2 gpurun_out/r02r_tests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02r_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r02r_smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02r_bench_n1.json 2> gpurun_out/r02r_bench_n1.err; echo "bench rc=$?"
timeout 900 python scripts/c5_team_sweep.py both > gpurun_out/r02r_c5_team_sweep.jsonl 2> gpurun_out/r02r_c5_team_sweep.err; echo "team sweep rc=$?"
timeout 900 python scripts/configs_c3_c5.py c5 > gpurun_out/r02r_c5_batch_sweep.jsonl 2> gpurun_out/r02r_c5_batch_sweep.err; echo "c5 rc=$?"
gcc -O2 -std=c99 -Iinclude integration/harness.c -Lzvdb_b200/lib -lzvdb_b200 -Wl,-rpath,$PWD/zvdb_b200/lib -o gpurun_out/r02r_harness && \
  for k in 10 100; do gpurun_out/r02r_harness 100000 128 10000 $k; done > gpurun_out/r02r_c_harness.txt 2>&1; rm -f gpurun_out/r02r_harness; grep -i "per second\|per call" gpurun_out/r02r_c_harness.txt
cut -c1-300 gpurun_out/r02r_bench_n1.json
