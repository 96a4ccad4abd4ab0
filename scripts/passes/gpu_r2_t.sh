#!/bin/bash
# round 2, pass T (1 GPU): evidence pass of the final tree (with K1L): all GPU tests, smoke, ncu launch list of the default bench command,
# the reference's own benchmark harness (dims x k sweep) through the Python mirror and through a compiled C caller
mkdir -p gpurun_out; rm -f gpurun_out/r02t_*
nvidia-smi --query-gpu=name,memory.total,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02t_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r02t_tests.log; tail -2 gpurun_out/r02t_tests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02t_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r02t_smoke.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02t_launches_bench_default.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu > /dev/null 2> gpurun_out/r02t_ncu_launch.err; echo "launch list rc=$?"
timeout 600 python -m zvdb_b200.benchmarks single > gpurun_out/r02t_reference_harness_sweep.txt 2> gpurun_out/r02t_reference_harness_sweep.err; echo "python harness rc=$?"
gcc -O2 -std=c99 -Iinclude integration/harness.c -Lzvdb_b200/lib -lzvdb_b200 -Wl,-rpath,$PWD/zvdb_b200/lib -o gpurun_out/r02t_harness && \
  for d in 128 512 768 1024; do for k in 10 25 50 100; do gpurun_out/r02t_harness 100000 $d 10000 $k; done; done > gpurun_out/r02t_c_harness_sweep.txt 2>&1; rm -f gpurun_out/r02t_harness
grep -c "search_team_kernel" gpurun_out/r02t_launches_bench_default.csv; grep -c "search_layer0_kernel" gpurun_out/r02t_launches_bench_default.csv
grep "Search per second" gpurun_out/r02t_c_harness_sweep.txt | tr '\n' ' '
