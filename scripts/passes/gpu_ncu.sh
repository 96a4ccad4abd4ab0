#!/bin/bash
# ncu evidence for K1: launch list of the bench command + one --set full capture per configuration
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_ref_ef64.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/ncu_launch_bench.json 2> gpurun_out/ncu_launch_bench.err; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:search_layer0 -s 6 -c 2 -o gpurun_out/prof_k1_ref_ef64 -f \
    python bench.py --steps 3 --warmup 3 --no-cpu > /dev/null 2> gpurun_out/ncu_full_ref.err; echo "full ref rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:search_layer0 -s 6 -c 1 -o gpurun_out/prof_k1_ref_ef512 -f \
    python bench.py --steps 3 --warmup 3 --ef 512 --no-cpu > /dev/null 2> gpurun_out/ncu_full_ref512.err; echo "full ref512 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:search_layer0 -s 6 -c 1 -o gpurun_out/prof_k1_q_ef128 -f \
    python bench.py --steps 3 --warmup 3 --ef 128 --graph quality --no-cpu > /dev/null 2> gpurun_out/ncu_full_q.err; echo "full q rc=$?"
ls -la gpurun_out
