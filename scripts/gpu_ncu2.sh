#!/bin/bash
# ncu evidence, second pass: launch list of the default bench command, --set full captures of K1 (both graphs) and K4
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bench_default.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/ncu_launch_bench.json 2> gpurun_out/ncu_launch_bench.err; echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:search_layer0 -s 6 -c 1 -o gpurun_out/prof2_k1_ref_ef64 -f \
    python bench.py --steps 3 --warmup 3 --no-cpu > /dev/null 2> gpurun_out/ncu_full_ref.err; echo "full ref rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:search_layer0 -s 6 -c 1 -o gpurun_out/prof2_k1_ref_ef512 -f \
    python bench.py --steps 3 --warmup 3 --ef 512 --no-cpu > /dev/null 2> gpurun_out/ncu_full_ref512.err; echo "full ref512 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:search_layer0 -s 6 -c 1 -o gpurun_out/prof2_k1_q_ef128 -f \
    python bench.py --steps 3 --warmup 3 --ef 128 --graph quality --no-cpu > /dev/null 2> gpurun_out/ncu_full_q.err; echo "full q rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bf_gemm_topk -s 1 -c 1 -o gpurun_out/prof2_k4_128 -f \
    python scripts/bench_bruteforce.py --steps 1 > /dev/null 2> gpurun_out/ncu_full_k4.err; echo "full k4 rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/launches_k4.csv \
    python scripts/bench_bruteforce.py --steps 2 > gpurun_out/ncu_launch_k4.json 2> gpurun_out/ncu_launch_k4.err; echo "k4 launch list rc=$?"
timeout 300 python scripts/bench_bruteforce.py --dim 768 --k 100 --steps 1 > gpurun_out/bench_k4_768_k100.json 2>&1; cat gpurun_out/bench_k4_768_k100.json
timeout 300 python scripts/bench_bruteforce.py --dim 768 --k 10 --steps 2 > gpurun_out/bench_k4_768_k10.json 2>&1; cat gpurun_out/bench_k4_768_k10.json
timeout 300 python scripts/bench_bruteforce.py --dim 128 --k 100 --steps 2 > gpurun_out/bench_k4_128_k100.json 2>&1; cat gpurun_out/bench_k4_128_k100.json
ls -la gpurun_out
