#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_bruteforce.py tests/test_gpu_host_buffers.py -x -q > gpurun_out/pytest_k4.log 2>&1; echo "pytest k4 rc=$?"; tail -8 gpurun_out/pytest_k4.log
timeout 120 python scripts/bench_bruteforce.py > gpurun_out/bench_k4.json 2> gpurun_out/bench_k4.err; echo "bench k4 rc=$?"; cat gpurun_out/bench_k4.json; tail -5 gpurun_out/bench_k4.err
timeout 120 python scripts/bench_bruteforce.py --filter > gpurun_out/bench_k4_filter.json 2>&1; cat gpurun_out/bench_k4_filter.json
timeout 120 python scripts/bench_bruteforce.py --filter --nq 65536 --steps 2 > gpurun_out/bench_k4_filter_64k.json 2>&1; cat gpurun_out/bench_k4_filter_64k.json
timeout 120 python scripts/bench_bruteforce.py --dim 768 --k 10 --steps 2 > gpurun_out/bench_k4_768_k10.json 2>&1; cat gpurun_out/bench_k4_768_k10.json
timeout 120 python scripts/bench_bruteforce.py --dim 768 --k 100 --steps 2 > gpurun_out/bench_k4_768_k100.json 2>&1; cat gpurun_out/bench_k4_768_k100.json
timeout 120 python scripts/bench_bruteforce.py --dim 128 --k 10 --nq 65536 --steps 2 > gpurun_out/bench_k4_128_nq64k.json 2>&1; cat gpurun_out/bench_k4_128_nq64k.json
