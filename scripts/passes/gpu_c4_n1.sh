#!/bin/bash
# C4 baseline: the whole 100M x 128 index on ONE GPU (reference insert: a serial host build, ~15 min), ef sweep
mkdir -p gpurun_out
timeout 2300 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu --no-recall --rows ${1:-100000000} --ef 512 --shard-gen --sweep > gpurun_out/bench_c4_ef512_n1.json 2> gpurun_out/bench_c4_ef512_n1.err; echo "rc=$?"
cut -c1-1500 gpurun_out/bench_c4_ef512_n1.json; tail -3 gpurun_out/bench_c4_ef512_n1.err; nvidia-smi --query-gpu=memory.used --format=csv
