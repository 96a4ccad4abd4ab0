#!/bin/bash
# C4 strong-scaling baseline: the 8-shard 100M x 128 index of the 8-GPU run, held and searched by ONE GPU
mkdir -p gpurun_out
free -g | head -2; nproc
timeout ${2:-1100} python scripts/c4_shards_one_gpu.py --rows ${1:-100000000} --shards 8 --min-free-gb ${3:-120} > gpurun_out/c4_shards8_n1.jsonl 2> gpurun_out/c4_shards8_n1.err; echo "rc=$?"
cut -c1-700 gpurun_out/c4_shards8_n1.jsonl; tail -3 gpurun_out/c4_shards8_n1.err; nvidia-smi --query-gpu=memory.used --format=csv
