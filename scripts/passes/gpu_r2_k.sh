#!/bin/bash
# round 2, pass K (1 GPU): C4 strong-scaling baseline on the incremental graph: the eight 12.5M-row shards of the 100M-row index held and searched by ONE GPU
mkdir -p gpurun_out; rm -f gpurun_out/r02k_*
free -g | head -2 > gpurun_out/r02k_host.txt
timeout 1500 python scripts/c4_shards_one_gpu.py --rows 100000000 --shards 8 --graph incremental --steps 10 --warmup 3 > gpurun_out/r02k_c4_n1_incremental.jsonl 2> gpurun_out/r02k_c4_n1_incremental.err; echo "rc=$?"
grep -v builder gpurun_out/r02k_c4_n1_incremental.err | tail -4; cut -c1-260 gpurun_out/r02k_c4_n1_incremental.jsonl
