#!/bin/bash
# incremental builder at 1M (build-ef sweep) and at C4 shard scale (12.5M x 128, one shard)
mkdir -p gpurun_out
timeout 600 python scripts/builder_incremental_eval.py 1000000 1 > gpurun_out/builder_incremental_1m.jsonl 2> gpurun_out/builder_incremental_1m.err; echo "1M rc=$?"; cut -c1-330 gpurun_out/builder_incremental_1m.jsonl | grep '"ef": 512'
free -g | head -2
ZVDB_BUILD_EF=256 timeout 1200 python scripts/builder_incremental_eval.py 12500000 1 > gpurun_out/builder_incremental_12m.jsonl 2> gpurun_out/builder_incremental_12m.err; echo "12.5M rc=$?"; cut -c1-420 gpurun_out/builder_incremental_12m.jsonl; tail -4 gpurun_out/builder_incremental_12m.err
