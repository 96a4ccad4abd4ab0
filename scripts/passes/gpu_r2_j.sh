#!/bin/bash
# round 2, pass J (1 GPU): descent at C4-shard scale, C3 (i.i.d. Gaussian and embedding-like low-rank rows), C5 batch sweep, the reference's own harness
mkdir -p gpurun_out; rm -f gpurun_out/r02j_*
timeout 900 python scripts/c4_shard_descent.py > gpurun_out/r02j_c4_shard_descent.jsonl 2> gpurun_out/r02j_c4_shard_descent.err; echo "descent rc=$?"; cat gpurun_out/r02j_c4_shard_descent.jsonl | cut -c1-220
timeout 900 python scripts/configs_c3_c5.py c3 > gpurun_out/r02j_c3.jsonl 2> gpurun_out/r02j_c3.err; echo "c3 rc=$?"; cut -c1-330 gpurun_out/r02j_c3.jsonl
timeout 600 python scripts/configs_c3_c5.py c5 > gpurun_out/r02j_c5.jsonl 2> gpurun_out/r02j_c5.err; echo "c5 rc=$?"; grep -E '"nq": (1|64|1024|65536),' gpurun_out/r02j_c5.jsonl
timeout 600 python -m zvdb_b200.benchmarks single --dims 128 --ks 10 > gpurun_out/r02j_reference_harness.txt 2> gpurun_out/r02j_reference_harness.err; echo "harness rc=$?"; tail -12 gpurun_out/r02j_reference_harness.txt
