#!/bin/bash
# 2-GPU check of the sliced e2e input path (driver-style launch)
mkdir -p gpurun_out
timeout 80 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_n2_sliced.json 2> gpurun_out/bench_n2_sliced.err; echo "bench N=2 rc=$?"; python -c "import json; d=json.load(open('gpurun_out/bench_n2_sliced.json')); print(d['value'], d['ms_per_step'], d['e2e'])"; tail -3 gpurun_out/bench_n2_sliced.err
