#!/bin/bash
# 8-GPU sanity of the final tree, launched like the driver does: N=8 then N=4 (fused peer-store exchange)
mkdir -p gpurun_out
for N in 8 4; do
  timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n${N}_p2p.json 2> gpurun_out/bench_n${N}_p2p.err; echo "bench N=$N rc=$?"; python -c "import json; d=json.load(open('gpurun_out/bench_n${N}_p2p.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['avg_launch_ms'], d['config']['workload'])"; tail -2 gpurun_out/bench_n${N}_p2p.err
done
