#!/bin/bash
# 8-GPU pass: sharded tests, bench at N=8 (both exchanges) and N=4, launched like the driver does
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo8.txt 2>&1
timeout 300 python -m pytest tests/test_gpu_sharded.py -x -q > gpurun_out/pytest_sharded8.log 2>&1; echo "pytest sharded rc=$?"; tail -5 gpurun_out/pytest_sharded8.log
run() { # N exchange
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $1 --steps 20 --warmup 3 --exchange $2 > gpurun_out/bench_n$1_$2.json 2> gpurun_out/bench_n$1_$2.err; echo "bench N=$1 $2 rc=$?"; cut -c1-1500 gpurun_out/bench_n$1_$2.json; tail -3 gpurun_out/bench_n$1_$2.err
}
run 8 p2p
run 8 nccl
run 4 p2p
