#!/bin/bash
# 2-GPU pass: sharded tests, then bench with both exchange formulations (launched like the driver does)
mkdir -p gpurun_out
N=${1:-2}
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_sharded.py -x -q > gpurun_out/pytest_sharded.log 2>&1; echo "pytest sharded rc=$?"; tail -15 gpurun_out/pytest_sharded.log
for ex in p2p nccl; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 --exchange $ex > gpurun_out/bench_n${N}_${ex}.json 2> gpurun_out/bench_n${N}_${ex}.err; echo "bench $ex rc=$?"; cat gpurun_out/bench_n${N}_${ex}.json; tail -4 gpurun_out/bench_n${N}_${ex}.err
done
