#!/bin/bash
# round 2, pass H (2 GPUs): sharded tests and N=2 bench after the tournament merge
mkdir -p gpurun_out; rm -f gpurun_out/r02h_*
python -m pytest tests/test_gpu_sharded.py tests/test_gpu_parity.py tests/test_gpu_metrics.py -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r02h_tests.log; tail -3 gpurun_out/r02h_tests.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02h_n2.json 2> gpurun_out/r02h_n2.err; tail -2 gpurun_out/r02h_n2.err
cut -c1-330 gpurun_out/r02h_n2.json
