#!/bin/bash
# round 2, pass B (2 GPUs): GPU tests incl. the two-GPU sharded parity test, N=2 bench lines (fused one-launch step vs round 1's three launches),
# the reference arm under torchrun
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r02_b_tests.log
tail -3 gpurun_out/r02_b_tests.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_b_n2_p2p.json 2> gpurun_out/r02_b_n2_p2p.err
tail -2 gpurun_out/r02_b_n2_p2p.err
$TR bench.py --gpus 2 --steps 20 --warmup 5 --exchange p2p3 --parity-queries 0 > gpurun_out/r02_b_n2_p2p3.json 2> gpurun_out/r02_b_n2_p2p3.err
$TR bench.py --gpus 2 --steps 20 --warmup 5 --exchange nccl --parity-queries 0 > gpurun_out/r02_b_n2_nccl.json 2> gpurun_out/r02_b_n2_nccl.err
$TR bench.py --impl reference --gpus 2 --steps 5 --warmup 1 > gpurun_out/r02_b_n2_ref.json 2> gpurun_out/r02_b_n2_ref.err
tail -1 gpurun_out/r02_b_n2_ref.err
