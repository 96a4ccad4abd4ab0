#!/bin/bash
# round 2, 8 GPUs, second pass: the default line with the record-form exchange + tournament merge, and the block + flag form next to it
mkdir -p gpurun_out; rm -f gpurun_out/r02n8b_*
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513"
timeout 600 $TR bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02n8b_default.json 2> gpurun_out/r02n8b_default.err; echo "default rc=$?"; tail -1 gpurun_out/r02n8b_default.err
timeout 600 $TR bench.py --gpus 8 --steps 20 --warmup 5 --exchange p2pb --parity-queries 0 > gpurun_out/r02n8b_p2pb.json 2> gpurun_out/r02n8b_p2pb.err; echo "p2pb rc=$?"
CUDA_VISIBLE_DEVICES=0,1,2,3 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r02n8b_n4.json 2> gpurun_out/r02n8b_n4.err; echo "n4 rc=$?"
for f in default p2pb n4; do cut -c1-260 gpurun_out/r02n8b_$f.json; done
