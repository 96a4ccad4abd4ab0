#!/bin/bash
# round 2, pass I (2 GPUs): record-form exchange: all GPU tests, N=2 bench (records vs blocks), kernel-speed A/B against the round-1 library
mkdir -p gpurun_out; rm -f gpurun_out/r02i_*
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r02i_tests.log; tail -3 gpurun_out/r02i_tests.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02i_n2.json 2> gpurun_out/r02i_n2.err; tail -1 gpurun_out/r02i_n2.err
$TR bench.py --gpus 2 --steps 20 --warmup 5 --exchange p2pb --parity-queries 0 > gpurun_out/r02i_n2_p2pb.json 2> gpurun_out/r02i_n2_p2pb.err
for f in r02i_n2 r02i_n2_p2pb; do cut -c1-240 gpurun_out/$f.json; done
CUDA_VISIBLE_DEVICES=0 ZVDB_TREE=$PWD/scripts/dev/r1_tree python scripts/dev/ab_time.py r1 >> gpurun_out/r02i_ab.jsonl 2>> gpurun_out/r02i_ab.err
CUDA_VISIBLE_DEVICES=0 python scripts/dev/ab_time.py today >> gpurun_out/r02i_ab.jsonl 2>> gpurun_out/r02i_ab.err
grep -v '"variant": 12' gpurun_out/r02i_ab.jsonl | grep reference
