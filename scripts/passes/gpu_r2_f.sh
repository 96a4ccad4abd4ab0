#!/bin/bash
# round 2, pass F (2 GPUs): all GPU tests (host step on two ranks), N=2 bench with the gather-to-owner e2e, N=1 default line,
# same-day A/B of the round-1 library against today's (global-visited modes)
mkdir -p gpurun_out; rm -f gpurun_out/r02f_*
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r02f_tests.log; tail -3 gpurun_out/r02f_tests.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02f_n2_p2p.json 2> gpurun_out/r02f_n2_p2p.err; tail -2 gpurun_out/r02f_n2_p2p.err
$TR bench.py --gpus 2 --steps 20 --warmup 5 --e2e-input replicated --parity-queries 0 > gpurun_out/r02f_n2_p2p_e2e_replicated.json 2> gpurun_out/r02f_n2_p2p_rep.err
python bench.py --steps 20 --warmup 5 > gpurun_out/r02f_n1.json 2> gpurun_out/r02f_n1.err; tail -2 gpurun_out/r02f_n1.err
CUDA_VISIBLE_DEVICES=0 ZVDB_TREE=$PWD/scripts/dev/r1_tree python scripts/dev/ab_time.py r1 >> gpurun_out/r02f_ab.jsonl 2>> gpurun_out/r02f_ab.err
CUDA_VISIBLE_DEVICES=0 python scripts/dev/ab_time.py today >> gpurun_out/r02f_ab.jsonl 2>> gpurun_out/r02f_ab.err
cat gpurun_out/r02f_ab.jsonl
