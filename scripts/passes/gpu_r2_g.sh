#!/bin/bash
# round 2, pass G (2 GPUs): N=2 bench lines with the gather-to-owner e2e and with the replicated e2e
mkdir -p gpurun_out; rm -f gpurun_out/r02g_*
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02g_n2_p2p.json 2> gpurun_out/r02g_n2_p2p.err; tail -2 gpurun_out/r02g_n2_p2p.err
$TR bench.py --gpus 2 --steps 20 --warmup 5 --e2e-input replicated --parity-queries 0 > gpurun_out/r02g_n2_p2p_e2e_replicated.json 2> gpurun_out/r02g_n2_p2p_rep.err; tail -2 gpurun_out/r02g_n2_p2p_rep.err
