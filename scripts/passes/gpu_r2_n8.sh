#!/bin/bash
# round 2, 8 GPUs: default bench line at N=8 (fused one-launch step, gather-to-owner e2e) with NVLink payload counters around it,
# the same with round 1's three launches, and C4 at full size (100M x 128, id-sharded over 8 GPUs) on both graphs
mkdir -p gpurun_out; rm -f gpurun_out/r02n8_*
nvidia-smi topo -m > gpurun_out/r02n8_topo.txt 2>&1; nvidia-smi nvlink -s -i 0 | head -24 >> gpurun_out/r02n8_topo.txt 2>&1
free -g | head -2 > gpurun_out/r02n8_host.txt; nproc >> gpurun_out/r02n8_host.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513"
timeout 600 python scripts/nvlink_counters.py gpurun_out/r02n8_nvlink_default.json -- $TR bench.py --gpus 8 --steps 20 --warmup 5 \
    > gpurun_out/r02n8_default.json 2> gpurun_out/r02n8_default.err; echo "default rc=$?"; tail -2 gpurun_out/r02n8_default.err
timeout 600 $TR bench.py --gpus 8 --steps 20 --warmup 5 --exchange p2p3 --parity-queries 0 > gpurun_out/r02n8_p2p3.json 2> gpurun_out/r02n8_p2p3.err; echo "p2p3 rc=$?"
timeout 900 $TR bench.py --gpus 8 --steps 10 --warmup 3 --rows 100000000 --shard-gen --graph incremental --ef 512 --sweep --no-cpu \
    > gpurun_out/r02n8_c4_incremental.json 2> gpurun_out/r02n8_c4_incremental.err; echo "c4 inc rc=$?"; grep -h "built" gpurun_out/r02n8_c4_incremental.err | head -2
timeout 900 $TR bench.py --gpus 8 --steps 10 --warmup 3 --rows 100000000 --shard-gen --graph reference --ef 512 --sweep --no-cpu \
    > gpurun_out/r02n8_c4_reference.json 2> gpurun_out/r02n8_c4_reference.err; echo "c4 ref rc=$?"; grep -h "built" gpurun_out/r02n8_c4_reference.err | head -2
cut -c1-400 gpurun_out/r02n8_default.json
