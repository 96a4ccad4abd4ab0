#!/bin/bash
# round 2: C4 (100M x 128, id-sharded) on the incremental quality graph at N = $1 GPUs (2 or 4): ef sweep, recall from K4 per shard + merge
N=${1:-4}
mkdir -p gpurun_out; rm -f gpurun_out/r02c4_n${N}_*
free -g | head -2 > gpurun_out/r02c4_n${N}_host.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29515"
timeout 1500 $TR bench.py --gpus $N --steps 10 --warmup 3 --rows 100000000 --shard-gen --graph incremental --ef 512 --sweep --no-cpu \
    > gpurun_out/r02c4_n${N}_incremental.json 2> gpurun_out/r02c4_n${N}_incremental.err; echo "c4 inc N=$N rc=$?"
grep -h "built" gpurun_out/r02c4_n${N}_incremental.err | head -2; cut -c1-300 gpurun_out/r02c4_n${N}_incremental.json
