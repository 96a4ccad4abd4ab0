#!/bin/bash
# C4 (scaled): id-sharded index of $2 rows x 128 over $1 GPUs, reference insert per shard, ef=512, plus the
# 1M-row index at ef=512 for the strong-scaling table. usage: gpu_c4.sh N TOTAL_ROWS
N=${1:-2}; T=${2:-32000000}
mkdir -p gpurun_out
free -g | head -2 > gpurun_out/host_mem_n$N.txt; nproc >> gpurun_out/host_mem_n$N.txt
run() { # tag, extra args
  if [ "$N" = 1 ]; then L="python"; else L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513"; fi
  timeout 1500 $L bench.py --gpus $N --steps 10 --warmup 3 --no-cpu $2 > gpurun_out/bench_$1_n$N.json 2> gpurun_out/bench_$1_n$N.err; echo "bench $1 N=$N rc=$?"; cut -c1-1100 gpurun_out/bench_$1_n$N.json; grep -h "built" gpurun_out/bench_$1_n$N.err | head -2
}
if [ "$N" = 2 ]; then timeout 300 python -m pytest tests/test_gpu_sharded.py -x -q > gpurun_out/pytest_sharded.log 2>&1; echo "pytest sharded rc=$?"; tail -3 gpurun_out/pytest_sharded.log; fi
[ -n "$SKIP_1M" ] || run 1m_ef512 "--ef 512"
run c4_ef512 "--rows $T --ef 512 --shard-gen --sweep"
