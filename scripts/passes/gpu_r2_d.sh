#!/bin/bash
# round 2, pass D: host-step tests (1 GPU part), A/B of K1 builds (fused tail compiled out / popped-key pointer) to locate a 10 % regression on the reference graph at ef=512
mkdir -p gpurun_out
python -m pytest tests/test_gpu_sharded.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r02d_tests.log; tail -3 gpurun_out/r02d_tests.log
for t in b200 ab_notail ab_notail_resptr; do
  ZVDB_B200_LIB=$PWD/zvdb_b200/lib/libzvdb_$t.so python scripts/dev/ab_time.py $t >> gpurun_out/r02d_ab.jsonl 2>> gpurun_out/r02d_ab.err
done
cat gpurun_out/r02d_ab.jsonl
