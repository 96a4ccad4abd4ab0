#!/bin/bash
# round 2, pass E: same-day A/B of the round-1 library (scratch tree) against today's builds, K1 global-visited modes
mkdir -p gpurun_out; rm -f gpurun_out/r02e_ab.jsonl
ZVDB_TREE=$PWD/scripts/dev/r1_tree python scripts/dev/ab_time.py r1 >> gpurun_out/r02e_ab.jsonl 2>> gpurun_out/r02e_ab.err
for t in b200 ab_t1 ab_t2; do
  ZVDB_B200_LIB=$PWD/zvdb_b200/lib/libzvdb_$t.so python scripts/dev/ab_time.py $t >> gpurun_out/r02e_ab.jsonl 2>> gpurun_out/r02e_ab.err
done
grep -v '"variant": 12' gpurun_out/r02e_ab.jsonl; tail -3 gpurun_out/r02e_ab.err
