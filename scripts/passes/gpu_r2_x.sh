#!/bin/bash
# round 2, pass X (1 GPU): K1 vs K1L on small batches at the C3 shape (1M x 768 cosine, M = 32, k = 100, ef = 128), incremental quality graph
mkdir -p gpurun_out; rm -f gpurun_out/r02x_*
C5_DIM=768 C5_M=32 C5_K=100 C5_EFS=128 C5_METRIC=1 C5_NQ=1,64,148,296 timeout 900 python scripts/c5_team_sweep.py incremental > gpurun_out/r02x_c3_team_sweep.jsonl 2> gpurun_out/r02x_c3_team_sweep.err; echo "rc=$?"; tail -2 gpurun_out/r02x_c3_team_sweep.err
cut -c1-400 gpurun_out/r02x_c3_team_sweep.jsonl
