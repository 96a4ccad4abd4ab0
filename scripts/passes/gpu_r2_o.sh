#!/bin/bash
# round 2, pass O (1 GPU): ncu --set full of K1L (one CTA per query) at nq = 8, sampling every 32 cycles; K1L parity tests
mkdir -p gpurun_out; rm -f gpurun_out/r02o_*

Q="--no-cpu --no-recall --no-track --parity-queries 0 --steps 3 --warmup 3 --nq 8 --variant 32768"
timeout 400 ncu --set full --clock-control none --import-source on --warp-sampling-interval 0 -k regex:search_team -s 6 -c 1 -o gpurun_out/r02o_k1l_nq8 -f \
    python bench.py $Q ${1:-} > gpurun_out/r02o_k1l_nq8.out 2> gpurun_out/r02o_k1l_nq8.err; echo "ncu rc=$?"; tail -n 3 gpurun_out/r02o_k1l_nq8.out
ncu -i gpurun_out/r02o_k1l_nq8.ncu-rep --page raw --csv > gpurun_out/r02o_k1l_nq8.raw.csv 2>/dev/null
ncu -i gpurun_out/r02o_k1l_nq8.ncu-rep --page source --csv > gpurun_out/r02o_k1l_nq8.source.csv 2>/dev/null
rm -f gpurun_out/r02o_k1l_nq8.ncu-rep
