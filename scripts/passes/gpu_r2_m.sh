#!/bin/bash
# round 2, pass M (1 GPU): re-entry verification of the final tree (all GPU tests, smoke, default bench line, reference arm) + one ncu --set full
# capture of K1 at nq = 1 (C5: where a single query's time goes)
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep gpurun_out/r02m_*
nvidia-smi --query-gpu=name,memory.total,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02m_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q --durations=8 2>&1 | tail -16 > gpurun_out/r02m_tests.log; tail -3 gpurun_out/r02m_tests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02m_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r02m_smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02m_bench_n1.json 2> gpurun_out/r02m_bench_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02m_bench_ref.json 2> gpurun_out/r02m_bench_ref.err; echo "ref rc=$?"
Q="--no-cpu --no-recall --no-track --parity-queries 0 --steps 3 --warmup 3 --nq 1"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:search_layer0 -s 6 -c 1 -o gpurun_out/r02m_k1_nq1 -f \
    python bench.py $Q > /dev/null 2> gpurun_out/r02m_k1_nq1.err; echo "ncu nq1 rc=$?"
ncu -i gpurun_out/r02m_k1_nq1.ncu-rep --page raw --csv > gpurun_out/r02m_k1_nq1.raw.csv 2>/dev/null
ncu -i gpurun_out/r02m_k1_nq1.ncu-rep --page source --csv > gpurun_out/r02m_k1_nq1.source.csv 2>/dev/null
rm -f gpurun_out/r02m_k1_nq1.ncu-rep
cut -c1-400 gpurun_out/r02m_bench_n1.json
