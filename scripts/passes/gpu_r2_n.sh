#!/bin/bash
# round 2, pass N (1 GPU): K1L (one CTA per query): parity tests, then the C5 A/B sweep against K1
mkdir -p gpurun_out; rm -f gpurun_out/r02n_*
timeout 600 python -m pytest tests/test_gpu_team_kernel.py -m gpu -q 2>&1 | tail -25 > gpurun_out/r02n_tests_team.log; tail -5 gpurun_out/r02n_tests_team.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/r02n_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02n_smoke.log
timeout 900 python scripts/c5_team_sweep.py ${1:-both} > gpurun_out/r02n_c5_team_sweep.jsonl 2> gpurun_out/r02n_c5_team_sweep.err; echo "sweep rc=$?"
cat gpurun_out/r02n_c5_team_sweep.jsonl | cut -c1-330; tail -3 gpurun_out/r02n_c5_team_sweep.err
