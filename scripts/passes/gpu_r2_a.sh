#!/bin/bash
# round 2, pass A: GPU tests on the new kernel (global-hash visited set, fused one-launch sharded step), variant sweep, default bench line
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02_a_tests.log
tail -4 gpurun_out/r02_a_tests.log
python scripts/sweep_variants.py > gpurun_out/r02_a_sweep.jsonl 2> gpurun_out/r02_a_sweep.err
tail -3 gpurun_out/r02_a_sweep.err
python bench.py --steps 20 --warmup 5 --no-track > gpurun_out/r02_a_bench.json 2> gpurun_out/r02_a_bench.err
tail -2 gpurun_out/r02_a_bench.err
