#!/bin/bash
# round 2, pass L (1 GPU): evidence pass for the final tree: all GPU tests, smoke, default bench line + reference arm, ncu launch list of the
# default bench command, ncu --set full captures of K1 (reference graph ef=64; incremental graph ef=128 and 512), visited-mode sweep, C harness
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep gpurun_out/r02l_*
nvidia-smi --query-gpu=name,memory.total,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02l_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r02l_tests.log; tail -2 gpurun_out/r02l_tests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02l_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r02l_smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02l_bench_n1.json 2> gpurun_out/r02l_bench_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02l_bench_ref.json 2> gpurun_out/r02l_bench_ref.err; echo "ref rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 --sweep --no-track > gpurun_out/r02l_bench_n1_sweep.json 2> gpurun_out/r02l_bench_n1_sweep.err; echo "sweep rc=$?"
summ() {
  ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1.raw.csv 2>/dev/null
  ncu -i gpurun_out/$1.ncu-rep --page source --csv > gpurun_out/$1.source.csv 2>/dev/null
  rm -f gpurun_out/$1.ncu-rep
}
Q="--no-cpu --no-recall --no-track --parity-queries 0 --steps 3 --warmup 3"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02l_launches_bench_default.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu > /dev/null 2> gpurun_out/r02l_ncu_launch.err; echo "launch list rc=$?"
cap() { local name=$1 skip=$2; shift 2
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:search_layer0 -s $skip -c 1 -o gpurun_out/$name -f \
      python bench.py $Q "$@" > /dev/null 2> gpurun_out/$name.err; echo "$name rc=$?"; summ $name; }
cap r02l_k1_ref_ef64 6
cap r02l_k1_inc_ef128 12 --graph incremental --ef 128
cap r02l_k1_inc_ef512 12 --graph incremental --ef 512
python scripts/summarise_ncu.py gpurun_out/r02l_k1_ref_ef64 gpurun_out/r02l_k1_inc_ef128 gpurun_out/r02l_k1_inc_ef512 > gpurun_out/r02l_k1_ncu.md 2>/dev/null
rm -f gpurun_out/r02l_*.source.csv
timeout 300 python scripts/sweep_variants.py > gpurun_out/r02l_visited_sweep.jsonl 2> gpurun_out/r02l_visited_sweep.err; echo "visited sweep rc=$?"
gcc -O2 -std=c99 -Iinclude integration/harness.c -Lzvdb_b200/lib -lzvdb_b200 -Wl,-rpath,$PWD/zvdb_b200/lib -o gpurun_out/r02l_harness && \
  for k in 10 100; do gpurun_out/r02l_harness 100000 128 10000 $k; done > gpurun_out/r02l_c_harness.txt 2>&1; rm -f gpurun_out/r02l_harness; cat gpurun_out/r02l_c_harness.txt
cut -c1-300 gpurun_out/r02l_bench_n1.json
