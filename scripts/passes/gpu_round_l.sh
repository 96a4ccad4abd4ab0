#!/bin/bash
# evidence pass for the session-4 kernel: all GPU tests, smoke, bench lines (reference + quality graph, reference arm),
# ncu launch list + --set full captures of K1 summarised on the box, the reference's own benchmark sweep, C3 and C5
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep gpurun_out/prof6_*
nvidia-smi --query-gpu=name,memory.total,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 600 python bench.py --steps 20 --warmup 3 --sweep > gpurun_out/bench_ref_l.json 2> gpurun_out/bench_ref_l.err; echo "bench ref rc=$?"
timeout 600 python bench.py --steps 10 --warmup 3 --graph quality --ef 128 --sweep > gpurun_out/bench_q_l.json 2> gpurun_out/bench_q_l.err; echo "bench q rc=$?"
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_refarm_l.json 2> gpurun_out/bench_refarm_l.err; echo "refarm rc=$?"
summ() {
  ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1.raw.csv 2>/dev/null
  ncu -i gpurun_out/$1.ncu-rep --page source --csv > gpurun_out/$1.source.csv 2>/dev/null
  rm -f gpurun_out/$1.ncu-rep
}
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bench_default.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/ncu_launch_bench.json 2> gpurun_out/ncu_launch_bench.err; echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:search_layer0 -s 6 -c 1 -o gpurun_out/prof6_k1_ref_ef64 -f \
    python bench.py --steps 3 --warmup 3 --no-cpu --no-recall > /dev/null 2> gpurun_out/ncu_full_ref.err; echo "full ref rc=$?"; summ prof6_k1_ref_ef64
timeout 600 ncu --set full --clock-control none --import-source on -k regex:search_layer0 -s 6 -c 1 -o gpurun_out/prof6_k1_ref_ef512 -f \
    python bench.py --steps 3 --warmup 3 --ef 512 --no-cpu --no-recall > /dev/null 2> gpurun_out/ncu_full_ref512.err; echo "full ref512 rc=$?"; summ prof6_k1_ref_ef512
timeout 600 ncu --set full --clock-control none --import-source on -k regex:search_layer0 -s 6 -c 1 -o gpurun_out/prof6_k1_q_ef128 -f \
    python bench.py --steps 3 --warmup 3 --ef 128 --graph quality --no-cpu --no-recall > /dev/null 2> gpurun_out/ncu_full_q.err; echo "full q rc=$?"; summ prof6_k1_q_ef128
python scripts/summarise_ncu.py gpurun_out/prof6_k1_ref_ef64 gpurun_out/prof6_k1_ref_ef512 gpurun_out/prof6_k1_q_ef128 > gpurun_out/k1_ncu_l.md 2>/dev/null
rm -f gpurun_out/prof6_*.source.csv
timeout 600 python -m zvdb_b200.benchmarks single > gpurun_out/reference_harness_single.txt 2> gpurun_out/reference_harness_single.err; echo "harness rc=$?"; tail -12 gpurun_out/reference_harness_single.txt
timeout 600 python -m zvdb_b200.benchmarks single --batched --dims 128 --ks 10 > gpurun_out/reference_harness_batched.txt 2>> gpurun_out/reference_harness_single.err; echo "harness batched rc=$?"; cat gpurun_out/reference_harness_batched.txt
timeout 600 python scripts/configs_c3_c5.py c5 > gpurun_out/c5.jsonl 2> gpurun_out/c5.err; echo "c5 rc=$?"
timeout 900 python scripts/configs_c3_c5.py c3 > gpurun_out/c3.jsonl 2> gpurun_out/c3.err; echo "c3 rc=$?"; cut -c1-300 gpurun_out/c3.jsonl
du -sh gpurun_out
