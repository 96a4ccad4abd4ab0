#!/bin/bash
# evidence pass: tests, smoke, bench lines (both graphs, descent, reference arm), K4 lines, ncu launch lists and --set full
# captures of K1 and K4 summarised on the box (gpurun_out/ is capped at 64 MiB), C3 at full size
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep
nvidia-smi --query-gpu=name,memory.total,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py --steps 20 --warmup 3 --sweep > gpurun_out/bench_ref_g.json 2> gpurun_out/bench_ref_g.err; echo "bench ref rc=$?"
timeout 600 python bench.py --steps 10 --warmup 3 --graph quality --ef 128 --sweep > gpurun_out/bench_q_g.json 2> gpurun_out/bench_q_g.err; echo "bench q rc=$?"
timeout 600 python bench.py --steps 20 --warmup 3 --descent > gpurun_out/bench_ref_descent_g.json 2> gpurun_out/bench_ref_descent_g.err; echo "bench descent rc=$?"
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_refarm_g.json 2> gpurun_out/bench_refarm_g.err; echo "refarm rc=$?"
for a in "--dim 128 --k 10" "--dim 128 --k 10 --nq 65536 --steps 2" "--dim 768 --k 10 --steps 2" "--dim 768 --k 100 --steps 2" "--dim 128 --k 100 --steps 2" "--dim 128 --k 10 --filter"; do
  timeout 200 python scripts/bench_bruteforce.py $a >> gpurun_out/bench_k4_g.jsonl 2>> gpurun_out/bench_k4_g.err; done; echo "k4 lines rc=$?"
summ() {
  ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1.raw.csv 2>/dev/null
  ncu -i gpurun_out/$1.ncu-rep --page source --csv > gpurun_out/$1.source.csv 2>/dev/null
  rm -f gpurun_out/$1.ncu-rep
}
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bench_default.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/ncu_launch_bench.json 2> gpurun_out/ncu_launch_bench.err; echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:search_layer0 -s 6 -c 1 -o gpurun_out/prof4_k1_ref_ef64 -f \
    python bench.py --steps 3 --warmup 3 --no-cpu > /dev/null 2> gpurun_out/ncu_full_ref.err; echo "full ref rc=$?"; summ prof4_k1_ref_ef64
timeout 600 ncu --set full --clock-control none --import-source on -k regex:search_layer0 -s 6 -c 1 -o gpurun_out/prof4_k1_ref_ef512 -f \
    python bench.py --steps 3 --warmup 3 --ef 512 --no-cpu > /dev/null 2> gpurun_out/ncu_full_ref512.err; echo "full ref512 rc=$?"; summ prof4_k1_ref_ef512
timeout 600 ncu --set full --clock-control none --import-source on -k regex:search_layer0 -s 6 -c 1 -o gpurun_out/prof4_k1_q_ef128 -f \
    python bench.py --steps 3 --warmup 3 --ef 128 --graph quality --no-cpu > /dev/null 2> gpurun_out/ncu_full_q.err; echo "full q rc=$?"; summ prof4_k1_q_ef128
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bf_gemm_topk -s 1 -c 1 -o gpurun_out/prof4_k4_pair_128 -f \
    python scripts/bench_bruteforce.py --steps 1 > /dev/null 2> gpurun_out/ncu_full_k4.err; echo "full k4 rc=$?"; summ prof4_k4_pair_128
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/launches_k4_pair.csv \
    python scripts/bench_bruteforce.py --steps 2 > gpurun_out/ncu_launch_k4.json 2> gpurun_out/ncu_launch_k4.err; echo "k4 launch list rc=$?"
timeout 900 python scripts/configs_c3_c5.py c3 > gpurun_out/c3.jsonl 2> gpurun_out/c3.err; echo "c3 rc=$?"
du -sh gpurun_out
