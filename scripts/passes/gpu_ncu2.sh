#!/bin/bash
# ncu evidence, second pass: launch list of the default bench command, --set full captures of K1 (both graphs) and K4.
# The .ncu-rep files are summarised on the box (raw + source pages as CSV) because gpurun_out/ is capped at 64 MiB.
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
summ() {  # $1 = report basename
  ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1.raw.csv 2>/dev/null
  ncu -i gpurun_out/$1.ncu-rep --page source --csv > gpurun_out/$1.source.csv 2>/dev/null
  ncu -i gpurun_out/$1.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/$1.cuda_source.csv 2>/dev/null
  [ "$2" = keep ] || rm -f gpurun_out/$1.ncu-rep
}
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bench_default.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/ncu_launch_bench.json 2> gpurun_out/ncu_launch_bench.err; echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:search_layer0 -s 6 -c 1 -o gpurun_out/prof2_k1_ref_ef64 -f \
    python bench.py --steps 3 --warmup 3 --no-cpu > /dev/null 2> gpurun_out/ncu_full_ref.err; echo "full ref rc=$?"; summ prof2_k1_ref_ef64
timeout 600 ncu --set full --clock-control none --import-source on -k regex:search_layer0 -s 6 -c 1 -o gpurun_out/prof2_k1_ref_ef512 -f \
    python bench.py --steps 3 --warmup 3 --ef 512 --no-cpu > /dev/null 2> gpurun_out/ncu_full_ref512.err; echo "full ref512 rc=$?"; summ prof2_k1_ref_ef512
timeout 600 ncu --set full --clock-control none --import-source on -k regex:search_layer0 -s 6 -c 1 -o gpurun_out/prof2_k1_q_ef128 -f \
    python bench.py --steps 3 --warmup 3 --ef 128 --graph quality --no-cpu > /dev/null 2> gpurun_out/ncu_full_q.err; echo "full q rc=$?"; summ prof2_k1_q_ef128
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bf_gemm_topk -s 1 -c 1 -o gpurun_out/prof2_k4_128 -f \
    python scripts/bench_bruteforce.py --steps 1 > /dev/null 2> gpurun_out/ncu_full_k4.err; echo "full k4 rc=$?"; summ prof2_k4_128 keep
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/launches_k4.csv \
    python scripts/bench_bruteforce.py --steps 2 > gpurun_out/ncu_launch_k4.json 2> gpurun_out/ncu_launch_k4.err; echo "k4 launch list rc=$?"
du -sh gpurun_out; ls -la gpurun_out
