#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
summ() {
  ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1.raw.csv 2>/dev/null
  ncu -i gpurun_out/$1.ncu-rep --page source --csv > gpurun_out/$1.source.csv 2>/dev/null
  ncu -i gpurun_out/$1.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/$1.cuda_source.csv 2>/dev/null
  rm -f gpurun_out/$1.ncu-rep
}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bf_gemm_topk -s 1 -c 1 -o gpurun_out/prof3_k4_pair_128 -f \
    python scripts/bench_bruteforce.py --steps 1 > /dev/null 2> gpurun_out/ncu_full_k4.err; echo "full k4 rc=$?"; summ prof3_k4_pair_128
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/launches_k4_pair.csv \
    python scripts/bench_bruteforce.py --steps 2 > gpurun_out/ncu_launch_k4.json 2> gpurun_out/ncu_launch_k4.err; echo "k4 launch list rc=$?"
