#!/bin/bash
# per-source-line instruction/stall profile of K1: reference graph ef=64 (shared hash) and quality graph ef=128 (bitmap)
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep
summ() {
  ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1.raw.csv 2>/dev/null
  ncu -i gpurun_out/$1.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/$1.cuda_source.csv 2>/dev/null
  python scripts/ncu_lines.py 70 < gpurun_out/$1.cuda_source.csv > gpurun_out/$1.lines.txt
  rm -f gpurun_out/$1.ncu-rep
}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:search_layer0 -s 6 -c 1 -o gpurun_out/prof5_k1_ref_ef64 -f \
    python bench.py --steps 3 --warmup 3 --no-cpu --no-recall > /dev/null 2> gpurun_out/ncu_full_ref.err; echo "full ref rc=$?"; summ prof5_k1_ref_ef64
timeout 600 ncu --set full --clock-control none --import-source on -k regex:search_layer0 -s 6 -c 1 -o gpurun_out/prof5_k1_q_ef128 -f \
    python bench.py --steps 3 --warmup 3 --ef 128 --graph quality --no-cpu --no-recall > /dev/null 2> gpurun_out/ncu_full_q.err; echo "full q rc=$?"; summ prof5_k1_q_ef128
du -sh gpurun_out; head -30 gpurun_out/prof5_k1_ref_ef64.lines.txt
