#!/bin/bash
# round 2, pass C (1 GPU): GPU tests (K4-fed builder, hierarchy + descent), ncu launch list of the default bench command, ncu --set full
# captures of K1 (reference graph ef=64; ef=512 on both graphs with the bitmap and with the global-hash visited set; incremental graph ef=128),
# recall-vs-evaluations curves (M = 16 / 32, with / without descent)
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep gpurun_out/r02c_*
nvidia-smi --query-gpu=name,memory.total,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02c_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r02c_tests.log; tail -3 gpurun_out/r02c_tests.log
summ() {
  ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1.raw.csv 2>/dev/null
  ncu -i gpurun_out/$1.ncu-rep --page source --csv > gpurun_out/$1.source.csv 2>/dev/null
  rm -f gpurun_out/$1.ncu-rep
}
Q="--no-cpu --no-recall --no-track --parity-queries 0 --steps 3 --warmup 3"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02c_launches_bench_default.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r02c_ncu_launch_bench.json 2> gpurun_out/r02c_ncu_launch_bench.err; echo "launch list rc=$?"
cap() {  # name skip bench-args...
  local name=$1 skip=$2; shift 2
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:search_layer0 -s $skip -c 1 -o gpurun_out/$name -f \
      python bench.py $Q "$@" > /dev/null 2> gpurun_out/$name.err; echo "$name rc=$?"; summ $name
}
cap r02c_k1_ref_ef64 6
cap r02c_k1_ref_ef512_bitmap 6 --ef 512 --variant 8
cap r02c_k1_ref_ef512_ghash 6 --ef 512 --variant 12
cap r02c_k1_inc_ef128_bitmap 12 --graph incremental --ef 128 --variant 8
cap r02c_k1_inc_ef512_bitmap 12 --graph incremental --ef 512 --variant 8
cap r02c_k1_inc_ef512_ghash 12 --graph incremental --ef 512 --variant 12
python scripts/summarise_ncu.py gpurun_out/r02c_k1_ref_ef64 gpurun_out/r02c_k1_ref_ef512_bitmap gpurun_out/r02c_k1_ref_ef512_ghash \
    gpurun_out/r02c_k1_inc_ef128_bitmap gpurun_out/r02c_k1_inc_ef512_bitmap gpurun_out/r02c_k1_inc_ef512_ghash > gpurun_out/r02c_k1_ncu.md 2>/dev/null
rm -f gpurun_out/r02c_*.source.csv
timeout 900 python scripts/recall_curve.py > gpurun_out/r02c_recall_curve.jsonl 2> gpurun_out/r02c_recall_curve.err; echo "recall curve rc=$?"; tail -2 gpurun_out/r02c_recall_curve.err
du -sh gpurun_out
