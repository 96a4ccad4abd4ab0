#!/bin/bash
# K1 L2-prefetch A/B (1M x 128 both graphs, then C3-shaped 1M x 768)
mkdir -p gpurun_out
timeout 600 python scripts/sweep_prefetch.py > gpurun_out/prefetch_ab.jsonl 2> gpurun_out/prefetch_ab.err; echo "ab rc=$?"; cat gpurun_out/prefetch_ab.jsonl; tail -3 gpurun_out/prefetch_ab.err
