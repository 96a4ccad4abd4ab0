#!/usr/bin/env python
"""NVLink payload bytes moved while a command runs: reads `nvidia-smi nvlink -gt d` (per-link data Tx/Rx KiB counters) on every
GPU before and after, prints one JSON line with the per-GPU and total deltas, and passes the command's output through.
usage: python scripts/nvlink_counters.py <out.json> -- <command ...>"""
import json, re, subprocess, sys

def read():
    out = subprocess.run(["nvidia-smi", "nvlink", "-gt", "d"], capture_output=True, text=True).stdout
    gpus, cur = {}, None
    for ln in out.splitlines():
        m = re.match(r"GPU (\d+):", ln)
        if m:
            cur = int(m.group(1)); gpus[cur] = {"tx_kib": 0, "rx_kib": 0}; continue
        m = re.search(r"Data Tx:\s*(\d+)\s*KiB", ln)
        if m and cur is not None: gpus[cur]["tx_kib"] += int(m.group(1))
        m = re.search(r"Data Rx:\s*(\d+)\s*KiB", ln)
        if m and cur is not None: gpus[cur]["rx_kib"] += int(m.group(1))
    return gpus, out

out_path, cmd = sys.argv[1], sys.argv[3:]
before, raw0 = read()
rc = subprocess.run(cmd).returncode
after, raw1 = read()
delta = {str(g): {k: after[g][k] - before[g][k] for k in after[g]} for g in after if g in before}
json.dump({"command": " ".join(cmd), "per_gpu_kib": delta, "total_tx_gib": sum(d["tx_kib"] for d in delta.values()) / 2**20,
           "total_rx_gib": sum(d["rx_kib"] for d in delta.values()) / 2**20,
           "note": "payload counters of nvidia-smi nvlink -gt d, all links of each GPU summed; covers the whole command (index build, warm-up, parity and e2e legs included)",
           "raw_sample": raw1[:600]}, open(out_path, "w"), indent=1)
sys.exit(rc)
