#!/bin/bash
# Round-2 A/B on one box: per-variant kernel throughput on C2 / C4 / C3 shapes (no per-variant pytest: winners are re-tested).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
for n in "$@"; do
  so=$PWD/needletail_b200/libntgpu_$n.so; [ -f "$so" ] || { echo "missing $n"; continue; }
  echo "== $n"
  NTGPU_SO=$so NT_MC_ONLY=${NT_MC_ONLY:-0,2,3} timeout 200 python tools/measure_configs.py 2>gpurun_out/ab2_$n.err | tee gpurun_out/ab2_$n.json | python -c '
import sys, json
for l in sys.stdin:
    d = json.loads(l); print("   %-60s %8.1f Gbases/s %8.1f GB/s %s" % (d["config"][:60], d["gbases_per_s"], d["gb_per_s"], json.dumps(d.get("stats", ""))))'
  tail -2 gpurun_out/ab2_$n.err
done
