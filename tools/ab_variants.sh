#!/bin/bash
# A/B of compile-time experiment builds of libntgpu on one GPU box (same box, back to back):
#   tools/ab_variants.sh build            # here (no GPU): builds needletail_b200/libntgpu_<name>.so for every variant
#   gpurun --timeout 900 -- 'bash tools/ab_variants.sh run'    # on the box: GPU parity tests + per-config throughput per variant
# (clean2 only matters for k = 51: run it with NT_MC_ONLY=4.)
# Variants are sets of -D flags of needletail_b200/csrc/fused.cuh; "default" is the shipped build.
set -u
cd "$(dirname "$0")/.."
declare -A V=(
  [default]=""
  [fp64min]="-DNTG_FP64_MIN=1"
  [ticket]="-DNTG_TICKET=1"
  [fp64min_ticket]="-DNTG_FP64_MIN=1 -DNTG_TICKET=1"
  [nodefer]="-DNTG_LB_DEFER=0"
  [stats]="-DNTG_STATS=1"
  [clean2]="-DNTG_CLEAN2=1"
  [wrap]="-DNTG_WRAP=1"
  [dc]="-DNTG_DC=1"
  [dc_ticket]="-DNTG_DC=1 -DNTG_TICKET=1"
  [dc_fp64min]="-DNTG_DC=1 -DNTG_FP64_MIN=1"
)
F="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -shared -Xcompiler -fPIC --expt-relaxed-constexpr -ldl -lz"
case "${1:-}" in
  build)
    for n in "${!V[@]}"; do /usr/local/cuda/bin/nvcc $F ${V[$n]} -o needletail_b200/libntgpu_$n.so needletail_b200/csrc/ntgpu.cu 2>/dev/null & done; wait
    ls -la needletail_b200/libntgpu_*.so ;;
  run)
    mkdir -p gpurun_out
    for n in default fp64min ticket fp64min_ticket nodefer stats clean2 wrap dc dc_ticket dc_fp64min; do
      so=$PWD/needletail_b200/libntgpu_$n.so; [ -f "$so" ] || continue
      echo "== $n"
      NTGPU_SO=$so timeout 300 python -m pytest tests -m gpu -q 2>&1 | tail -1
      NTGPU_SO=$so NT_MC_ONLY=${NT_MC_ONLY:-0,1,2} timeout 150 python tools/measure_configs.py 2>/dev/null | tee gpurun_out/ab_$n.json | cut -c1-125
    done ;;
  clean) rm -f needletail_b200/libntgpu_*.so ;;
  *) echo "usage: $0 build|run|clean"; exit 2 ;;
esac
