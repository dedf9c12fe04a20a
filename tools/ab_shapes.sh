#!/bin/bash
# A/B of CTA shapes of the fused kernel (walker threads / CTAs per SM / tile KiB / newline capacity), same box, back to back:
#   tools/ab_shapes.sh build                                   # here (no GPU)
#   gpurun --timeout 900 -- 'bash tools/ab_shapes.sh run'      # on the box
set -u
cd "$(dirname "$0")/.."
declare -A V=(
  [s288x2]=""
  [s192x3]="-DNTG_NTW=192 -DNTG_CTAS=3 -DNTG_TILE_KB=60 -DNTG_NLMAX=1024"
  [s128x4]="-DNTG_NTW=128 -DNTG_CTAS=4 -DNTG_TILE_KB=40 -DNTG_NLMAX=768"
  [s96x5]="-DNTG_NTW=96 -DNTG_CTAS=5 -DNTG_TILE_KB=30 -DNTG_NLMAX=512"
  [s160x3]="-DNTG_NTW=160 -DNTG_CTAS=3 -DNTG_TILE_KB=52 -DNTG_NLMAX=1024"
)
F="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -shared -Xcompiler -fPIC --expt-relaxed-constexpr -ldl -lz"
case "${1:-}" in
  build)
    for n in "${!V[@]}"; do /usr/local/cuda/bin/nvcc $F ${V[$n]} -o needletail_b200/libntgpu_$n.so needletail_b200/csrc/ntgpu.cu 2>gpurun_out/build_$n.err & done; wait
    ls -la needletail_b200/libntgpu_*.so ;;
  run)
    shift
    bash tools/ab_run2.sh "${@:-s288x2 s192x3 s128x4 s96x5 s160x3 s288x2}" ;;
  clean) rm -f needletail_b200/libntgpu_*.so ;;
  *) echo "usage: $0 build|run|clean"; exit 2 ;;
esac
