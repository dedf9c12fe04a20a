#!/bin/bash
# One gpurun call: per-config kernel throughput, one ncu --set full capture of the fused kernel and the launch list.
# Usage: gpurun --timeout 1500 -- 'bash tools/gpu_round_check.sh TAG'
TAG=${1:-x}
timeout 500 python tools/measure_configs.py > gpurun_out/configs_$TAG.json 2> gpurun_out/configs_$TAG.err
cut -c1-150 gpurun_out/configs_$TAG.json; tail -2 gpurun_out/configs_$TAG.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_fused -s 2 -c 1 -o gpurun_out/prof_fused_$TAG python bench.py --steps 1 --warmup 3 --reads 10000000 --no-cpu --no-e2e --no-verify > gpurun_out/ncu_full_$TAG.log 2>&1; tail -1 gpurun_out/ncu_full_$TAG.log | cut -c1-120
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --reads 20000000 --no-cpu --no-e2e --no-verify > gpurun_out/ncu_launches_$TAG.log 2>&1; tail -1 gpurun_out/ncu_launches_$TAG.log | cut -c1-120
ls -la gpurun_out/*.ncu-rep
