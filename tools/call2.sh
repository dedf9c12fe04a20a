timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
NT_MC_ONLY=0 bash tools/ab_run2.sh wv0 wv4 wv8 wv12 stats wv0
NTGPU_SO=$PWD/needletail_b200/libntgpu_nolb.so timeout 120 python tools/time_only.py 2>&1 | tail -2
timeout 120 python tools/time_only.py 2>&1 | tail -1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_fused -s 2 -c 1 -o gpurun_out/prof_r2b python bench.py --steps 1 --warmup 3 --reads 10000000 --no-cpu --no-e2e > gpurun_out/ncu_r2b.log 2>&1; tail -2 gpurun_out/ncu_r2b.log | cut -c1-200; ls -la gpurun_out/*.ncu-rep
