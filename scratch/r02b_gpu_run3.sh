#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== ffma2 probe"
timeout 120 ./scratch/ffma2_probe
echo "== ncu fused_mlp cs16"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_mlp_kernel -s 2 -c 1 -f -o gpurun_out/r02b_fused_mlp_cs16 python bench.py --workload sphere_dd3_pd3_ld6 --models 2 --steps 100 --warmup 5 --no-sub --no-ab --no-cpu-baseline > gpurun_out/r02b_ncu_mlp.log 2>&1; echo rc=$?
tail -3 gpurun_out/r02b_ncu_mlp.log
ls -la gpurun_out/*.ncu-rep
