#!/bin/bash
# One GPU box: A/B of the phase-C bank-conflict change, smoke(), the full GPU test suite, ncu captures for profiles/.
cd "$(dirname "$0")/.."
for lib in "" /root/repo/vae_training_b200/libvae_b200_exp1.so; do
  for i in 1 2; do
    VAE_B200_LIB=$lib python bench.py --steps 2000 --warmup 5 --no-sub --no-ab --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('lib=$lib', round(d['value']/1e6,2))"
  done
done
VAE_B200_LIB=/root/repo/vae_training_b200/libvae_b200_exp1.so python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fused_small or shared_eps or fused_fixed" 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu_final.log 2>&1; echo pytest rc=$?; tail -4 gpurun_out/r02_pytest_gpu_final.log
ncu --set full --clock-control none --import-source on -k regex:fused_small -s 3 -c 1 -f -o gpurun_out/r02_fused_small_k20 python bench.py --steps 20 --warmup 5 --no-sub --no-ab --no-cpu-baseline > gpurun_out/r02_ncu_small_k20.log 2>&1; echo rc=$?
ncu --set full --clock-control none --import-source on -k regex:fused_mlp_tc -s 2 -c 1 -f -o gpurun_out/r02_fused_mlp_tc python bench.py --workload sphere_dd3_pd3_ld6 --steps 100 --warmup 5 --no-sub --no-ab --no-cpu-baseline > gpurun_out/r02_ncu_tc.log 2>&1; echo rc=$?
ncu --set full --clock-control none --import-source on -k regex:fused_mlp_tc -s 2 -c 1 -f -o gpurun_out/r02_fused_mlp_tc_k20 python bench.py --workload sphere_dd3_pd3_ld6 --steps 20 --warmup 5 --no-sub --no-ab --no-cpu-baseline > gpurun_out/r02_ncu_tc_k20.log 2>&1; echo rc=$?
ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/r02_launches_default.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_ncu_list.log 2>&1; echo rc=$?
wc -l gpurun_out/r02_launches_default.csv
