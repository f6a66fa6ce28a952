#!/bin/bash
# Final GPU run of the round: full GPU suite, smoke(), driver-contract bench line, one rank's share of the 8-GPU sweep,
# ncu captures (--set full of the two changed kernels, launch list of the default bench command).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== full GPU suite"
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02b_pytest_gpu.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/r02b_pytest_gpu.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== bench --steps 20 --warmup 5"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02b_bench_default_k20_n1.json 2> gpurun_out/r02b_bench_default_k20_n1.err; echo rc=$?
python -c "
import json
d=json.loads(open('gpurun_out/r02b_bench_default_k20_n1.json').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], 'clocks', d['clocks'])
for k,v in d['sub_results'].items(): print(' ', k, v.get('value'), v.get('unit'), v.get('ms_per_step'))
"
echo "== bench --steps 2000"
timeout 600 python bench.py --steps 2000 --warmup 5 --no-sub --no-ab --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('k2000', round(d['value']/1e6,2))"
echo "== reference arm"
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02b_bench_reference_arm_k20.json 2>/dev/null; echo rc=$?; cut -c1-300 gpurun_out/r02b_bench_reference_arm_k20.json
echo "== sweep share of rank 0 of 8"
timeout 600 python scripts/full_sweep.py --emulate 0/8 --out gpurun_out/r02b_full_sweep_rank0_of_8.json 2>&1 | tail -3
echo "== ncu fused_small k20"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_small -s 3 -c 1 -f -o gpurun_out/r02b_fused_small_k20 python bench.py --steps 20 --warmup 5 --no-sub --no-ab --no-cpu-baseline > gpurun_out/r02b_ncu_small_k20.log 2>&1; echo rc=$?
echo "== ncu fused_mlp cs16"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_mlp_kernel -s 2 -c 1 -f -o gpurun_out/r02b_fused_mlp_cs16 python bench.py --workload sphere_dd3_pd3_ld6 --models 2 --steps 100 --warmup 5 --no-sub --no-ab --no-cpu-baseline > gpurun_out/r02b_ncu_mlp.log 2>&1; echo rc=$?
echo "== launch list of the default bench command"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/r02b_launches_default.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02b_ncu_list.log 2>&1; echo rc=$?
wc -l gpurun_out/r02b_launches_default.csv
