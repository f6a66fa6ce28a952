#!/bin/bash
# Round 2, last session: full GPU suite + smoke + driver-contract bench line with the flax-keyed init.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== full GPU suite"
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r02c_pytest_gpu.log 2>&1; echo pytest rc=$?; tail -4 gpurun_out/r02c_pytest_gpu.log
echo "== smoke"
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench --steps 20 --warmup 5"
timeout 500 python bench.py --steps 20 --warmup 5 > gpurun_out/r02c_bench_default_k20_n1.json 2> gpurun_out/r02c_bench_default_k20_n1.err; echo rc=$?
python -c "
import json
d=json.loads(open('gpurun_out/r02c_bench_default_k20_n1.json').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], 'clocks', d['clocks'])
for k,v in d['sub_results'].items(): print(' ', k, v.get('value'), v.get('unit'), v.get('ms_per_step'))
"
