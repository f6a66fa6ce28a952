#!/bin/bash
# GPU call 2: overlapped schedule + bulk (TMA) operand copies + batched bias Adam / fast index math in fused_mlp;
# leader-z eps sharing now the default of fused_small.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
R=/root/repo/vae_training_b200
echo "== mlp correctness"
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fused_cluster and not shared_eps" 2>&1 | tail -3
echo "== mlp A/B"
VAE_MLP_DUAL=0 timeout 300 python scratch/mlp_ab.py 2>&1 | grep -v Warning | sed 's/^/seq  /'
timeout 300 python scratch/mlp_ab.py 2>&1 | grep -v Warning | sed 's/^/dual /'
echo "== mlp timing"
VAE_B200_LIB=$R/libvae_b200_timing.so timeout 300 python scratch/mlp_timing.py 2>&1 | tail -34
echo "== full suite"
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02b_pytest_gpu_call2.log 2>&1; echo pytest rc=$?; tail -4 gpurun_out/r02b_pytest_gpu_call2.log
echo "== small bench"
timeout 300 python bench.py --steps 20 --warmup 5 --no-sub --no-ab --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('k20', round(d['value']/1e6,2), 'e2e', round(d['e2e']['value']/1e6,2))"
