#!/bin/bash
# GPU call 1 of the re-entry session: correctness of the overlapped ("dual") schedule of fused_mlp, A/B against the
# sequential schedule, cycle breakdown, A/B of the leader-z eps sharing of fused_small, then the full GPU suite.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
R=/root/repo/vae_training_b200
echo "== mlp correctness (dual default)"
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fused_cluster and not shared_eps" 2>&1 | tail -3
echo "== mlp A/B"
VAE_MLP_DUAL=0 timeout 300 python scratch/mlp_ab.py 2>&1 | grep -v Warning | sed 's/^/seq  /'
timeout 300 python scratch/mlp_ab.py 2>&1 | grep -v Warning | sed 's/^/dual /'
echo "== mlp timing"
VAE_B200_LIB=$R/libvae_b200_timing.so timeout 300 python scratch/mlp_timing.py 2>&1 | tail -40
echo "== small A/B"
for i in 1 2; do
  for lib in "" $R/libvae_b200_leadz.so; do
    VAE_B200_LIB=$lib timeout 300 python bench.py --steps 20 --warmup 5 --no-sub --no-ab --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('k20 lib=$lib', round(d['value']/1e6,2), 'e2e', round(d['e2e']['value']/1e6,2))"
    VAE_B200_LIB=$lib timeout 300 python bench.py --steps 2000 --warmup 5 --no-sub --no-ab --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('k2000 lib=$lib', round(d['value']/1e6,2))"
  done
done
echo "== small correctness with leadz"
VAE_B200_LIB=$R/libvae_b200_leadz.so timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_ref_golden.py -m gpu -q -x -k "fused_small or shared_eps or fused_fixed or golden" 2>&1 | tail -3
echo "== full suite (default lib)"
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02b_pytest_gpu_call1.log 2>&1; echo pytest rc=$?; tail -4 gpurun_out/r02b_pytest_gpu_call1.log
