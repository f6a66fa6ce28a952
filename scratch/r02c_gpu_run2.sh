#!/bin/bash
# A/B of the fused_small launch set-up: every set-up load issued before the zero fill (default) vs the sequential set-up
# (libvae_b200_nopf.so = -DVAE_SMALL_PREFETCH=0); then the full GPU suite on the default library.
cd "$(dirname "$0")/.."
R=/root/repo/vae_training_b200
val() { python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']/1e6,2), 'M steps/s  launch_ms', round(d['timing']['launch_ms_median'],5), ' e2e', round(d['e2e']['value']/1e6,2))"; }
for rep in 1 2; do
  echo -n "prefetch k20: "; timeout 200 python bench.py --steps 20 --warmup 5 --no-sub --no-ab --no-cpu-baseline 2>/dev/null | val
  echo -n "nopf     k20: "; VAE_B200_LIB=$R/libvae_b200_nopf.so timeout 200 python bench.py --steps 20 --warmup 5 --no-sub --no-ab --no-cpu-baseline 2>/dev/null | val
done
echo -n "prefetch k2000: "; timeout 200 python bench.py --steps 2000 --warmup 5 --no-sub --no-ab --no-cpu-baseline 2>/dev/null | val
echo -n "nopf     k2000: "; VAE_B200_LIB=$R/libvae_b200_nopf.so timeout 200 python bench.py --steps 2000 --warmup 5 --no-sub --no-ab --no-cpu-baseline 2>/dev/null | val
echo -n "prefetch single chain k20: "; timeout 200 python bench.py --models 1 --steps 20 --warmup 5 --no-sub --no-ab --no-cpu-baseline 2>/dev/null | val
echo -n "nopf     single chain k20: "; VAE_B200_LIB=$R/libvae_b200_nopf.so timeout 200 python bench.py --models 1 --steps 20 --warmup 5 --no-sub --no-ab --no-cpu-baseline 2>/dev/null | val
echo "== full suite (default lib)"
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r02c_pytest_gpu_prefetch.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/r02c_pytest_gpu_prefetch.log
