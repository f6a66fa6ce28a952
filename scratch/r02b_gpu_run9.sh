#!/bin/bash
cd "$(dirname "$0")/.."
R=/root/repo/vae_training_b200
echo "== mlp correctness (Adam warps: 16 parameters per thread in flight, own function)"
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_ref_golden.py tests/test_gpu_frontend.py -m gpu -q -x -k "fused_cluster and not shared_eps or golden or schedule" 2>&1 | tail -3
VAE_B200_LIB=$R/libvae_b200_nodefer.so timeout 300 python scratch/mlp_ab.py 2>&1 | grep -v Warning | grep "models=[12] " | sed 's/^/nodefer /'
timeout 300 python scratch/mlp_ab.py 2>&1 | grep -v Warning | sed 's/^/defer16 /'
echo "== timing"
VAE_B200_LIB=$R/libvae_b200_timing.so timeout 300 python scratch/mlp_timing.py 2>&1 | tail -17
echo "== full suite (default lib)"
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02b_pytest_gpu_call9.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/r02b_pytest_gpu_call9.log
