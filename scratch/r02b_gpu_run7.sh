#!/bin/bash
cd "$(dirname "$0")/.."
R=/root/repo/vae_training_b200
echo "== mlp correctness (Adam of layer r-1 on two warps during the tiles of layer r)"
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_ref_golden.py tests/test_gpu_frontend.py -m gpu -q -x -k "fused_cluster and not shared_eps or golden or schedule" 2>&1 | tail -3
VAE_B200_LIB=$R/libvae_b200_nodefer.so timeout 300 python scratch/mlp_ab.py 2>&1 | grep -v Warning | sed 's/^/nodefer /'
timeout 300 python scratch/mlp_ab.py 2>&1 | grep -v Warning | sed 's/^/defer   /'
echo "== timing"
VAE_B200_LIB=$R/libvae_b200_timing.so timeout 300 python scratch/mlp_timing.py 2>&1 | tail -34
