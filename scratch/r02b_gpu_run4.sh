#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== mlp correctness (default: scalar FFMA at 16 CTAs, FFMA2 below)"
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fused_cluster and not shared_eps" 2>&1 | tail -3
echo "== A/B FMA form"
VAE_MLP_FFMA2=1 timeout 300 python scratch/mlp_ab.py 2>&1 | grep -v Warning | sed 's/^/ffma2  /'
VAE_MLP_FFMA2=0 timeout 300 python scratch/mlp_ab.py 2>&1 | grep -v Warning | sed 's/^/scalar /'
VAE_MLP_FFMA2=0 VAE_MLP_DUAL=0 timeout 300 python scratch/mlp_ab.py 2>&1 | grep -v Warning | sed 's/^/scalar-seq /'
