#!/bin/bash
cd "$(dirname "$0")/.."
echo "== mlp correctness (single bulk copy per array)"
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fused_cluster and not shared_eps" 2>&1 | tail -3
timeout 300 python scratch/mlp_ab.py 2>&1 | grep -v Warning | sed 's/^/bulk1 /'
