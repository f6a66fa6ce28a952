#!/bin/bash
cd "$(dirname "$0")/.."
echo "== mlp correctness (slices dealt from the last rank down)"
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_ref_golden.py tests/test_gpu_frontend.py -m gpu -q -x -k "fused_cluster and not shared_eps or golden or schedule" 2>&1 | tail -3
timeout 300 python scratch/mlp_ab.py 2>&1 | grep -v Warning | sed 's/^/revslice /'
