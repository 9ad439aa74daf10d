#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu -k skm > $O/c26_multi.log 2>&1
tail -5 $O/c26_multi.log
bash tools/gpu_call17.sh 2
