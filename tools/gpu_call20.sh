#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests/test_gpu_dropin.py -x -q -m gpu > $O/c20_dropin.log 2>&1
tail -6 $O/c20_dropin.log
timeout 2400 python -m pytest tests/test_gpu_fullsize.py -x -q -m gpu > $O/c20_full.log 2>&1
tail -12 $O/c20_full.log
