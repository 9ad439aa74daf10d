#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out; mkdir -p $O
PY=$(which python)
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 $PY -m pytest tests/test_gpu_sliced.py -x -q -m gpu -k "minimizer_lengths and 31-1-5" > $O/c8_memcheck.log 2>&1
grep -E "Invalid|at |by thread|Address|ERROR SUMMARY" $O/c8_memcheck.log | head -30
