#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out; mkdir -p $O
SDTGPU_TRACE=1 timeout 600 python bench.py --config C5 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-parity > $O/c36.json 2> $O/c36.err
tail -n 60 $O/c36.err | cut -c1-220
