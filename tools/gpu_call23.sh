#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out; mkdir -p $O
export SDTGPU_TRACE=1
timeout 900 python bench.py --config C4 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-parity > $O/c23_C4.json 2> $O/c23_C4.err; grep -v "^\[sdtgpu\]   *\(retry\|sub\|split\)" $O/c23_C4.err | tail -n 40 | cut -c1-200
timeout 900 python bench.py --config C3 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-parity > $O/c23_C3.json 2> $O/c23_C3.err; grep "sdtgpu\]" $O/c23_C3.err | grep -v "sub-slices\|  split\|^\[sdtgpu\]  list\|^\[sdtgpu\]  merge  *[0-9]\.\|build all items  *[0-9]\." | tail -n 30 | cut -c1-120
