#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out; mkdir -p $O
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'skm_merge|skm_build2' -s 0 -c 2 -o $O/r2_mb -f python bench.py --pairs 4000000 --transcripts 3200 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-parity > $O/c38.log 2>&1
tail -3 $O/c38.log | cut -c1-200
ls -la $O/r2_mb.ncu-rep
