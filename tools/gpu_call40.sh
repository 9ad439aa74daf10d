#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out; mkdir -p $O
timeout 420 ncu --section SourceCounters --section WarpStateStats --section SpeedOfLight --section SchedulerStats --clock-control none --import-source on -k regex:'skm_merge' -s 0 -c 1 -o $O/r2_merge_c2 -f python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-parity > $O/c40.log 2>&1
tail -3 $O/c40.log | cut -c1-200
ls -la $O/r2_merge_c2.ncu-rep
