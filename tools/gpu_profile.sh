#!/bin/bash
# Round-end ncu captures on one B200 (under gpurun).  Numbers printed by a run under ncu are never bench values.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out; mkdir -p $O
# (1) launch list of the default bench command (short): every kernel launch with its duration
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-parity > $O/r2_launches_bench.log 2>&1
# (2) DRAM traffic + instruction counts of the pipeline's kernels on the REAL C2 config (one step)
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum,sm__inst_executed_pipe_lsu.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k "regex:skm_|chain_|slice_scan" --csv --log-file $O/r2_traffic_c2.csv python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-parity > $O/r2_traffic_bench.log 2>&1
# (3) full set with source for the three hot kernels (smaller transcriptome at C2's coverage so that the replays fit)
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:skm_emit|skm_merge|skm_build" -s 2 -c 5 -f -o $O/r2_skm python bench.py --pairs 4000000 --transcripts 3200 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-parity > $O/r2_ncu_full.log 2>&1
ls -la $O | grep r2_
