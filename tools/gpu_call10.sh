#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=memory.total,memory.used --format=csv
for C in C3; do
SDTGPU_TRACE=1 timeout 600 python bench.py --config $C --path sliced --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > $O/c10_bench_$C.json 2> $O/c10_bench_$C.err; tail -n 30 $O/c10_bench_$C.err | cut -c1-300
done
