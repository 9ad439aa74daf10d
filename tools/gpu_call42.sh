#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out; mkdir -p $O
BENCH_TRACE_RANK0=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --no-parity > $O/c42.json 2> $O/c42.err
grep "sdtgpu" $O/c42.err | tail -n 70 | cut -c1-120
