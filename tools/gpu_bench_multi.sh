#!/bin/bash
# Multi-GPU bench lines on one box (under `gpurun --gpus N -- 'bash tools/gpu_bench_multi.sh N'`): the driver's
# command at N GPUs (weak scaling) and, at 8, the strong-scaling line.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out; mkdir -p $O
N=${1:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 5 --warmup 3 > $O/r2_bench_c2_${N}gpu.json 2> $O/bench_multi.err
tail -n 2 $O/bench_multi.err | cut -c1-200
if [ "$N" = 8 ]; then
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 5 --warmup 3 --scaling strong --no-e2e --no-parity > $O/r2_bench_c2_${N}gpu_strong.json 2> $O/bench_multi_strong.err
tail -n 2 $O/bench_multi_strong.err | cut -c1-200
fi
python - $N <<'PY'
import json,sys
for f in ['gpurun_out/r2_bench_c2_%sgpu.json' % sys.argv[1], 'gpurun_out/r2_bench_c2_%sgpu_strong.json' % sys.argv[1]]:
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        ph=j['roofline']['sliced']['phases']
        print(f, j['scaling'], round(j['value']/1e9,2), round(j['ms_per_step'],2), {k:round(v['ms_per_step'],2) for k,v in ph.items()}, j['collective_ms_per_step'], (j['e2e'] or {}).get('value'), j['step_wall_ms_rank0'], j['clocks'])
    except Exception as e: print(f, 'ERR', e)
PY
