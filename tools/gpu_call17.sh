#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out; mkdir -p $O
N=${1:-2}
[ -n "$TRACE" ] && export SDTGPU_TRACE=1
timeout 900 env BENCH_NO_SAMPLER=$NOSAMP BENCH_TRACE_RANK0=$TR0 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline > $O/c17_bench_${N}gpu.json 2> $O/c17_bench_${N}gpu.err
grep "sdtgpu\]" $O/c17_bench_${N}gpu.err | tail -n 24 | cut -c1-200
tail -n 3 $O/c17_bench_${N}gpu.err | cut -c1-300
python - $N <<'PY'
import json,sys
try:
    j=json.loads(open(f'gpurun_out/c17_bench_{sys.argv[1]}gpu.json').read().strip().splitlines()[-1])
    print(round(j['value']/1e9,2), round(j['ms_per_step'],2), 'e2e', j['e2e'] and round(j['e2e']['value']/1e9,2), 'coll', j.get('collective_ms_per_step'))
    ph=j['roofline']['sliced']['phases']
    print({k:round(v['ms_per_step'],2) for k,v in ph.items()}, j['roofline']['sliced']['geometry'])
except Exception as e: print('ERR', e)
PY
python - $N <<'PY'
import json,sys
j=json.loads(open(f'gpurun_out/c17_bench_{sys.argv[1]}gpu.json').read().strip().splitlines()[-1])
print('host phases rank0 (includes e2e steps):', j.get('exchange_host_ms_per_step_rank0'))
PY
