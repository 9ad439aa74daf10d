#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > $O/c16_multi.log 2>&1
tail -25 $O/c16_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 1 --no-cpu-baseline > $O/c16_bench_2gpu.json 2> $O/c16_bench_2gpu.err
tail -n 5 $O/c16_bench_2gpu.err | cut -c1-300
python - <<'PY'
import json
try:
    j=json.loads(open('gpurun_out/c16_bench_2gpu.json').read().strip().splitlines()[-1])
    print(round(j['value']/1e9,2), round(j['ms_per_step'],2), 'e2e', j['e2e'] and round(j['e2e']['value']/1e9,2), 'coll', j.get('collective_ms_per_step'))
    ph=j['roofline']['sliced']['phases']
    print({k:round(v['ms_per_step'],2) for k,v in ph.items()}, j['roofline']['sliced']['geometry'])
except Exception as e: print('ERR', e)
PY
