#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu -k "skm or sliced" > $O/c45_tests.log 2>&1
tail -3 $O/c45_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 4 --warmup 3 --no-cpu-baseline --no-e2e > $O/c45.json 2> $O/c45.err
tail -n 2 $O/c45.err | cut -c1-200
python - <<'PY'
import json
try:
    j=json.loads(open('gpurun_out/c45.json').read().strip().splitlines()[-1])
    ph=j['roofline']['sliced']['phases']
    print(round(j['value']/1e9,2), round(j['ms_per_step'],2), j['parity_checked'], {k:round(v['ms_per_step'],2) for k,v in ph.items()}, j['collective_ms_per_step'], j['step_wall_ms_rank0'])
except Exception as e: print('ERR', e)
PY
