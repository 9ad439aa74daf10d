#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_sliced.py -x -q -m gpu > $O/c27_tests.log 2>&1
tail -3 $O/c27_tests.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/c27_a.json 2> $O/c27_a.err; tail -n 2 $O/c27_a.err | cut -c1-200
python - <<'PY'
import json
j=json.loads(open('gpurun_out/c27_a.json').read().strip().splitlines()[-1])
ph=j['roofline']['sliced']['phases']
print(round(j['value']/1e9,2), round(j['ms_per_step'],2), 'e2e', round(j['e2e']['value']/1e9,2), {k:round(v['ms_per_step'],2) for k,v in ph.items()})
PY
