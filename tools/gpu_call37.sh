#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out; mkdir -p $O
for i in 1 2 3; do
timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --no-parity > $O/c37.json 2> $O/c37.err
python - <<'PY'
import json
j=json.loads(open('gpurun_out/c37.json').read().strip().splitlines()[-1])
print(round(j['value']/1e9,2), round(j['ms_per_step'],2), j['step_wall_ms_rank0'], j['clocks'])
PY
done
SDTGPU_TRACE=1 timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --no-parity > $O/c37.json 2> $O/c37t.err
grep -E "total|emit|build level" $O/c37t.err | cut -c1-120 | tail -20
