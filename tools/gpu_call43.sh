#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out; mkdir -p $O
for i in 1 2; do
SDTGPU_TRACE=1 timeout 600 python bench.py --config C5 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-parity > $O/c43.json 2> $O/c43_$i.err
python - <<'PY'
import json
j=json.loads(open('gpurun_out/c43.json').read().strip().splitlines()[-1])
print(round(j['value']/1e9,2), round(j['ms_per_step'],2), j['step_wall_ms_rank0'])
PY
grep -c "sdtgpu" $O/c43_$i.err
done
