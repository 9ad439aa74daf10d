#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out; mkdir -p $O
for i in 1 2 3 4; do
timeout 600 python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-e2e --no-parity > $O/c46.json 2> $O/c46.err
python - <<'PY'
import json
j=json.loads(open('gpurun_out/c46.json').read().strip().splitlines()[-1])
g=j['roofline']['sliced']['geometry']
print(round(j['value']/1e9,2), round(j['ms_per_step'],2), j['step_wall_ms_rank0'], g['epochs_emitted_again'], g['device_allocations'])
PY
done
