#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out; mkdir -p $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 24 --warmup 3 --no-cpu-baseline --no-e2e --no-parity > $O/c47.json 2> $O/c47.err
python - <<'PY'
import json
j=json.loads(open('gpurun_out/c47.json').read().strip().splitlines()[-1])
print(round(j['value']/1e9,2), round(j['ms_per_step'],2))
prev=None
for w, e, v in zip(j["step_wall_ms_rank0"], j["exchange_host_ms_steps_rank0"], j["step_events_rank0"]):
    d = {k: round(v[4][k] - (prev or {}).get(k, 0), 1) for k in v[4]} if prev else None
    prev = v[4]
    print(w, e, v[:4], d)
PY
