#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out; mkdir -p $O
run() { timeout 900 env $1 python bench.py --config C3 --steps 2 --warmup 2 --no-cpu-baseline --no-e2e --no-parity > $O/c25.json 2> $O/c25.err; tail -n 1 $O/c25.err | cut -c1-200
python - "$1" <<'PY'
import json,sys
try:
    j=json.loads(open('gpurun_out/c25.json').read().strip().splitlines()[-1])
    ph=j['roofline']['sliced']['phases']; g=j['roofline']['sliced']['geometry']
    print(sys.argv[1], round(j['value']/1e9,2), round(j['ms_per_step'],2), {k:round(v['ms_per_step'],2) for k,v in ph.items()}, g['n_slices'], g['work_items'], g['retried_items'])
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
}
run "SDTGPU_SLICE_LOAD=0.12"
run "SDTGPU_SLICE_LOAD=0.25"
run "SDTGPU_SLICE_LOAD=0.4"
run "SDTGPU_SLICE_LOAD=0.25 SDTGPU_ITEM_LOAD=1.1"
