#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_sliced.py -x -q -m gpu > $O/c34_tests.log 2>&1
tail -3 $O/c34_tests.log
run() { timeout 600 env $1 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e $2 > $O/c34.json 2> $O/c34.err; tail -n 1 $O/c34.err | cut -c1-200
python - "$1" <<'PY'
import json,sys
try:
    j=json.loads(open('gpurun_out/c34.json').read().strip().splitlines()[-1])
    ph=j['roofline']['sliced']['phases']; g=j['roofline']['sliced']['geometry']
    print(sys.argv[1], round(j['value']/1e9,2), round(j['ms_per_step'],2), j['parity_checked'], {k:round(v['ms_per_step'],2) for k,v in ph.items()}, g['n_slices'], g['work_items'], g['retried_items'], g['n_records_merged'])
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
}
run "X=1" ""
run "X=2" "--no-parity"
