#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_sliced.py -x -q -m gpu > $O/c39_tests.log 2>&1
tail -3 $O/c39_tests.log
run() { timeout 600 env $1 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e $2 > $O/c39.json 2> $O/c39.err; tail -n 1 $O/c39.err | cut -c1-200
python - "$1" <<'PY'
import json,sys
try:
    j=json.loads(open('gpurun_out/c39.json').read().strip().splitlines()[-1])
    ph=j['roofline']['sliced']['phases']; g=j['roofline']['sliced']['geometry']
    print(sys.argv[1], round(j['value']/1e9,2), round(j['ms_per_step'],2), j['parity_checked'], {k:round(v['ms_per_step'],2) for k,v in ph.items()}, g['n_slices'], g['work_items'], g['retried_items'], g['n_records_merged'], j['step_wall_ms_rank0'])
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
}
run "X=1" ""

run "X=C5" "--config C5"
