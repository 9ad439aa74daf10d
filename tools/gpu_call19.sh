#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out; mkdir -p $O
summ() { python - "$1" "$2" <<'PY'
import json,sys
f=sys.argv[1]
try:
    j=json.loads(open(f).read().strip().splitlines()[-1])
    ph=j['roofline']['sliced']['phases']
    g=j['roofline']['sliced']['geometry']
    print(sys.argv[2], round(j['value']/1e9,2), round(j['ms_per_step'],2), {k:round(v['ms_per_step'],2) for k,v in ph.items()}, g['n_slices'], g['work_items'], g['retried_items'], 'frac', round(j['roofline']['frac'],3))
except Exception as e: print(f, 'ERR', e)
PY
}
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-parity > $O/c19_a.json 2> $O/c19_a.err; tail -n 2 $O/c19_a.err | cut -c1-200; summ $O/c19_a.json "hintfree"
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-parity --hint 983000000 > $O/c19_b.json 2> $O/c19_b.err; tail -n 2 $O/c19_b.err | cut -c1-200; summ $O/c19_b.json "closed-form hint"
timeout 1500 python -m pytest tests -x -q -m gpu > $O/c19_tests.log 2>&1
tail -8 $O/c19_tests.log
