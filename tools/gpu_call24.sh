#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out; mkdir -p $O
for C in C3 C4; do
timeout 900 python bench.py --config $C --steps 3 --warmup 3 --no-cpu-baseline > $O/r2_bench_$C.json 2> $O/r2_bench_$C.err; tail -n 2 $O/r2_bench_$C.err | cut -c1-300
python - $C <<'PY'
import json,sys
try:
    j=json.loads(open(f'gpurun_out/r2_bench_{sys.argv[1]}.json').read().strip().splitlines()[-1])
    ph=j['roofline']['sliced']['phases']; g=j['roofline']['sliced']['geometry']
    print(sys.argv[1], round(j['value']/1e9,2), round(j['ms_per_step'],2), 'e2e', round(j['e2e']['value']/1e9,2), 'parity', j['parity_checked'], 'frac', round(j['roofline']['frac'],3), {k:round(v['ms_per_step'],2) for k,v in ph.items()}, g['n_slices'], g['work_items'], g['retried_items'], g['n_nodes'])
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
