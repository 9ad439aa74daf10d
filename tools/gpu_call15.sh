#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out; mkdir -p $O
summ() { python - "$1" <<'PY'
import json,sys
f=sys.argv[1]
try:
    j=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, round(j['value']/1e9,2), round(j['ms_per_step'],2), 'e2e', j['e2e'] and round(j['e2e']['value']/1e9,2), 'parity', j.get('parity_checked'), 'coll', j.get('collective_ms_per_step'))
    ph=j['roofline']['sliced']['phases']
    print({k:round(v['ms_per_step'],2) for k,v in ph.items()}, j['roofline']['sliced']['geometry'])
    print('cpu', j.get('cpu_baseline') and j['cpu_baseline']['value']/1e6, 'files', j.get('e2e_from_files'))
except Exception as e: print(f, 'ERR', e)
PY
}
timeout 900 python bench.py --steps 3 --warmup 2 > $O/c15_bench_C2.json 2> $O/c15_bench_C2.err; tail -n 5 $O/c15_bench_C2.err | cut -c1-300; summ $O/c15_bench_C2.json
for C in C3 C4 C5; do
timeout 900 python bench.py --config $C --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > $O/c15_bench_$C.json 2> $O/c15_bench_$C.err; tail -n 3 $O/c15_bench_$C.err | cut -c1-300; summ $O/c15_bench_$C.json
done
