#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out; mkdir -p $O
summ() { python - "$1" <<'PY'
import json,sys
f=sys.argv[1]
try:
    j=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, round(j['value']/1e9,2), round(j['ms_per_step'],2), 'e2e', j['e2e'] and round(j['e2e']['value']/1e9,2))
    ph=j['roofline']['sliced']['phases']
    print({k:round(v['ms_per_step'],2) for k,v in ph.items()}, j['roofline']['sliced']['geometry'])
except Exception as e: print(f, 'ERR', e)
PY
}
for C in C2 C3; do
timeout 600 python bench.py --config $C --path sliced --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > $O/c14_bench_$C.json 2> $O/c14_bench_$C.err; tail -n 3 $O/c14_bench_$C.err | cut -c1-250; summ $O/c14_bench_$C.json
done
SDTGPU_TRACE=1 timeout 600 python bench.py --config C4 --path sliced --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > $O/c14_bench_C4.json 2> $O/c14_bench_C4.err; tail -n 25 $O/c14_bench_C4.err | cut -c1-250; summ $O/c14_bench_C4.json
