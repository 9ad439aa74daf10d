#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_sliced.py -x -q -m gpu > $O/c11_tests.log 2>&1
tail -4 $O/c11_tests.log
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
for C in C3 C4; do
SDTGPU_TRACE=1 timeout 600 python bench.py --config $C --path sliced --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > $O/c11_bench_$C.json 2> $O/c11_bench_$C.err; tail -n 14 $O/c11_bench_$C.err | cut -c1-250; summ $O/c11_bench_$C.json
done
