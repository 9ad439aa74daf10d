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
    print(sys.argv[2], round(j['value']/1e9,2), round(j['ms_per_step'],2), {k:round(v['ms_per_step'],2) for k,v in ph.items()}, g['n_slices'], g['work_items'], g['retried_items'])
except Exception as e: print(f, 'ERR', e)
PY
}
for LOAD in 0.12 0.2 0.3 0.45; do
SDTGPU_SLICE_LOAD=$LOAD timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --no-parity > $O/c18_a.json 2> $O/c18_a.err; tail -n 2 $O/c18_a.err | cut -c1-200; summ $O/c18_a.json "hintfree load=$LOAD"
done
for IL in 0.9 1.2; do
SDTGPU_ITEM_LOAD=$IL SDTGPU_SLICE_LOAD=0.2 timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --no-parity > $O/c18_a.json 2> $O/c18_a.err; summ $O/c18_a.json "hintfree load=0.2 item_load=$IL"
done
for C in C3 C4; do
timeout 900 python bench.py --config $C --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > $O/c18_bench_$C.json 2> $O/c18_bench_$C.err; tail -n 2 $O/c18_bench_$C.err | cut -c1-300; summ $O/c18_bench_$C.json $C
done
