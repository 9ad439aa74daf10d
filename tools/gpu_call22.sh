#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out; mkdir -p $O
for C in C1 C3 C4 C5; do
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
timeout 900 python bench.py --files-full --steps 3 --warmup 3 > $O/r2_bench_c2_files.json 2> $O/r2_bench_c2_files.err; tail -n 2 $O/r2_bench_c2_files.err | cut -c1-300
python - <<'PY'
import json
j=json.loads(open('gpurun_out/r2_bench_c2_files.json').read().strip().splitlines()[-1])
print(round(j['value']/1e9,2), j['e2e_from_files'])
PY
