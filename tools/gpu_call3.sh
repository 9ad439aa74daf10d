#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out; mkdir -p $O
timeout 400 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > $O/c3_bench_new.json 2> $O/c3_bench_new.err
tail -c 600 $O/c3_bench_new.err
python - <<'PY'
import json
for f in ('gpurun_out/c3_bench_new.json',):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, j['value']/1e9, j['ms_per_step'])
        ph=j['roofline']['sliced']['phases']
        print({k:round(v['ms_per_step'],2) for k,v in ph.items()}, j['roofline']['sliced']['geometry'])
        print(j['roofline']['sliced'].get('build_phase_cycles'))
    except Exception as e: print(f, 'ERR', e)
PY
timeout 500 ncu --set full --clock-control none --import-source on -k "regex:skm_build" -s 2 -c 1 -f -o $O/c3_build python bench.py --pairs 4000000 --transcripts 3200 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > $O/c3_ncu.log 2>&1
