#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out; mkdir -p $O
SDTGPU_TRACE=1 timeout 400 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > $O/c7_bench.json 2> $O/c7_bench.err; tail -n 60 $O/c7_bench.err
python - <<'PY'
import json
j=json.loads(open('gpurun_out/c7_bench.json').read().strip().splitlines()[-1])
print(round(j['value']/1e9,2), round(j['ms_per_step'],2))
ph=j['roofline']['sliced']['phases']
print({k:round(v['ms_per_step'],2) for k,v in ph.items()})
PY
SDTGPU_NO_SPLIT=1 timeout 400 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > $O/c7_bench_ns.json 2> $O/c7_bench_ns.err
python - <<'PY'
import json
j=json.loads(open('gpurun_out/c7_bench_ns.json').read().strip().splitlines()[-1])
print('nosplit', round(j['value']/1e9,2), round(j['ms_per_step'],2))
ph=j['roofline']['sliced']['phases']
print({k:round(v['ms_per_step'],2) for k,v in ph.items()})
PY
