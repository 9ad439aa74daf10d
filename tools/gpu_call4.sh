#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_sliced.py -x -q -m gpu > $O/c4_tests.log 2>&1
tail -5 $O/c4_tests.log
SDTGPU_ORD64=1 timeout 300 python -m pytest tests/test_gpu_sliced.py -x -q -m gpu -k "parity_ragged or hot" > $O/c4_tests64.log 2>&1
tail -3 $O/c4_tests64.log
summ() { python - "$1" <<'PY'
import json,sys
f=sys.argv[1]
try:
    j=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, round(j['value']/1e9,2), round(j['ms_per_step'],2))
    ph=j['roofline']['sliced']['phases']
    print({k:round(v['ms_per_step'],2) for k,v in ph.items()}, j['roofline']['sliced']['geometry']['slice_slots'], j['roofline']['sliced']['geometry']['retried_items'])
except Exception as e: print(f, 'ERR', e)
PY
}
timeout 400 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > $O/c4_bench.json 2> $O/c4_bench.err; tail -c 400 $O/c4_bench.err; summ $O/c4_bench.json
SDTGPU_ORD64=1 timeout 400 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > $O/c4_bench64.json 2> $O/c4_bench64.err; summ $O/c4_bench64.json
