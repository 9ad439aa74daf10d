#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_sliced.py tests/test_gpu_parity.py -x -q -m gpu > $O/c2_tests.log 2>&1
tail -5 $O/c2_tests.log
timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/c2_bench_new.json 2> $O/c2_bench_new.err
tail -c 600 $O/c2_bench_new.err
python - <<'PY'
import json
for f in ('gpurun_out/c2_bench_new.json',):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, j['value']/1e9, j['ms_per_step'])
        ph=j['roofline']['sliced']['phases']
        print({k:round(v['ms_per_step'],2) for k,v in ph.items()}, j['roofline']['sliced']['geometry'])
    except Exception as e: print(f, 'ERR', e)
PY
